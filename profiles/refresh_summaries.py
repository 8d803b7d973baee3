"""Regenerate the committed profile summaries from the captures of `profiles/run_profiles.sh <tag>`:

    python profiles/refresh_summaries.py <tag>

reads  gpurun_out/prof_<tag>.ncu-rep, prof_tpl_<tag>.ncu-rep, launches_<tag>.csv, bench_<tag>.json
writes profiles/r1_final_ncu.md, r1_final_launches.md, r1_template_kernels_ncu.md, ncu_traffic.json
(needs `ncu` on PATH to export the raw / source pages of the reports)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(REPO, "gpurun_out")
sys.path.insert(0, HERE)
import block_summary                                                  # noqa: E402
import stall_summary                                                  # noqa: E402
import summarize_ncu                                                  # noqa: E402


def export(rep, page, dst, extra=()):
    with open(dst, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), stdout=f, stderr=subprocess.DEVNULL, check=False)


def captured(fn, *args):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn(*args)
    return buf.getvalue().splitlines()


def table(raw_csv):
    lines = [l for l in captured(summarize_ncu.main, raw_csv) if l.startswith("|")]
    hdr = [c.strip() for c in lines[0].strip("|").split("|")]
    rows = [[c.strip() for c in l.strip("|").split("|")] for l in lines[2:]]
    return hdr, rows


def main(tag):
    # ---- K2
    rep = os.path.join(OUT, "prof_%s.ncu-rep" % tag)
    raw, src = os.path.join(OUT, "raw_%s.csv" % tag), os.path.join(OUT, "src_%s_scan.csv" % tag)
    export(rep, "raw", raw)
    export(rep, "source", src, ["--launch-skip", "0", "--launch-count", "1"])
    hdr, rows = table(raw)
    rawrows = list(csv.reader(open(raw)))
    rh = rawrows[0]

    def rawv(name, i):
        return rawrows[2 + i][rh.index(name)]
    out = ["# ncu `--set full` summary, round 1 final kernels (`k_unbinned_mma<2>`: DMMA K2, 4 m-tiles per warp, tiled TMA, sticky point groups)", "",
           "Command: `profiles/run_profiles.sh %s` -> `ncu --set full --clock-control none --import-source on -k regex:k_unbinned_mma -c 2 -o gpurun_out/prof_%s python profiles/profile_driver.py 1`" % (tag, tag),
           "(after the same program exited 0 without ncu).  Column 1: config-2 scan (4096 points x 99 957 events, K = C*S = 8 terms).",
           "Column 2: one point over 8 Mi events (HBM-bound regime).  Condensed by `profiles/refresh_summaries.py`", "",
           "| metric | scan | P = 1 stream |", "|---|---|---|"]
    for i, h in enumerate(hdr[1:], 1):
        out.append("| %s | %s | %s |" % (h, rows[0][i], rows[1][i]))
    pipe = 'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active'
    out.append("| shared FP64/DMMA pipe cycles active %% (sm__pipe_shared_cycles_active) | %s | %s |" % (rawv(pipe, 0), rawv(pipe, 1)))
    out.append("| warps active per scheduler | %s | %s |" % (rawv('smsp__warps_active.avg.per_cycle_active', 0), rawv('smsp__warps_active.avg.per_cycle_active', 1)))
    out += ["", "## Where the warps are (scan; samples by basic block)", "", "```"] + captured(block_summary.main, src)[:12] + ["```", "",
            "The first block is the branch-free group body of a full unit (32 DMMA.8x8x4 + 32 DMUL + range checks for 32 points x 32",
            "events); the rest is per-tile / per-superblock / per-group overhead and the lighter remainder units.", "",
            "## Stall reasons and hottest instructions (scan)", "", "```"] + [l[:230] for l in captured(stall_summary.main, src)[:22]] + ["```", "",
            "Reading: the FP64/DMMA pipe is busy ~72 % of the time (math-pipe throttle is the top stall, then the fixed-latency `wait` of the",
            "dependent product-tree `DMUL`s); 2 * K flops of every 2 * K + 2 pipe slots are algorithmic, so 0.72 * 0.89 ~ 0.63 is what this",
            "schedule can reach -- bench.py measures 0.61.  DRAM traffic per scan launch against 40.0 MB algorithmic: see the table;",
            "at P = 1 the kernel moves ~543 MB for 537 MB algorithmic (bench.py, CUDA events: 0.82-0.83 of the measured HBM peak)."]
    open(os.path.join(HERE, "r1_final_ncu.md"), "w").write("\n".join(out) + "\n")
    traffic = {}
    for key, i in (("scan", 0), ("p1_stream", 1)):
        rd, wr = float(rawv('dram__bytes_read.sum', i)) * 1e6, float(rawv('dram__bytes_write.sum', i)) * 1e6
        traffic[key] = {"kernel": "k_unbinned_mma<2>", "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                        "duration_us_under_ncu": float(rawv('gpu__time_duration.sum', i))}
    traffic["_how"] = ("profiles/run_profiles.sh %s: ncu --set full --clock-control none --import-source on -k regex:k_unbinned_mma -c 2 "
                       "python profiles/profile_driver.py 1; scan = config-2 4096-point scan (algorithmic bytes 40.0 MB: the anchor tensor "
                       "once), p1_stream = one point over 8 Mi events (algorithmic 536.9 MB)" % tag)
    json.dump(traffic, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)

    # ---- template kernels
    rep = os.path.join(OUT, "prof_tpl_%s.ncu-rep" % tag)
    raw = os.path.join(OUT, "raw_tpl_%s.csv" % tag)
    export(rep, "raw", raw)
    hdr, rows = table(raw)
    keep = [rows[0], rows[-1]]
    out = ["# ncu `--set full` summary of the template-space kernels (round 1, final)", "",
           "Command: `profiles/run_profiles.sh %s` -> `ncu --set full --clock-control none --import-source on -k regex:\"k_mixture_partials|k_template_partials\" -c 4 -o gpurun_out/prof_tpl_%s python profiles/template_profile.py`" % (tag, tag),
           "(after the same program exited 0 without ncu).  Column 1: K5b `k_mixture_partials<1,2>`, config-5 shape, P = 1 over 1e8 events",
           "(2.0 GB of prepared events).  Column 2: K5 `k_template_partials<1,2>` (packed templates, 256-bit gathers), config-4 shape,",
           "1e5 toys x ~1000 events, one point per toy.", "",
           "| metric | K5b mixture stream (P = 1, 1e8 events) | K5 toys (1e5 toys) |", "|---|---|---|"]
    for i, h in enumerate(hdr[1:], 1):
        out.append("| %s | %s | %s |" % (h, keep[0][i], keep[1][i]))
    out += ["", "K5b: 2.0 GB DRAM read for 2.0 GB algorithmic (20 B per prepared event); duration under ncu (cold caches, serialised) vs",
            "~0.345 ms measured by bench.py with CUDA events: 5.8 TB/s = 0.88 of the measured HBM copy bandwidth (6.55 TB/s).",
            "K5 (toys): bound by scattered L2 requests; the packed template layout brings all four lookup corners of a row with one",
            "256-bit load (24 requests per event at config 4; 20.2 ms for the same sweep with 96 64-bit gathers, 16.3 ms with 48 128-bit ones)."]
    open(os.path.join(HERE, "r1_template_kernels_ncu.md"), "w").write("\n".join(out) + "\n")

    # ---- launch list of the bench command
    rows = [r for r in csv.reader(open(os.path.join(OUT, "launches_%s.csv" % tag), errors="ignore")) if len(r) > 10]
    h = rows[0]
    ik, iv, ig, ib = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ik], [0, 0.0, None, None])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1000.0
        a[2], a[3] = r[ig], r[ib]
    bench = json.load(open(os.path.join(OUT, "bench_%s.json" % tag)))
    lines = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --skip-cpu --skip-other` (round 1, final kernels)", "",
             "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_%s.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-other`" % tag,
             "(profiles/run_profiles.sh %s, after the same command exited 0 without ncu).  Per-launch times under ncu are cold-cache and" % tag,
             "serialised; what must agree with bench.py is the kernel's SHARE of a step.  The bench runs the 4-launch step (K1 `k_point_setup_warp`,",
             "`k_plan_units`, K2 `k_unbinned_mma<2>`, `k_unbinned_finalize_warp`) for warm-up + timed steps + the e2e arm (replayed as a CUDA graph),",
             "then K2 alone, the P = 1 streaming case, and the peak probes (`--skip-other`: the config-1/3/4/5 sections have their own capture,",
             "`profiles/r1_template_kernels_ncu.md`).", "",
             "| kernel | launches | total us | avg us | grid (last) | block |", "|---|---|---|---|---|---|"]
    for k, a in agg.items():
        lines.append("| `%s` | %d | %.1f | %.1f | %s | %s |" % (k[:70].replace("|", "/"), a[0], a[1], a[1] / a[0], a[2], a[3]))
    lines += ["", "bench.py of the same pass (`gpurun_out/bench_%s.json`, CUDA events, no profiler): value %.3e point-events/s, %.3f ms per step,"
              % (tag, bench["value"], bench["ms_per_step"]),
              "K2 %.3f ms = %.2f of the step (`roofline.share_of_step`), e2e %.3e (%.3f ms per call)."
              % (bench["roofline"]["ms"], bench["roofline"]["share_of_step"], bench["e2e"]["value"], bench["e2e"]["ms_per_step"])]
    open(os.path.join(HERE, "r1_final_launches.md"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main(sys.argv[1])

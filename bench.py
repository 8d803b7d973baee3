#!/usr/bin/env python
"""Benchmark of the likelihood-evaluation hot path (BASELINE.json metric: logL evals/s = points x events / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points P] [--events N]

Workload at N=1 = BASELINE.json configs[1]: 2 sources, 2-D 100x100 histogram templates, 2 shape
nuisances x 5 anchors (G = 25), ~100k events, 4096-point profile scan.  One "step" = one pass of the
hot path over the whole scan.

  value     device-resident: parameter points and schedule already in HBM; timed region (CUDA events on
            the launching stream) = K1 point set-up + K2 fused morph/mixture/log/reduce + finalize.
  e2e       ll.batch(host ndarray) -> host ndarray through the public API: planning, H2D of points and
            schedule from pinned memory, the same kernels, D2H of logL + status, stream sync.
  roofline  dominant kernel (K2) timed alone; FP64-pipe bound for a shared-dataset scan
            (SURVEY.md section 8d), peak = FP64 FMA micro-benchmark measured in this run; the same
            kernel's algorithmic HBM bytes and the streaming (P=1, HBM-bound) regime are reported too.
  cpu_baseline  the oracle pipeline (same SciPy calls as the reference) on a bounded sample, 1 core.

The L2 (126 MB) is flushed between timed steps by writing a 512 MB buffer (the 40 MB anchor tensor
would otherwise stay L2-resident); the flush is outside the timed intervals.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import bench_workloads as wl   # noqa: E402

METRIC = "logL evals/sec (points x events / s)"
UNIT = "point-events/s"
N_SOURCES, N_SHAPE, ANCHORS, BINS = 2, 2, wl.ANCHORS_5, (100, 100)


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle pipeline on the host cores
# ------------------------------------------------------------------------------------------------
_ORACLE = None
_REF_LL = None


def build_oracle(n_events, seed=1):
    from oracle.pipeline import UnbinnedOracle
    axes, edges, templates, mus = wl.c2_arrays(N_SOURCES, N_SHAPE, ANCHORS, BINS)
    x, y = wl.c2_events(templates, mus, edges, n_events, seed)
    t0 = time.perf_counter()
    orc = UnbinnedOracle(axes, mus).set_data_from_templates(templates, edges, [x, y])
    return orc, len(x), time.perf_counter() - t0


def build_reference_ll(n_events, seed=1):
    """The config-2 workload through the API of the UNMODIFIED reference package (oracle/_ref: a copy of
    /root/reference/blueice made by __graft_entry__.build(), with the two test-only stand-ins for multihist /
    atomicwrites).  Returns (lf, n_events, set_data seconds) or None when oracle/_ref is absent."""
    from oracle.build_ref import import_reference
    if import_reference() is None:
        return None
    from blueice.likelihood import UnbinnedLogLikelihood          # the reference
    from blueice.source import HistogramPdfSource
    from multihist import Histdd
    os.chdir(tempfile.mkdtemp(prefix="bi_ref_"))                  # the reference writes ./pdf_cache
    axes, edges, templates, mus = wl.c2_arrays(N_SOURCES, N_SHAPE, ANCHORS, BINS)
    space, params = ['cs1', 'cs2'], ['shift%d' % (i + 1) for i in range(N_SHAPE)]
    cls = wl.array_source_class(HistogramPdfSource, Histdd, axes, edges, space, mus, templates, None, params)
    lf = UnbinnedLogLikelihood(wl.array_model_config(cls, edges, space, N_SOURCES, params, 'linear'))
    for s_ in range(N_SOURCES):
        lf.add_rate_parameter('src%d' % s_)
    for p_ in params:
        lf.add_shape_parameter(p_, ANCHORS)
    lf.prepare()
    x, y = wl.c2_events(templates, mus, edges, n_events, seed)
    d = np.zeros(len(x), dtype=[('cs1', float), ('cs2', float), ('source', int)])
    d['cs1'], d['cs2'] = x, y
    t0 = time.perf_counter()
    lf.set_data(d)
    return lf, len(x), time.perf_counter() - t0


def _reference_points(lf, zs, mult):
    out = np.empty(len(zs))
    for i in range(len(zs)):
        kw = {'shift%d' % (j + 1): float(zs[i, j]) for j in range(N_SHAPE)}
        kw.update({'src%d_rate_multiplier' % j: float(mult[i, j]) for j in range(N_SOURCES)})
        out[i] = lf(**kw)
    return out


def _oracle_chunk(args):
    zs, mult = args
    if _REF_LL is not None:
        return _reference_points(_REF_LL, zs, mult)
    return _ORACLE.batch(zs, mult)


def cpu_baseline_single_core(n_events, budget_s=12.0):
    """The reference evaluation path on ONE core (it is single-threaded): the unmodified reference package when
    oracle/_ref is present (kind 'reference'), else the oracle port (kind 'port')."""
    zs, mult = wl.scan_points(4096, N_SHAPE, N_SOURCES, seed=2)
    ref = build_reference_ll(n_events)
    if ref is not None:
        lf, n, set_data_s = ref
        kind, what = "reference", "unmodified reference package (oracle/_ref) called point by point, lf(**params)"

        def one(i):
            return _reference_points(lf, zs[i:i + 1], mult[i:i + 1])[0]
    else:
        orc, n, set_data_s = build_oracle(n_events)
        kind, what = "port", "oracle.pipeline.UnbinnedOracle (scipy RegularGridInterpolator + NumPy, as the reference)"

        def one(i):
            return orc(zs[i], mult[i])
    one(0)                                                  # warm-up
    done, t0 = 0, time.perf_counter()
    while done < len(zs) and time.perf_counter() - t0 < budget_s:
        one(done)
        done += 1
    dt = time.perf_counter() - t0
    return {"value": done * n / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d of the 4096 scan points x %d events, %.1f s; %s; set_data %.2f s"
                      % (done, n, dt, what, set_data_s),
            "ms_per_point": 1e3 * dt / done, "set_data_s": set_data_s}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores -- the unmodified
    reference package (oracle/_ref) when present, else the oracle port -- on the own arm's workload: every step is the
    whole 4096-point scan of config 2, its points dealt out to one worker process per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    global _ORACLE, _REF_LL
    n_events = args.events
    ref = build_reference_ll(n_events)
    check = None
    if ref is not None:
        _REF_LL, n, set_data_s = ref
        kind = "reference"
        what = "unmodified reference package (oracle/_ref: /root/reference/blueice + stand-ins for multihist / atomicwrites)"
        # the oracle restatement must give the reference's numbers (bit for bit) on this very workload
        orc, _, _ = build_oracle(n_events)
        z4, m4 = wl.scan_points(4, N_SHAPE, N_SOURCES, seed=2)
        check = bool(np.array_equal(_reference_points(_REF_LL, z4, m4), orc.batch(z4, m4)))
    else:
        _ORACLE, n, set_data_s = build_oracle(n_events)
        kind, what = "port", "oracle.pipeline.UnbinnedOracle"
    cores = os.cpu_count() or 1
    n_points = args.points
    zs, mult = wl.scan_points(n_points, N_SHAPE, N_SOURCES, seed=2)
    chunks = [(zs[i::cores], mult[i::cores]) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(min(args.warmup, 1)):                    # a CPU has no clocks to ramp: one warm-up pass
            pool.map(_oracle_chunk, chunks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_oracle_chunk, chunks)
        dt = time.perf_counter() - t0
    value = args.steps * len(zs) * n / dt
    sample = ("%d points x %d events per step (the whole scan) over %d worker processes (disjoint point chunks); %s; "
              "set_data %.2f s" % (len(zs), n, cores, what, set_data_s))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(n, len(zs)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "reference_equals_oracle_port": check},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def ncu_traffic(kernel_key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel, taken from
    the committed `ncu --set full` capture (profiles/ncu_traffic.json; the capture command is recorded there)."""
    path = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(kernel_key, {}).get("dram_bytes_per_launch")


def workload_config(n_events, n_points):
    return {"workload": "config 2: 2-source 2D (cs1,cs2) HistogramPdfSource templates 100x100 bins, "
                        "2 shape nuisances x 5 anchors, ~100k events, 4096-point profile scan",
            "n_events": int(n_events), "n_points_per_gpu": int(n_points), "n_sources": N_SOURCES,
            "n_anchors": len(ANCHORS) ** N_SHAPE, "lookup": "linear",
            "l2": "flushed between timed steps (512 MB write, outside the timed intervals)",
            "parallelism": "weak: every GPU evaluates its own 4096 points of the scan on a replicated dataset (point sharding "
                           "has no data-path collective); device arm: each step replays the four launches of the evaluation as one CUDA "
                           "graph and stores its results on every rank over NVLink (bi_peer_broadcast), one barrier after "
                           "the timed loop; e2e arm: PointShardedLikelihood.batch, "
                           "every call ends with one bi_peer_exchange launch (stores + flags + wait + delivery to pinned "
                           "host memory) inside the call's CUDA graph; e2e also carries the strongly scaled scan, the "
                           "event-sharded config 5 and the toy-sharded config 4 as flat keys"}


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """Samples SM clock and throttle reasons with NVML every 20 ms in a thread during the timed region."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = threading.Event()
        self.thread = None
        self.sm_max = None
        self.error = None

    def prepare(self):
        """NVML init + device handle (tens of ms): done before the timed region, not inside the sampling thread."""
        try:
            import pynvml
            import torch
            self.nvml = pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES-style remapping by matching the PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(self.gpu_index).pci_bus_id
            handle = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                    handle = h
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.handle = handle
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        except Exception as exc:                                   # pragma: no cover
            self.error = repr(exc)

    def _sample(self):
        pynvml, handle = self.nvml, self.handle
        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
        mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle))
        for name, bit in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)
        try:
            self.power.append(pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0)
        except Exception:
            pass

    def _run(self):
        try:
            while not self.stop_flag.is_set():
                self._sample()
                time.sleep(0.002)
        except Exception as exc:                                   # pragma: no cover
            self.error = repr(exc)

    def start(self):
        if getattr(self, "handle", None) is None:
            self.prepare()
        if self.error:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=5)
            if not self.error:
                try:
                    self._sample()                                 # at least one sample right at the end of the region
                except Exception as exc:                           # pragma: no cover
                    self.error = repr(exc)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
               "samples": len(self.samples), "reasons": sorted(self.reasons),
               "power_w_max": max(self.power) if self.power else None}
        if self.error:
            out["error"] = self.error
        return out


# ------------------------------------------------------------------------------------------------
# own arm
_PAD = {}


def _tight_barrier(world):
    """Barrier in front of a timed e2e call at N > 1: the NCCL barrier, then a signal-pad barrier over NVLink peer memory
    (torch symmetric memory, the PeerGather plumbing) followed by a stream synchronize.  The ranks leave it within ~2 us of
    each other; after dist.barrier() alone they leave ~6 us apart with outliers of several hundred us (measured at N = 8,
    profiles/r2/diag_ngpu.py), and a wall clock started behind it charges that skew to the call."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    dist.barrier()
    pad = _PAD.get("pad")
    if pad is None:
        from blueice_b200.distributed import PeerGather
        pad = _PAD["pad"] = PeerGather(1)
    pad.barrier()
    torch.cuda.synchronize()


def _median_max_over_ranks(fn, n_warm, n_rep, world, device):
    """Wall time of fn() (median of n_rep calls after n_warm warm-ups; every call starts behind a barrier), the maximum
    over the ranks, and the same for the device span (CUDA events around the call)."""
    import torch
    import torch.distributed as dist
    for _ in range(n_warm):
        fn()
    wall, dev = [], []
    for _ in range(n_rep):
        torch.cuda.synchronize()
        _tight_barrier(world)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        fn()
        b.record()
        wall.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        dev.append(a.elapsed_time(b) * 1e-3)
    t = torch.tensor([float(np.median(wall)), float(np.median(dev))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), float(t[1])


def config5_event_sharded(args, rank, world, device):
    """BASELINE config 5 as north_star splits it: `--c5-events` events IN TOTAL (strong scaling), sharded over the ranks
    in superblock-aligned contiguous slices; every minimiser step evaluates its P points on the rank's shard (mixture
    engine, K5b) and ends with ONE exchange launch (bi_peer_exchange: the shards' log sums to every rank over NVLink peer
    memory, summed in rank order on the device), all inside one CUDA graph per step."""
    import torch
    import torch.distributed as dist
    from blueice_b200 import distributed as bdist
    ll, _, names = wl.c2_api(6, 4, wl.ANCHORS_5, BINS, n_events=1000, seed=5,
                             likelihood_config={'unbinned_engine': 'mixture'})
    base_mu = float(np.sum(ll.base_model.expected_events()))
    n_target = args.c5_events
    lt = n_target / base_mu
    t0 = time.perf_counter()
    td = ll.base_model.simulate_toys(1, livetime_days=lt, seed=50)          # the same events on every rank (Philox)
    N = td.n_events
    lo, hi = bdist.shard_bounds(N, world, align=512)[rank]
    d = np.zeros(hi - lo, dtype=[('source', int)] + [(name, float) for name in td.dims])
    host = td.coords[:, lo:hi].cpu().numpy()
    for k, name in enumerate(td.dims):
        d[name] = host[k]
    del td, host
    torch.cuda.empty_cache()
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ll.set_data(d)
    torch.cuda.synchronize()
    set_data_s = time.perf_counter() - t0
    sharded = bdist.EventShardedLikelihood(ll)
    rng = np.random.default_rng(51)
    x0 = np.concatenate([rng.uniform(0.8, 1.2, size=6), rng.uniform(-1.9, 1.9, size=4)])
    fd = np.repeat(x0[None, :], 11, 0)
    for j in range(10):
        fd[j + 1, j] += 1.4901161193847656e-08
    scan = np.column_stack([rng.uniform(0.8, 1.2, size=(64, 6)), rng.uniform(-1.9, 1.9, size=(64, 4))])
    res = {}
    for P, table in ((1, fd[:1]), (11, fd), (64, scan)):
        got = sharded.batch(table, names, livetime_days=lt)
        rows = bdist.all_gather_rows(got)
        assert all(np.array_equal(rows[0], r) for r in rows), "ranks disagree on the event-sharded result"
        wall, dev = _median_max_over_ranks(lambda: sharded.batch(table, names, livetime_days=lt), 4, 20, world, device)
        res["P%d" % P] = {"e2e_ms": wall * 1e3, "device_span_ms": dev * 1e3, "point_events_per_s_e2e": P * N / wall,
                          "finite": bool(np.all(np.isfinite(got))), "logl0": float(got[0])}
    # the exchange launch alone, all ranks in the same loop (P = 11 values per rank)
    pg = sharded._gathers[11]
    x = torch.zeros(11, dtype=torch.float64, device=device)
    for _ in range(10):
        pg.reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        pg.reduce(x)
    b.record()
    torch.cuda.synchronize()
    exchange_us = a.elapsed_time(b) * 1e3 / 200
    err = pg.error()
    return {"workload": "large-dataset fit, events sharded: %d events in total over %d GPUs (%d on this rank), 6 sources, "
                        "4 shape parameters x 5 anchors (625 anchors), 100x100 templates" % (N, world, hi - lo),
            "api": "distributed.EventShardedLikelihood.batch", "n_events_total": int(N), "n_gpus": world,
            "exchange": "bi_peer_exchange mode sum (%s)" % ("nvlink peer memory" if pg.fallback is None else "nccl fallback: " + pg.fallback),
            "exchange_launch_us": exchange_us, "exchange_error_word": int(err), **res,
            "generate_s": gen_s, "set_data_s": set_data_s}


def sharded_summary(world, P, n_events, e2e_obj, strong, other):
    """Flat scalars of the three multi-GPU splits for the e2e object: config 2 strongly scaled (P points in total), config 5
    event-sharded (all events in total, P = 11 minimiser step including the exchange), config 4 toy-sharded."""
    out = {}
    if world == 1:
        out["c2_strong_points_total"] = P
        out["c2_strong_ms"] = e2e_obj["ms_per_step"]
        out["c2_strong_pe_per_s"] = e2e_obj["value"]
    elif strong is not None:
        out["c2_strong_points_total"] = strong["points_total"]
        out["c2_strong_ms"] = strong["e2e_ms"]
        out["c2_strong_pe_per_s"] = strong["point_events_per_s_e2e"]
    if not other:
        return out
    c4 = other.get("config4_toys")
    if c4:
        out["c4_toys_total"] = c4["toys_per_gpu"] * world
        out["c4_sweep_ms_e2e"] = c4["e2e"]["ms"]
        out["c4_toys_per_s_e2e"] = c4["e2e"]["toys_per_s"]
        out["c4_toys_per_s_device"] = c4["device"]["toys_per_s"]
    c5 = other.get("config5_event_sharded") or other.get("config5_large_dataset")
    if c5:
        out["c5_events_total"] = c5.get("n_events_total", c5.get("n_events"))
        for P5 in (1, 11, 64):
            r = c5.get("P%d" % P5)
            if r:
                out["c5_P%d_step_ms_e2e" % P5] = r["e2e_ms"]
        if "P11" in c5:
            out["c5_P11_pe_per_s_e2e"] = c5["P11"]["point_events_per_s_e2e"]
        if "exchange_launch_us" in c5:
            out["c5_exchange_launch_us"] = c5["exchange_launch_us"]
    return out


def long_contraction_scan(device, steps):
    """Config-2-shaped scan of a model whose contraction is long: 5 shape parameters x 3 anchors (243 anchors, 32 corners
    per hypercube cell) x 5 sources = 160 terms per point-event, 4096 points over 5e4 events (synthetic per-event pdf
    tensor, filled on the device).  Such contractions run the K-chunk DMMA kernel (k_unbinned_mma_wide); round 2 started
    with the streaming kernel here (one pass over the cell's 160 rows per POINT)."""
    import torch
    from blueice_b200.engine import MorphGrid, UnbinnedEngine
    D, S, N, P = 5, 5, 50000, 4096
    rng = np.random.default_rng(7)
    axes = [np.sort(rng.uniform(-2, 2, 3)) for _ in range(D)]
    eng = UnbinnedEngine(MorphGrid(axes), rng.uniform(5, 500, (3 ** D, S)), device=device)
    eng.allocate_ps_anchor(N)
    gen = torch.Generator(device=device)
    gen.manual_seed(7)
    eng.ps_anchor.copy_(torch.exp(torch.randn(eng.ps_anchor.shape, generator=gen, device=device, dtype=torch.float64) * 2 - 4))
    zs = np.column_stack([rng.uniform(a[0], a[-1], P) for a in axes])
    mult = rng.uniform(0.5, 2, (P, S))
    z_d, m_d, _, _, _ = eng._upload_points(zs, mult, None, None)
    z_d, m_d = z_d.clone(), m_d.clone()
    outs = eng.run_fused(P, z_d, m_d, None, None)
    _, v = eng.mma_workspace(P)
    torch.cuda.synchronize()
    finite = bool(torch.isfinite(outs["logl"]).all())
    n = max(steps, 5)
    whole, k2 = [], []
    for _ in range(2 + n):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        eng.run_fused(P, z_d, m_d, None, None)                     # K1 -> schedule -> pack -> K2 -> finalize
        b.record()
        eng.mma_plan(P, v, outs["status"])
        torch.cuda.synchronize()
        a2 = torch.cuda.Event(enable_timing=True)
        a2.record()
        eng.mma_k2(v)                                              # pack + K2 alone
        c.record()
        torch.cuda.synchronize()
        whole.append(a.elapsed_time(b))
        k2.append(a2.elapsed_time(c))
    whole, k2 = float(np.mean(whole[2:])), float(np.mean(k2[2:]))
    K = S << D
    flops = 2.0 * K * P * N
    r = eng.evaluate(zs[:64], mult[:64])
    eng.force_kernel = 'stream'
    t0 = time.perf_counter()
    r_s = eng.evaluate(zs[:64], mult[:64])
    torch.cuda.synchronize()
    stream_s = time.perf_counter() - t0
    eng.force_kernel = None
    return {"workload": "4096-point scan, 5 shape parameters x 3 anchors x 5 sources = %d contraction terms, %d events "
                        "(synthetic per-event pdf tensor, %.0f MB)" % (K, N, eng.ps_anchor.numel() * 8 / 1e6),
            "kernel": "k_wide_pack_coef + k_unbinned_mma_wide (K-chunk loop: CTA-shared event tiles, both DMMA operands "
                      "staged by TMA bulk copies, accumulators carried across the chunks)",
            "n_terms": K, "n_events": N, "n_points": P, "evaluation_ms": whole, "point_events_per_s": P * N / (whole * 1e-3),
            "finite": finite, "max_abs_diff_vs_streaming_kernel_64_points": float(np.max(np.abs(r - r_s))),
            "streaming_kernel_64_points_ms": stream_s * 1e3,
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA.8x8x4)", "ms": k2, "achieved": flops / (k2 * 1e-3) / 1e12,
                         "unit": "TFLOP/s", "flops_alg_per_point_event": 2 * K, "share_of_evaluation": k2 / whole}}


def other_configs(args, rank, world, device):
    """BASELINE configs 4 and 5 through the public API on the template-space engine (K5 / K5b), bounded sizes.

    config 4  toy-MC Neyman construction: `--toys` toys PER GPU x ~1000 events (Model.simulate_toys on the device,
              counter-based RNG keyed by the global toy id), 3 sources, 3 shape parameters x 5 anchors, one
              parameter point per toy; toys are sharded over the ranks, no collective in the evaluation.
    config 5  large-dataset fit: `--c5-events` events on this GPU (the full config is 1e8 over 8 GPUs = 1.25e7 per
              GPU), 6 sources, 4 shape parameters x 5 anchors; one minimiser evaluation (P = 1) and one
              forward-difference batch (P = 11) with the mixture engine; rank 0 only."""
    import torch
    import torch.distributed as dist
    out = {}
    # ---- config 4 ----
    t0 = time.perf_counter()
    ll, _, names = wl.c2_api(3, 3, wl.ANCHORS_5, BINS, n_events=1000, seed=4)
    build_s = time.perf_counter() - t0
    base_mu = float(np.sum(ll.base_model.expected_events()))
    lt = 1000.0 / base_mu                                            # ~1000 events per toy
    T = args.toys
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:
        # toys sharded over the ranks (weak: `--toys` toys per GPU): rank r generates and holds toys [r T, (r + 1) T)
        from blueice_b200.distributed import ToyShardedLikelihood
        sharded_toys = ToyShardedLikelihood(ll)
        toys = sharded_toys.simulate(T * world, livetime_days=lt, seed=40)
        torch.cuda.synchronize()
        gen_s, load_s = time.perf_counter() - t0, 0.0                # generation + load (event preparation) together
    else:
        toys = ll.base_model.simulate_toys(T, livetime_days=lt, seed=40, first_toy=rank * T)
        torch.cuda.synchronize()
        gen_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        ll.set_toy_data(toys)
        torch.cuda.synchronize()
        load_s = time.perf_counter() - t0
    zs_all4, mult_all4 = wl.scan_points(T * world, 3, 3, seed=41)
    table_all4 = np.ascontiguousarray(np.column_stack([mult_all4, zs_all4]))
    zs, mult = zs_all4[rank * T:(rank + 1) * T], mult_all4[rank * T:(rank + 1) * T]
    table = table_all4[rank * T:(rank + 1) * T]
    if world > 1:
        def toy_call():
            return sharded_toys.batch_toys(table_all4, names, livetime_days=lt)[rank * T:(rank + 1) * T]
    else:
        def toy_call():
            return ll.batch_toys(table, names, livetime_days=lt)
    res = toy_call()
    eng = ll._toy_engine
    # device-resident: K1 + K5 + finalize on uploaded points
    zs_d, mult_d, scale_d, _, _ = eng._upload_points(zs, mult, np.full(T, lt / ll.pdf_base_config['livetime_days']), None)
    sched = eng.toy_schedule()
    dev_ms = []
    for k in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.run_one_call(T, sched, zs_d, mult_d, scale_d, None)      # bi_template_ll_batch: K1 + K5 + finalize
        b.record()
        torch.cuda.synchronize()
        if k:
            dev_ms.append(a.elapsed_time(b))
    for _ in range(3):                                             # warm-up: the third call captures the CUDA graph
        toy_call()
    e2e = []
    for _ in range(5):
        torch.cuda.synchronize()
        _tight_barrier(world)
        t0 = time.perf_counter()
        res = toy_call()
        e2e.append(time.perf_counter() - t0)
    t_dev, t_e2e = float(np.mean(dev_ms)) * 1e-3, float(np.median(e2e))
    if world > 1:
        t = torch.tensor([t_dev, t_e2e, gen_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, gen_s = float(t[0]), float(t[1]), float(t[2])
    # per-toy maximum-likelihood fits in lock step (the Neyman construction's inner loop) on a slice of the toys
    from blueice_b200.inference import bestfit_toys
    n_fit = min(args.fit_toys, T) if world == 1 else 0
    fit_info = None
    if n_fit:
        sub = ll.base_model.simulate_toys(n_fit, livetime_days=lt, seed=40, first_toy=rank * T)
        ll.set_toy_data(sub)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, fit_ll, info = bestfit_toys(ll, livetime_days=lt)
        fit_s = time.perf_counter() - t0
        fit_info = {"toys": n_fit, "free_parameters": 6, "seconds": fit_s, "toys_per_s": n_fit / fit_s,
                    "iterations": int(info["iterations"]), "likelihood_evaluations": int(info["evaluations"]),
                    "evaluations_per_s": info["evaluations"] / fit_s, "converged": int(info["converged"].sum())}
        ll.set_toy_data(toys)
    n_ev = toys.n_events
    # the sweep against HBM: SURVEY.md 8d counts 20 bytes per event (prepared event: bin + two fractions); the bin-major form
    # moves 24 (bin-sorted event) + 8 written + 8 read (density) per event.  It is bound by the L1 data pipe (shared-memory
    # row loads + the scattered per-toy records), see profiles/r2_k5c_toys_ncu.md; DRAM traffic from the committed capture.
    c4_roofline = {"bound": "hbm", "bytes_alg": 20.0 * n_ev, "achieved": 20.0 * n_ev / t_dev / 1e9, "peak": measured_peaks()[0]["hbm_gbs"],
                   "unit": "GB/s", "frac": 20.0 * n_ev / t_dev / 1e9 / measured_peaks()[0]["hbm_gbs"],
                   "traffic": ncu_traffic("c4_bin_major_density") if sched.get("bm") is not None else None,
                   "limiter": "L1 data pipe (LSU wavefronts 64 % of peak, FP64 pipe 40 %): 24 rows x 32 B of shared-memory "
                              "operands per event; FP64 floor of the reference's operation order 1.96 ms per 1e8 events",
                   "note": "whole sweep (K1, records, densities, tree, finalize) against SURVEY.md 8d's 20 B per event"}
    out["config4_toys"] = {
        "workload": "toy-MC: %d toys per GPU x ~1000 events, 3 sources, 3 shape parameters x 5 anchors (125 anchors), "
                    "100x100 templates, one parameter point per toy" % T,
        "api": "ll.batch_toys" if world == 1 else "distributed.ToyShardedLikelihood.batch_toys (toys sharded over the ranks, "
               "results of all ranks gathered over NVLink by one bi_peer_exchange launch inside the evaluation's CUDA graph)",
        "kernel": ("k_bm_density<2,3,3> + k_template_partials<1,2,pre> (densities formed bin-major: the rows of one bin staged in "
                   "shared memory by one TMA bulk copy, events bucketed by hypercube cell, per-toy records; range test + canonical "
                   "tree + rare path by K5's own code; bit-identical to K3 + K2)") if sched.get("bm") is not None else
                  "k_template_partials<1,2> (fused template lookup + morph + log-sum, bit-identical to K3 + K2)",
        "roofline": c4_roofline,
        "toys_per_gpu": T, "events_per_gpu": int(n_ev), "n_gpus": world, "finite_results": int(np.isfinite(res).sum()),
        "device": {"ms": t_dev * 1e3, "toys_per_s": world * T / t_dev, "point_events_per_s": world * n_ev / t_dev},
        "e2e": {"ms": t_e2e * 1e3, "toys_per_s": world * T / t_e2e, "point_events_per_s": world * n_ev / t_e2e,
                "h2d_bytes": int(eng.last_h2d_bytes), "d2h_bytes": int(eng.last_d2h_bytes)},
        "generate_s": gen_s, "load_s": load_s, "model_build_s": build_s,
        "extrapolated_1e6_toys_s": 1e6 / (world * T / t_e2e), "lock_step_fits": fit_info}
    del ll, toys, eng
    torch.cuda.empty_cache()
    if world > 1:
        # the other two splits of north_star, strongly scaled; configs 1 / 3 and the single-GPU config-5 breakdown
        # are reported by the N = 1 run only
        out["config5_event_sharded"] = config5_event_sharded(args, rank, world, device)
        return out
    # ---- config 1 (the reference's own CPU-runnable case): latency of one evaluation and a 4096-point batch ----
    ll1, d1, names1 = wl.c1_api(seed=0)
    rng1 = np.random.default_rng(1)
    pts1 = np.column_stack([rng1.uniform(0.5, 2.0, 4096), rng1.uniform(-2.0, 2.0, 4096)])
    kw1 = dict(zip(names1, [float(v) for v in pts1[0]]))
    for _ in range(20):
        ll1(**kw1)
    t0 = time.perf_counter()
    for _ in range(300):
        ll1(**kw1)
    lat = (time.perf_counter() - t0) / 300
    for _ in range(4):                                             # the second call captures the CUDA graph
        b1 = ll1.batch(pts1, names1)
    t0 = time.perf_counter()
    for _ in range(20):
        b1 = ll1.batch(pts1, names1)
    tb = (time.perf_counter() - t0) / 20
    out["config1_gaussian"] = {
        "workload": "conf_for_test() Gaussian-source UnbinnedLogLikelihood, shape parameter mu (3 anchors), %d events" % len(d1),
        "n_events": int(len(d1)), "single_call_latency_us": lat * 1e6, "single_call_point_events_per_s": len(d1) / lat,
        "batch_points": 4096, "batch_ms": tb * 1e3, "batch_point_events_per_s": 4096 * len(d1) / tb,
        "batch_equals_single_call": bool(b1[0] == ll1(**kw1))}
    # the caller this regime lives in: one_parameter_interval (inference.py:332-389) = a brentq search whose every step is
    # a conditional fit of sequential ll(**params) calls
    n_calls = [0]
    inner_call = type(ll1).__call__

    def counting_call(self, *a_, **k_):
        n_calls[0] += 1
        return inner_call(self, *a_, **k_)
    type(ll1).__call__ = counting_call
    limit, interval_s = float('nan'), float('nan')
    try:
        t0 = time.perf_counter()
        limit = ll1.one_parameter_interval('s0_rate_multiplier', 2.0, kind='upper')
        interval_s = time.perf_counter() - t0
    except Exception as exc:
        out["config1_gaussian"]["one_parameter_interval_error"] = repr(exc)
    finally:
        type(ll1).__call__ = inner_call
    out["config1_gaussian"]["one_parameter_interval"] = {
        "call": "ll.one_parameter_interval('s0_rate_multiplier', 2.0, kind='upper')  (mu profiled)", "upper_limit": float(limit),
        "seconds": interval_s, "likelihood_calls": n_calls[0], "us_per_call": 1e6 * interval_s / max(n_calls[0], 1)}
    if not args.skip_cpu:
        # the same search through the unmodified reference package on the same data, when oracle/_ref travelled here
        try:
            from oracle.build_ref import import_reference
            if import_reference() is not None:
                from blueice.likelihood import UnbinnedLogLikelihood as RefLL
                from blueice.test_helpers import conf_for_test as ref_conf
                os.chdir(tempfile.mkdtemp(prefix="bi_ref1_"))
                lr = RefLL(ref_conf(n_sources=1))
                lr.add_rate_parameter('s0')
                lr.add_shape_parameter('mu', {-2: -2, 0: 0, 2: 2})
                lr.prepare()
                lr.set_data(d1)
                for _ in range(20):
                    lr(**kw1)
                t0 = time.perf_counter()
                for _ in range(300):
                    lr(**kw1)
                ref_lat = (time.perf_counter() - t0) / 300
                t0 = time.perf_counter()
                ref_limit = lr.one_parameter_interval('s0_rate_multiplier', 2.0, kind='upper')
                ref_interval_s = time.perf_counter() - t0
                out["config1_gaussian"]["reference"] = {
                    "what": "unmodified reference package (oracle/_ref), same data, one core",
                    "single_call_us": ref_lat * 1e6, "one_parameter_interval_s": ref_interval_s,
                    "upper_limit": float(ref_limit), "logl_abs_diff": float(abs(lr(**kw1) - ll1(**kw1)))}
        except Exception as exc:                                    # the baseline leg must never break the bench line
            out["config1_gaussian"]["reference"] = {"error": repr(exc)}
        # cpu_baseline leg for this config: the oracle port (same third-party calls as the reference), one core
        from oracle.pipeline import UnbinnedOracle
        axes1, mus1, ps1, x1 = wl.c1_arrays(seed=0)
        orc1 = UnbinnedOracle(axes1, mus1).set_ps(ps1)
        for _ in range(20):
            orc1([0.3], [1.1])
        t0 = time.perf_counter()
        for _ in range(500):
            orc1([0.3], [1.1])
        cpu_lat = (time.perf_counter() - t0) / 500
        out["config1_gaussian"]["cpu_port_single_call_us"] = cpu_lat * 1e6
        out["config1_gaussian"]["cpu_port_events"] = int(len(x1))
    del ll1
    # ---- config 3 (binned + Beeston-Barlow; K4) ----
    from blueice_b200.engine import BinnedEngine, MorphGrid, capture_graph
    t0 = time.perf_counter()
    axes, edges, mus3, pmf, n_model, observed = wl.c3_arrays((200, 200, 20), 4, 3, (-1., 0., 1.), seed=3)
    beng = BinnedEngine(MorphGrid(axes), mus3.reshape(27, 4), pmf, n_model, 0)
    beng.set_observed(observed)
    build_s = time.perf_counter() - t0
    del pmf, n_model
    P3 = 256
    zs3, mult3 = wl.scan_points(P3, 3, 4, seed=31, z_range=(-1., 1.), mult_range=(0.8, 1.2))
    r3 = beng.evaluate(zs3, mult3)
    B, C3, S3 = beng.n_bins, beng.grid.n_corners, beng.n_sources
    hbm = measured_peaks()[0]["hbm_gbs"]
    k4 = {}
    for P in (1, P3):
        zs_d, mult_d, _, _, _ = beng._upload_points(zs3[:P], mult3[:P], None, None)
        zs_d, mult_d = zs_d.clone(), mult_d.clone()
        # device time: the launch sequence of one evaluation (K1, schedule, pass A, total, pass B, total) replayed as a CUDA
        # graph -- what BinnedEngine.evaluate does from its third call on -- between two events (no host work inside)
        for _ in range(2):
            beng.run_device(P, zs_d, mult_d, None, None)
        torch.cuda.synchronize()
        g3 = torch.cuda.CUDAGraph()
        with capture_graph(torch, g3):
            beng.run_device(P, zs_d, mult_d, None, None)
        dm = []
        for k in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g3.replay()
            b.record()
            torch.cuda.synchronize()
            if k > 1:
                dm.append(a.elapsed_time(b))
        del g3
        for _ in range(4):                                         # the third call captures the evaluation's CUDA graph
            beng.evaluate(zs3[:P], mult3[:P])
        ts = []
        for _ in range(10 if P == 1 else 3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            beng.evaluate(zs3[:P], mult3[:P])
            ts.append(time.perf_counter() - t0)
        k4[P] = (float(np.mean(dm)) * 1e-3, float(np.mean(ts)))
    # the Beeston-Barlow pass A alone at P = 1 (the HBM-bound kernel: it streams exactly SURVEY.md 8d's per-point bytes,
    # 8 B (C S + C + 1): pmf corners, calibration-count corners, observed), timed by CUDA events around a launch sequence
    # that is identical except for that kernel being skipped is not possible through the C-ABI, so the kernel's own
    # duration comes from the committed ncu launch list (profiles/r2_k4_ncu.md); here: whole evaluations
    bytes_pt = 8.0 * B * (C3 * S3 + C3 + 1)
    flop_pt = float(B) * (2 * C3 * S3 + 60)
    t1, t256 = k4[1][0], k4[P3][0]
    out["config3_binned_bb"] = {
        "workload": "BinnedLogLikelihood + Beeston-Barlow: 200x200x20 bins, 4 sources, 3 shape parameters x 3 anchors "
                    "(27 anchors); one point and a %d-point scan" % P3,
        "kernel": "k_binned_tile<1> + k_binned_tile<2> (anchor rows of a 256-bin tile staged in shared memory by TMA bulk "
                  "copies once per group of <= 32 points of one hypercube cell; pass A keeps t_b, pass B adds the Poisson terms)",
        "points": P3, "bins": int(B), "finite_results": int(np.isfinite(r3).sum()),
        "single_point": {"device_ms": t1 * 1e3, "e2e_ms": k4[1][1] * 1e3,
                         "roofline": {"bound": "hbm", "bytes_alg": bytes_pt, "achieved": bytes_pt / t1 / 1e9, "peak": hbm,
                                      "unit": "GB/s", "frac": bytes_pt / t1 / 1e9 / hbm,
                                      "note": "whole 5-launch evaluation (schedule, pass A, total, pass B, total) against "
                                              "SURVEY.md 8d's per-point bytes; pass A alone streams those bytes"}},
        "scan": {"device_ms": t256 * 1e3, "e2e_ms": k4[P3][1] * 1e3, "point_bins_sources_per_s": P3 * B * S3 / t256,
                 "roofline": {"bound": "fp64 pipe (shared-data scan: the %d points read one anchor tensor set of %.2f GB through "
                                       "L2 / shared memory, HBM does not bind)" % (P3, 8.0 * B * 27 * (S3 + 1) / 1e9),
                              "flops_alg": flop_pt * P3, "achieved_tflops": flop_pt * P3 / t256 / 1e12,
                              "note": "SURVEY.md 8d's flop count (2 C S + ~60 per point-bin); division, sqrt and log expand to "
                                      "~20-45 FP64 instructions each and the reference's separately rounded multiply-adds "
                                      "cannot be fused, so the instruction-level pipe utilisation is in the ncu summary",
                              "dram_bytes_shared": 8.0 * B * 27 * (S3 + 1)}},
        "build_s": build_s}
    del beng
    torch.cuda.empty_cache()
    # ---- config 5 ----
    t0 = time.perf_counter()
    ll, _, names = wl.c2_api(6, 4, wl.ANCHORS_5, BINS, n_events=1000, seed=5,
                             likelihood_config={'unbinned_engine': 'mixture'})
    build_s = time.perf_counter() - t0
    base_mu = float(np.sum(ll.base_model.expected_events()))
    n_target = args.c5_events
    t0 = time.perf_counter()
    td = ll.base_model.simulate_toys(1, livetime_days=n_target / base_mu, seed=50)
    d = td.to_records()
    gen_s = time.perf_counter() - t0
    del td
    t0 = time.perf_counter()
    ll.set_data(d)
    torch.cuda.synchronize()
    set_data_s = time.perf_counter() - t0
    N = len(d)
    lt = n_target / base_mu
    rng = np.random.default_rng(51)
    x0 = np.concatenate([rng.uniform(0.8, 1.2, size=6), rng.uniform(-1.9, 1.9, size=4)])
    fd = np.repeat(x0[None, :], 11, 0)
    for j in range(10):
        fd[j + 1, j] += 1.4901161193847656e-08
    res5 = {}
    eng = ll._engine
    hbm = measured_peaks()[0]["hbm_gbs"]
    scan = np.column_stack([rng.uniform(0.8, 1.2, size=(64, 6)), rng.uniform(-1.9, 1.9, size=(64, 4))])   # 64-point scan
    for P, table in ((1, fd[:1]), (11, fd), (64, scan)):
        for _ in range(4):                                          # warm-up: the third call captures the CUDA graph
            ll.batch(table, names, livetime_days=lt)
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = ll.batch(table, names, livetime_days=lt)
            ts.append(time.perf_counter() - t0)
        # device-resident: K1 + mix + K5b + finalize
        zs, mult = ll._rows_from_params(table, names)
        sched, _ = eng.single_schedule(zs)
        zs_d, mult_d, scale_d, _, _ = eng._upload_points(zs, mult, np.full(P, lt / ll.pdf_base_config['livetime_days']), None)
        dm = []
        for k in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.run_one_call(P, sched, zs_d, mult_d, scale_d, None)  # bi_template_ll_batch: K1 + mix + K5b + finalize
            b.record()
            torch.cuda.synchronize()
            if k:
                dm.append(a.elapsed_time(b))
        o = eng._setup_terms(P, zs_d, mult_d, scale_d, None)       # stage by stage once: leaves the mixture templates
        eng.run_schedule(sched, o)                                 # in the workspace of mixture_kernel_only
        km = []
        for k in range(6):                                         # the streaming kernel alone
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.mixture_kernel_only(sched, o)
            b.record()
            torch.cuda.synchronize()
            if k:
                km.append(a.elapsed_time(b))
        t_dev, t_e2e, t_k = float(np.mean(dm)) * 1e-3, float(np.mean(ts)), float(np.mean(km)) * 1e-3
        bytes_ev = 4 + 8 * eng.n_space                             # prepared event: bin (i32) + fractions
        res5["P%d" % P] = {"device_ms": t_dev * 1e3, "e2e_ms": t_e2e * 1e3, "point_events_per_s_device": P * N / t_dev,
                           "point_events_per_s_e2e": P * N / t_e2e, "finite": bool(np.all(np.isfinite(r))),
                           "roofline": {"kernel": ("k_mixture_partials<1,2>" if sched["group_points"] == 1 else
                                                   "k_mixture_partials_mma<2,%d> (DMMA.8x8x4 over the lookup corners, "
                                                   "%d-point groups)" % (sched["group_points"] // 8, sched["group_points"])),
                                        "bound": "hbm",
                                        "ms": t_k * 1e3, "bytes_alg": sched["n_groups"] * N * bytes_ev,
                                        "achieved": sched["n_groups"] * N * bytes_ev / t_k / 1e9, "peak": hbm,
                                        "unit": "GB/s", "frac": sched["n_groups"] * N * bytes_ev / t_k / 1e9 / hbm,
                                        "share_of_evaluation": t_k / t_dev}}
    out["config5_large_dataset"] = {
        "workload": "large-dataset fit: %d events on this GPU (full config: 1e8 over 8 GPUs = 1.25e7 per GPU), 6 sources, "
                    "4 shape parameters x 5 anchors (625 anchors), 100x100 templates" % N,
        "kernel": "k_template_mix + k_mixture_partials<1,2> / k_mixture_partials_mma<2,MT> (mixture form: one lookup per "
                  "point-event, events sorted by bin, prepared events streamed once per group of <= 16 points; groups on "
                  "the FP64 tensor pipe)",
        "bytes_per_event": 4 + 8 * eng.n_space, "n_events": int(N), **res5,
        "generate_s": gen_s, "set_data_s": set_data_s, "model_build_s": build_s}
    del eng
    torch.cuda.empty_cache()
    out["config2_long_contraction"] = long_contraction_scan(device, args.steps)
    return out


# ------------------------------------------------------------------------------------------------
def run_own_arm(args):
    import torch
    import torch.distributed as dist
    from blueice_b200 import _cabi
    from blueice_b200.engine import MorphGrid, UnbinnedEngine, capture_graph

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    os.chdir(tempfile.mkdtemp(prefix="bi_bench_"))

    sampler = ClockSampler(local_rank)
    sampler.prepare()
    # ---- build the workload through the public API (prepare + set_data are not part of a step) ----
    t0 = time.perf_counter()
    ll, d, names = wl.c2_api(N_SOURCES, N_SHAPE, ANCHORS, BINS, n_events=args.events, seed=1)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ll.set_data(d)
    torch.cuda.synchronize()
    set_data_s = time.perf_counter() - t0
    eng = ll._engine
    eng.force_kernel = args.kernel
    n_events = len(d)
    P = args.points
    # weak scaling: every rank evaluates its own P points of a (world * P)-point scan
    zs_all, mult_all = wl.scan_points(P * world, N_SHAPE, N_SOURCES, seed=2)
    zs, mult = zs_all[rank * P:(rank + 1) * P], mult_all[rank * P:(rank + 1) * P]
    table = np.ascontiguousarray(np.column_stack([mult, zs]))

    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=device)

    def flush_l2():
        flush.fill_(1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -----------------------------------------------------------------------
    plan = eng.plan(zs)
    zs_d, mult_d, _, _, _ = eng._upload_points(zs, mult, None, None)
    zs_d, mult_d = zs_d.clone(), mult_d.clone()
    plan_dev = tuple(None if t is None else t.clone() for t in eng.upload_plan(plan)[:3])
    # N > 1: the P results of every rank are gathered on every rank by P2P stores over NVLink + one signal-pad barrier
    # (blueice_b200.distributed.PeerGather: bi_peer_broadcast into torch symmetric memory; NCCL all_gather if that
    # is unavailable); no collective inside the evaluation itself
    from blueice_b200.distributed import PeerGather, PointShardedLikelihood
    peer_gather = PeerGather(P) if world > 1 else None
    gathered_dev = None

    def device_step():
        nonlocal gathered_dev
        logl = eng.run_device(P, zs_d, mult_d, None, None, plan, plan_dev)
        if world > 1:
            # this rank's rows go to every rank (stores only): point sharding has no data-path collective, the ranks
            # do not wait for each other inside a step; delivery is confirmed by ONE barrier after the timed loop
            gathered_dev = peer_gather.gather(logl) if args.gather_wait else peer_gather.broadcast(logl)
        return logl

    launches0 = eng.launches
    for _ in range(max(args.warmup, 3)):
        flush_l2()
        device_step()
    launches_per_step = (eng.launches - launches0) // max(args.warmup, 3) + (1 if world > 1 else 0)   # + bi_peer_broadcast
    # the four launches of the evaluation replayed as ONE CUDA graph (as the e2e path does): with N processes sharing the
    # host, eager launches let single steps stall for hundreds of microseconds between kernels (0.54 ms steps next to
    # 0.32 ms ones at N = 8), which the device-resident figure should not depend on.  The P2P stores stay a launch of
    # their own (their slot parity is tracked on the host).
    step_graph, logl_static = None, None
    if not args.no_step_graph:
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with capture_graph(torch, g):
                logl_static = eng.run_device(P, zs_d, mult_d, None, None, plan, plan_dev)
            step_graph = g
        except Exception:
            step_graph = None
            torch.cuda.synchronize()

    def timed_step():
        nonlocal gathered_dev
        if step_graph is None:
            return device_step()
        step_graph.replay()
        if world > 1:
            gathered_dev = peer_gather.gather(logl_static) if args.gather_wait else peer_gather.broadcast(logl_static)
        return logl_static

    for _ in range(3):
        flush_l2()
        timed_step()
    barrier()
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for k in range(args.steps):
        flush_l2()
        starts[k].record()
        logl = timed_step()
        ends[k].record()
    if world > 1:
        peer_gather.barrier()
    barrier()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = float(np.sum(step_ms))
    launches_timed = launches_per_step * args.steps
    result_dev = logl.cpu().numpy().copy()
    gathered_host = gathered_dev.cpu().numpy().copy() if world > 1 else None

    # ---- end-to-end arm (public API, host buffers) ---------------------------------------------------
    # N > 1: PointShardedLikelihood.batch over the whole (world * P)-point scan -- every rank evaluates its P points,
    # all ranks end up with all results (gathered on the device before the D2H)
    if world > 1:
        sharded = PointShardedLikelihood(ll)
        table_all = np.ascontiguousarray(np.column_stack([mult_all, zs_all]))

        def e2e_call():
            return sharded.batch(table_all, names)
    else:
        def e2e_call():
            return ll.batch(table, names)
    # (a sharded call rotates over pinned landing buffers, each with its own CUDA graph, captured on its second use: the
    # warm-up holds its results like the timed loop does, so that both buffers of the rotation are captured before it)
    res = None
    for _ in range(6 if world > 1 else 3):
        flush_l2()
        torch.cuda.synchronize()
        res = e2e_call()
    e2e_s = []
    for k in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        _tight_barrier(world)
        t0 = time.perf_counter()
        res = e2e_call()
        e2e_s.append(time.perf_counter() - t0)
    if world > 1:
        assert np.array_equal(res.reshape(world, P), gathered_host), "sharded e2e result differs from the device gather"
        res = res[rank * P:(rank + 1) * P]
    assert np.array_equal(res, result_dev), "device-resident and e2e arms disagree"
    e2e_total = float(np.sum(e2e_s))
    h2d, d2h = int(eng.last_h2d_bytes), int(eng.last_d2h_bytes)

    # ---- strong scaling of the same scan: P points IN TOTAL, sharded over the ranks (N = 1: the e2e arm itself) ------
    strong = None
    if world > 1:
        table_strong = table_all[:P]
        for _ in range(6):
            flush_l2()
            res_s = sharded.batch(table_strong, names)
        ts, ds = [], []
        for k in range(args.steps):
            flush_l2()
            torch.cuda.synchronize()
            _tight_barrier(world)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            res_s = sharded.batch(table_strong, names)
            b.record()
            ts.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
            ds.append(a.elapsed_time(b) * 1e-3)
        # rank 0 evaluated rows [0, P / world) of the weak scan too: the sharded result must reproduce them bit for bit
        if rank == 0:
            assert np.array_equal(res_s[:P // world], result_dev[:P // world]), "strong-scaled scan differs"
        t = torch.tensor([float(np.sum(ts)), float(np.sum(ds))], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong = {"points_total": P, "e2e_ms": float(t[0]) * 1e3 / args.steps, "device_span_ms": float(t[1]) * 1e3 / args.steps,
                  "point_events_per_s_e2e": P * n_events * args.steps / float(t[0])}

    # ---- dominant kernel alone (K2), for the roofline --------------------------------------
    S, C = eng.n_sources, eng.grid.n_corners
    stream = eng._stream()

    def mma_stage_fns(e, n_pts, z_dev, m_dev):
        """(plan, k2) launchers of the fused path's middle stages, on the workspace a fused call has filled."""
        outs = e.run_fused(n_pts, z_dev, m_dev, None, None)            # K1 outputs + status now live in the workspace
        _, v = e.mma_workspace(n_pts)

        def plan_fn():
            e.mma_plan(n_pts, v, outs["status"])

        def k2_fn():
            e.mma_k2(v)
        return plan_fn, k2_fn, v

    def stream_k2(e, pl, pl_dev, setup, part):
        if len(pl.stream_points):
            _cabi.check(e.lib.bi_unbinned_partials_stream(
                _cabi.dev_ptr(e.ps_anchor), e.ld, e.n_events, S, C, _cabi.dev_ptr(pl_dev[0]),
                len(pl.stream_points), _cabi.dev_ptr(setup["corner"]), _cabi.dev_ptr(setup["weight"]),
                _cabi.dev_ptr(setup["mus"]), _cabi.dev_ptr(setup["status"]), e.outlier_likelihood,
                _cabi.dev_ptr(part), e._stream()), "bi_unbinned_partials_stream")

    if plan.kernel == 'mma':
        plan_fn, k2_only, views = mma_stage_fns(eng, P, zs_d, mult_d)
    else:
        o = eng._setup(P, zs_d, mult_d, None, None)
        partial = eng.ws.get("partial", P * eng.n_super, torch.float64)

        def plan_fn():
            pass

        def k2_only():
            stream_k2(eng, plan, plan_dev, o, partial)

    k2_ms = []
    n_in_kernel = len(plan.stream_points) if plan.kernel != 'mma' else P
    sched = None
    if n_in_kernel:
        for _ in range(3):
            plan_fn()
            k2_only()
        if plan.kernel == 'mma':
            torch.cuda.synchronize()
            hdr = views["header"][:8].cpu().numpy()
            sched = {"point_groups": int(hdr[0]), "superblock_ranges": int(hdr[1]), "superblocks_per_range": int(hdr[2]),
                     "work_units": int(hdr[3]), "evaluable_points": int(hdr[5])}
            n_in_kernel = int(hdr[5])
        for _ in range(max(args.steps, 10)):
            flush_l2()
            plan_fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            k2_only()
            b.record()
            torch.cuda.synchronize()
            k2_ms.append(a.elapsed_time(b))
    clocks = sampler.stop()

    # ---- roofline denominators measured in this run ---------------------------------------------------
    peaks, peak_src = measured_peaks()
    sink = torch.zeros(8, dtype=torch.float64, device=device)
    ms = np.zeros(1, dtype=np.float32)
    flops = np.zeros(1, dtype=np.float64)
    fp64_tflops = []
    for _ in range(3):
        _cabi.check(eng.lib.bi_bench_fp64_fma(1 << 17, 148 * 16, _cabi.dev_ptr(sink), _cabi.host_ptr(ms),
                                              _cabi.host_ptr(flops), stream), "bi_bench_fp64_fma")
        fp64_tflops.append(flops[0] / (ms[0] * 1e-3) / 1e12)
    for _ in range(3):
        _cabi.check(eng.lib.bi_bench_fp64_mma(1 << 15, 148 * 4, _cabi.dev_ptr(sink), _cabi.host_ptr(ms),
                                              _cabi.host_ptr(flops), stream), "bi_bench_fp64_mma")
        fp64_tflops.append(flops[0] / (ms[0] * 1e-3) / 1e12)
    fp64_peak = max(fp64_tflops)

    # streaming regime: P = 1 over a tensor much larger than L2 (HBM-bound), same G/S/C
    stream_info = None
    if rank == 0 and not args.skip_stream:
        n_big = args.stream_events
        big = UnbinnedEngine(MorphGrid(eng.grid.axes), eng.mus_anchor_host, device=device)
        big.allocate_ps_anchor(n_big)
        reps = -(-n_big // n_events)
        src = eng.ps_anchor[:, :, :n_events]
        for r in range(reps):
            lo = r * n_events
            hi = min(lo + n_events, n_big)
            big.ps_anchor[:, :, lo:hi].copy_(src[:, :, :hi - lo])
        big.force_kernel = args.stream_kernel
        z1, m1 = zs[:1], mult[:1]
        plan1 = big.plan(z1)
        z1_d, m1_d, _, _, _ = big._upload_points(z1, m1, None, None)
        z1_d, m1_d = z1_d.clone(), m1_d.clone()
        if plan1.kernel == 'mma':
            plan1_fn, k2_1, _ = mma_stage_fns(big, 1, z1_d, m1_d)
        else:
            plan1_dev = big.upload_plan(plan1)[:3]
            o1 = big._setup(1, z1_d, m1_d, None, None)
            part1 = big.ws.get("partial", big.n_super, torch.float64)

            def plan1_fn():
                pass

            def k2_1():
                stream_k2(big, plan1, plan1_dev, o1, part1)

        def stream_only():
            k2_1()

        for _ in range(3):
            plan1_fn()
            stream_only()
        sms = []
        for _ in range(10):
            plan1_fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            stream_only()
            b.record()
            torch.cuda.synchronize()
            sms.append(a.elapsed_time(b))
        bytes_alg = 8.0 * C * S * n_big
        gbs = bytes_alg / (np.mean(sms) * 1e-3) / 1e9
        # plain streaming read of the same number of bytes with this library's own reader
        n_read = int(min(bytes_alg // 8, big.ps_anchor.numel()))
        rd = []
        for _ in range(5):
            _cabi.check(big.lib.bi_bench_stream_read(_cabi.dev_ptr(big.ps_anchor), n_read, _cabi.dev_ptr(sink),
                                                     _cabi.host_ptr(ms), big._stream()), "bi_bench_stream_read")
            rd.append(n_read * 8 / (ms[0] * 1e-3) / 1e9)
        stream_info = {"kernel": "k_unbinned_mma (P=1: one 8-point m-tile per warp, TMA ring)" if plan1.kernel == 'mma'
                       else "k_unbinned_stream (P=1, lanes=events)", "bound": "hbm",
                       "n_events": int(n_big), "bytes_per_point_event": 8 * C * S,
                       "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                       "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if peak_src == "measured"
                       else "fallback 6.65 TB/s", "ms": float(np.mean(sms)),
                       "point_events_per_s": n_big / (np.mean(sms) * 1e-3),
                       "plain_read_gbs_this_run": float(max(rd))}
        del big

    # ---- the other BASELINE configs on the template-space engine (informational; the headline stays config 2) ----
    other = None
    if not args.skip_other:
        del flush
        torch.cuda.empty_cache()
        other = other_configs(args, rank, world, device)
        if other and other.get("config2_long_contraction"):
            rl = other["config2_long_contraction"]["roofline"]
            rl["peak"] = fp64_peak
            rl["frac"] = rl["achieved"] / fp64_peak

    # ---- reduce over ranks -------------------------------------------------------------------------
    total_ms_max, e2e_total_max = total_ms, e2e_total
    if world > 1:
        t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms_max, e2e_total_max = float(t[0]), float(t[1])

    if rank == 0:
        units = float(world) * P * n_events * args.steps
        value = units / (total_ms_max * 1e-3)
        e2e_value = units / e2e_total_max
        G = eng.grid.n_anchors
        roofline = None
        if k2_ms:
            k2 = float(np.mean(k2_ms)) * 1e-3
            flops_alg = float(n_in_kernel) * n_events * 2.0 * C * S
            bytes_alg = 8.0 * S * n_events * G
            kname = ("k_unbinned_mma<K4=%d> (DMMA.8x8x4 contraction over K=C*S=%d, per-warp TMA ring)"
                     % ((C * S + 3) // 4, C * S) if plan.kernel == 'mma'
                     else "k_unbinned_stream<%d> (one pass over the anchor tensor per point)" % C)
            roofline = {"kernel": kname, "bound": "tensor", "pipe": "fp64 (DMMA.8x8x4 = the FP64 tensor instruction; it shares one pipe with DFMA)", "achieved": flops_alg / k2 / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": flops_alg / k2 / 1e12 / fp64_peak,
                        "traffic": ncu_traffic("scan") if plan.kernel == 'mma' else None,
                        "peak_source": "max(bi_bench_fp64_fma, bi_bench_fp64_mma) measured in this run: DFMA and DMMA "
                                       "share one pipe on sm_100a (MEASURED_PEAKS.json holds no FP64 figure)",
                        "flops_alg_per_point_event": 2 * C * S,
                        "flops_alg_note": "SURVEY.md 8d: 2*C*S flop + 1 log per point-event; the log is replaced by "
                                          "one FP64 multiply per point-event (product tree, one log per 512 events) "
                                          "which is NOT counted, so frac <= 2CS/(2CS+2) = %.3f" % (C * S / (C * S + 1.0)), "ms": k2 * 1e3,
                        "share_of_step": k2 * 1e3 / (total_ms / args.steps),
                        "points_in_kernel": int(n_in_kernel),
                        "hbm": {"bytes_alg": bytes_alg, "achieved_gbs": bytes_alg / k2 / 1e9,
                                "peak_gbs": peaks["hbm_gbs"], "frac": bytes_alg / k2 / 1e9 / peaks["hbm_gbs"],
                                "note": "shared-dataset scan: the anchor tensor is read once per launch, "
                                        "so this kernel is FP64-pipe bound, not HBM bound (SURVEY.md 8d)"}}
        cpu = cpu_baseline_single_core(args.events) if not args.skip_cpu else None
        e2e_obj = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": 1e3 * e2e_total_max / args.steps}
        # flat scalar copies of the multi-GPU splits (the driver's record keeps scalars of the contract objects only)
        e2e_obj.update(sharded_summary(world, P, n_events, e2e_obj, strong, other))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(n_events, P),
                "e2e": e2e_obj, "config2_strong_scaling": strong,
                "gpu_launches": int(launches_timed),
                "launches_per_step": launches_timed / args.steps,
                "roofline": roofline, "roofline_stream": stream_info, "cpu_baseline": cpu,
                "clocks": clocks, "fp64_fma_peak_tflops": fp64_peak,
                "set_data_s": set_data_s, "model_build_s": build_s,
                "plan": sched if sched is not None else
                {"stream_points": int(len(plan.stream_points))},
                "step_ms_min_max": [float(np.min(step_ms)), float(np.max(step_ms))],
                "gather_transport": None if world == 1 else ("nvlink p2p stores (symmetric memory)" if peer_gather.fallback
                                                              is None else "nccl all_gather (%s)" % peer_gather.fallback),
                "other_configs": other}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--events", type=int, default=None, help="fix the number of events (default: Poisson ~100k)")
    ap.add_argument("--stream-events", type=int, default=8 * 1024 * 1024)
    ap.add_argument("--skip-stream", action="store_true")
    ap.add_argument("--stream-kernel", default=None, choices=[None, "mma", "stream"],
                    help="kernel for the P=1 HBM-bound measurement (default: the engine's own choice)")
    ap.add_argument("--kernel", default=None, choices=[None, "mma", "stream"],
                    help="force a K2 kernel for the scan (default: the engine's own choice)")
    ap.add_argument("--gather-wait", action="store_true",
                    help="N > 1: wait for all ranks' results inside every step (a barrier per step) instead of once at the end")
    ap.add_argument("--no-step-graph", action="store_true", help="device arm: eager launches instead of a CUDA graph replay")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-other", action="store_true", help="skip the config-4 / config-5 template-engine runs")
    ap.add_argument("--toys", type=int, default=100000, help="config 4: toys per GPU")
    ap.add_argument("--fit-toys", type=int, default=5000, help="config 4: toys fitted in lock step (bestfit_toys)")
    ap.add_argument("--c5-events", type=int, default=100000000, help="config 5: events on this GPU")
    ap.add_argument("--ref-points-per-core", type=int, default=8)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()

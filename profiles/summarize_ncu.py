"""Condense an `ncu --page raw --csv` export into the handful of counters DESIGN.md / bench.py cite.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/raw.csv
    python profiles/summarize_ncu.py gpurun_out/raw.csv > profiles/<round>_<name>.md
"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe (inst) % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "TMA pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join(label for _, label in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        cells = []
        for key, _ in KEYS:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append("%s %s" % (v, units[i]) if units[i] not in ("", "%") else v)
            else:
                cells.append("n/a")
        kname = r[name_col].split("(")[0].replace("void ", "")
        print("| `%s` | " % kname + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])

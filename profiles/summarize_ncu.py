#!/usr/bin/env python
"""Summarise .ncu-rep captures (gpurun_out/prof_*.ncu-rep) into a markdown table:
    python profiles/summarize_ncu.py gpurun_out/prof_k_push_rows.ncu-rep ... > profiles/rX_ncu_summary.md
Needs only `ncu -i` (no GPU)."""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "cycles / issued inst (per warp)"),
    # warp-state stall reasons: average number of warps per scheduler in that state per issue-active cycle
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    print("| kernel | " + " | ".join(n for _, n in METRICS) + " |")
    print("|---|" + "---|" * len(METRICS))
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            cells = []
            for m, _ in METRICS:
                if m in idx:
                    v, u = r[idx[m]], units[idx[m]]
                    try:
                        v = "%.4g" % float(v.replace(",", ""))
                    except ValueError:
                        pass
                    cells.append((v + " " + u).strip())
                else:
                    cells.append("n/a")
            print("| `%s` | %s |" % (r[idx["Kernel Name"]][:60], " | ".join(cells)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (from gpurun_out/) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\nkernel: {d.get('Kernel Name')}  grid={d.get('Grid Size')} block={d.get('Block Size')}\n")
            for h, u, v in zip(hdr, units, r):
                if h in KEYS or any(h.startswith(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")):
                    f.write(f"  {h:85s} {v} {u}\n")
    print(open(out).read())


if __name__ == "__main__":
    main()

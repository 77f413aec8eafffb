#!/usr/bin/env python3
"""Compact per-kernel summary of an .ncu-rep (`ncu --set full`): duration, DRAM traffic, pipe utilisation,
issue-slot use, occupancy, instruction counts.  Usage: ncu_kernel_summary.py report.ncu-rep > summary.csv"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
] + ["smsp__pcsamp_warps_issue_stalled_" + r for r in (
    "barrier", "branch_resolving", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "no_instructions", "not_selected",
    "selected", "wait", "dispatch_stall", "mio_throttle", "lg_throttle")]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    keep = [i for i, c in enumerate(hdr) if c in ("ID", "Kernel Name") or c in METRICS]
    w = csv.writer(sys.stdout)
    for r in rows:
        w.writerow([(r[i] if i < len(r) else "")[:90] for i in keep])


if __name__ == "__main__":
    main()

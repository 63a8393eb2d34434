#!/usr/bin/env python
"""Condense an `ncu --set full` report into the small metric,value,unit CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index] > profiles/rNN_<what>_ncu_full_summary.csv
    python tools/ncu_summary.py --stalls gpurun_out/prof.ncu-rep [launch_index]     # top source lines by stall samples

Runs here (no GPU needed): it only reads the report with `ncu -i ... --page raw --csv`.
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "launch__waves_per_multiprocessor",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[hdr], rows[hdr + 1], rows[hdr + 2:]


def summary(rep, which=0):
    names, units, data = raw_rows(rep)
    row = data[which]
    col = {n: i for i, n in enumerate(names)}
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "value", "unit"])
    w.writerow(["Kernel Name", row[col["Kernel Name"]], ""])
    w.writerow(["Block Size", row[col["Block Size"]], ""])
    w.writerow(["Grid Size", row[col["Grid Size"]], ""])
    for k in KEEP:
        if k in col:
            w.writerow([k, row[col[k]], units[col[k]]])


def stalls(rep, which=0, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    if not out.strip():
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
    sys.stdout.write(out)


if __name__ == "__main__":
    a = sys.argv[1:]
    if a and a[0] == "--stalls":
        stalls(a[1], int(a[2]) if len(a) > 2 else 0)
    else:
        summary(a[0], int(a[1]) if len(a) > 1 else 0)

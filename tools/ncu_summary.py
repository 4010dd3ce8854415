#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list (per-kernel time share) and a full capture.

    python tools/ncu_summary.py launches gpurun_out/<tag>/launches.csv > profiles/<name>_launches.txt
    python tools/ncu_summary.py full gpurun_out/<tag>/prof.ncu-rep > profiles/<name>_full.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_elapsed.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
]


def launches(path):
    rows = []
    with open(path, newline="") as handle:
        text = "".join(line for line in handle if line.startswith('"'))
    for row in csv.DictReader(io.StringIO(text)):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((row["Kernel Name"].split("(")[0], float(row["Metric Value"].replace(",", ""))))
    per = OrderedDict()
    for name, ns in rows:
        n, t = per.get(name, (0, 0.0))
        per[name] = (n + 1, t + ns)
    total = sum(t for _, t in per.values())
    print("# ncu --metrics gpu__time_duration.sum launch list: %d launches, %.3f ms of kernel time" % (len(rows), total / 1e6))
    print("# (per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes)")
    print("%-64s %8s %12s %12s %7s" % ("kernel", "launches", "total_ms", "mean_us", "share"))
    for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print("%-64s %8d %12.3f %12.1f %6.1f%%" % (name[:64], n, t / 1e6, t / n / 1e3, 100 * t / total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    print("# ncu --set full capture: %s" % path)
    for r in rows[2:]:
        print("== %s  grid=%s block=%s" % (r[header.index("Kernel Name")], r[header.index("Grid Size")],
                                          r[header.index("Block Size")]))
        for m in METRICS:
            if m in header:
                i = header.index(m)
                print("  %-72s %s %s" % (m, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

#!/usr/bin/env python
"""
Summarise ncu captures into small text files that can be committed under profiles/.

    python tools/ncu_summary.py report  gpurun_out/prof.ncu-rep   > profiles/<name>.txt
    python tools/ncu_summary.py launches gpurun_out/launches.csv [skip] > profiles/<name>.txt

`report`   one block per profiled launch: duration, registers, occupancy limits, pipe utilisation,
           issue activity, DRAM bytes (from `ncu --set full`).
`launches` per-kernel totals of a `--metrics gpu__time_duration.sum` launch list (optionally skipping the
           first `skip` launches = warm-up) with each kernel's share of the total.
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none ; source: {}".format(path))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("{}  block={} grid={}".format(d["Kernel Name"], d["Block Size"], d["Grid Size"]))
        for k in KEYS:
            if k in d:
                print("    {:78s} {:>18s} {}".format(k, d[k], units[hdr.index(k)]))


def launches(path, skip=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = OrderedDict()
    for r in rows[1 + skip:]:
        name = r[ik]
        ns = float(r[iv].replace(",", ""))
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in tot.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none ; source: {} ; first {} launches skipped".format(path, skip))
    print("# per-launch times are cold-cache and serialised: compare SHARES")
    print("{:>8s} {:>14s} {:>8s}  kernel".format("launches", "total_us", "share"))
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("{:8d} {:14.1f} {:7.2f}%  {}".format(n, ns / 1e3, 100.0 * ns / total, name[:110]))


if __name__ == "__main__":
    if sys.argv[1] == "report":
        report(sys.argv[2])
    else:
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)

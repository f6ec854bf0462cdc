#!/usr/bin/env python
"""Summaries for profiles/: (1) aggregate an ncu gpu__time_duration launch list by kernel, (2) key metrics of a
`--set full` capture (ncu -i X.ncu-rep --page raw --csv piped in)."""
import csv
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hdr]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0][:70]
        e = agg.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} {v[0]:8d} {v[1] / 1e6:10.2f} {100 * v[1] / tot:6.1f}%")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64_op_dmma.sum", "sm__pipe_fp64_op_dmma_cycles_active.avg.pct_of_peak_sustained_active"]


def raw():
    """stdin: `ncu -i X.ncu-rep --page raw --csv`; argv[2] (optional): substring of the kernel name, argv[3]: which match"""
    rows = list(csv.reader(sys.stdin))
    hdr, unit = rows[0], rows[1]
    body = rows[2:]
    if len(sys.argv) > 2 and "Kernel Name" in hdr:
        kn = hdr.index("Kernel Name")
        body = [r for r in body if len(r) > kn and sys.argv[2] in r[kn]]
    vals = body[int(sys.argv[3]) if len(sys.argv) > 3 else -1]
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for i, h in enumerate(hdr):
        if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            print(f"{h:90s} {vals[i]:>18s} {unit[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw()

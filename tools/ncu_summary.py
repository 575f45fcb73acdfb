#!/usr/bin/env python
"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv   > profiles/launches_r01.txt
    python tools/ncu_summary.py raw gpurun_out/tome_merge_r01.ncu-rep  > profiles/tome_merge_r01.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        name = row["Kernel Name"]
        agg[name][0] += 1
        agg[name][1] += v
        order.append((name, v))
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(order)} launches, {tot:.1f} us total (ncu per-launch times are cold-cache and serialised: compare SHARES)")
    mine = sum(v[1] for k, v in agg.items() if "tokred" in k)
    print(f"# tokred kernels: {mine:.1f} us = {100 * mine / tot:.2f}% of the captured launches")
    print(f"{'us':>10} {'share':>7} {'count':>6}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        tag = " <== tokred" if "tokred" in k else ""
        print(f"{v[1]:10.1f} {100 * v[1] / tot:6.2f}% {v[0]:6d}  {k[:110]}{tag}")
    print("\n# tokred launches in order")
    for name, v in order:
        if "tokred" in name:
            print(f"{v:10.1f} us  {name[:120]}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        print(f"## {rec.get('Kernel Name', '')[:100]}  grid={rec.get('Grid Size')} block={rec.get('Block Size')}")
        for k in hdr:
            if any(k == key or k.startswith(key) for key in KEYS):
                print(f"{k:80s} {rec[k]:>18s} {units[hdr.index(k)]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])

#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv   > profiles/rNN_launches_X.txt
  python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep   > profiles/rNN_kernel_X.txt
"""
import collections
import csv
import re
import subprocess
import sys

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
]


def short_name(name: str) -> str:
    s = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    m = re.search(r"native::([A-Za-z_0-9]+)", s)
    if s.startswith("at::") and m:
        return "at::" + m.group(1)
    m = re.search(r"([A-Za-z_0-9:]+)\s*(<|\()", s)
    return m.group(1) if m else s[:60]


def launches(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row["Metric Unit"]
        us = v / 1000 if u.startswith("n") else (v if u.startswith("u") else v * 1000)
        k = short_name(row["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)")
    print(f"# total {tot:.0f} us over {sum(v[0] for v in agg.values())} launches")
    print(f"{'us':>11} {'share':>6} {'n':>5} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:11.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d} {v[1] / v[0]:9.1f}  {k}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        print("kernel:", row[hdr.index("Kernel Name")][:160])
        for m in RAW_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:75s} {row[i]:>16s} {units[i]}")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        print()


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])

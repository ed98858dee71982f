"""Summaries of ncu outputs for profiles/.
  launches:  python profiles/ncu_summarise.py launches gpurun_out/X_launches.csv
  raw:       ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_summarise.py raw
"""
import collections
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes.sum", "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor"]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ix = {n: i for i, n in enumerate(hdr)}
    by = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        by.setdefault(name, []).append(us)
    ours = sum(sum(v) for k, v in by.items() if "mptv::" in k and "k_int_peak" not in k)
    print(f"# {path}: per-kernel device time (cold-cache, serialised under ncu: compare SHARES)")
    for k, v in by.items():
        share = 100 * sum(v) / ours if ("mptv::" in k and "k_int_peak" not in k) else float("nan")
        print(f"{k[-60:]:60s} launches={len(v):4d} mean_us={sum(v) / len(v):10.1f} total_us={sum(v):10.1f} share_of_hot_path={share:5.1f}%")


def raw():
    rows = list(csv.reader(sys.stdin))
    hdr = rows[0]
    ix = {n: i for i, n in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[ix["Kernel Name"]][:100], "| id", r[ix["ID"]])
        for k in KEYS:
            if k in ix:
                print(f"    {k:70s} {r[ix[k]]:>16s} {rows[1][ix[k]]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw()

"""Turns ncu outputs brought back in gpurun_out/ into the short text summaries kept in profiles/.

  launches: python tools/ncu_summarise.py launches gpurun_out/x.csv "<command>" > profiles/rNN_..._launches_summary.txt
            (csv from: ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file x.csv <command>)
  full:     python tools/ncu_summarise.py full gpurun_out/x.raw.csv "<command>" [kernel-substring] > profiles/rNN_..._ncu.txt
            (csv from: ncu -i x.ncu-rep --page raw --csv > x.raw.csv)
"""
import csv
import sys
from collections import OrderedDict

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "lts__t_bytes.sum", "pcie__read_bytes.sum", "pcie__write_bytes.sum")


def rows_of(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.reader(lines))


def launches(path, command):
    rows = rows_of(path)
    hdr = next(r for r in rows if r and r[0] == "ID")
    acc = OrderedDict()
    for r in rows:
        if len(r) != len(hdr) or r[0] == "ID":
            continue
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        if d.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        acc.setdefault(d["Kernel Name"], []).append(v)
    total = sum(sum(v) for v in acc.values())
    print(f"# {command}")
    print("# kernel, launches, mean ns, share of GPU time (cold-cache, serialised: compare shares)")
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k}, {len(v)}, {sum(v) / len(v):.0f}, {sum(v) / total:.4f}")


def full(path, command, kernel=""):
    rows = rows_of(path)
    hdr = rows[0]
    units = rows[1]
    name_col = hdr.index("Kernel Name")
    out = OrderedDict()
    n = 0
    for r in rows[2:]:
        if len(r) != len(hdr) or kernel not in r[name_col]:
            continue
        n += 1
        for h, u, v in zip(hdr, units, r):
            if h in KEEP:
                out.setdefault((h, u), []).append(v)
    names = sorted({r[name_col] for r in rows[2:] if len(r) == len(hdr) and kernel in r[name_col]})
    print(f"# {command}")
    print(f"# kernels: {'; '.join(names)} ({n} captured launches, values per launch)")
    for (h, u), v in sorted(out.items()):
        print(h, u, v)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")

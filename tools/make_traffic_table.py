"""ncu raw pages (gpurun_out/r02_<name>.raw.csv, from tools/ncu_round2.sh) -> profiles/r02_roofline_traffic.json:
DRAM bytes per launch and per frame of the blend kernel for every benched workload, which bench.py
puts into `roofline.traffic` / `frac_dram`; plus a short text summary per capture in profiles/.

  python tools/make_traffic_table.py
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

wl = graft.load_package().workloads
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6,
        "nsecond": 1.0, "msecond": 1e6}
# capture name -> (table key, frames per launch, bench arguments)
CAPTURES = {
    "cfg3": ("4k_nv12_fullwidth_batch32", 32, "--config 3"),
    "cfg3_distinct": ("4k_nv12_fullwidth_batch32+distinct_cues", 32, "--config 3 --distinct-cues"),
    "cfg5": ("256x1080p_i420_streams", 256, "--config 5"),
    "cfg2": ("1080p_nv12_3_regions", 1, "--config 2"),
    "cfg1": ("720p_i420_single_cue", 1, "--config 1"),
    "cfg4_rgba": ("4k_packed_per_span_colours:RGBA", 8, "--config 4 --format RGBA"),
    "cfg4_ayuv": ("4k_packed_per_span_colours:AYUV", 8, "--config 4 --format AYUV"),
}
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic")


def rows_of(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.reader(lines))


def main():
    table = {}
    for name, (key, frames, args) in CAPTURES.items():
        path = os.path.join(ROOT, "gpurun_out", f"r02_{name}.raw.csv")
        if not os.path.exists(path):
            print("missing", path)
            continue
        rows = rows_of(path)
        hdr, units = rows[0], rows[1]
        kcol = hdr.index("Kernel Name")
        launches = [r for r in rows[2:] if len(r) == len(hdr) and "ttmlblend_group_kernel" in r[kcol]]
        if not launches:
            print("no launches in", path)
            continue

        def col(metric):
            i = hdr.index(metric)
            return [float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0) for r in launches]

        rd, wr, t = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
        per_launch = sum(a + b for a, b in zip(rd, wr)) / len(launches)
        table[key] = {
            "dram_bytes_per_launch": per_launch, "frames_per_launch": frames,
            "dram_bytes_per_frame": per_launch / frames,
            "dram_read_bytes_per_launch": sum(rd) / len(rd), "dram_write_bytes_per_launch": sum(wr) / len(wr),
            "ncu_duration_us": sum(t) / len(t) / 1e3, "launches_captured": len(launches),
            "kernel": launches[0][kcol],
            "source": f"profiles/r02_{name}_ncu.txt (ncu --set full --clock-control none, python bench.py --steps 4 "
                      f"--warmup 3 --no-cpu-baseline --no-e2e --no-extras {args}; dram__bytes_read.sum + "
                      f"dram__bytes_write.sum, mean of {len(launches)} launches)"}
        with open(os.path.join(ROOT, "profiles", f"r02_{name}_ncu.txt"), "w") as out:
            out.write(f"# ncu --set full --clock-control none -k regex:ttmlblend_group_kernel -s 12 -c 2, "
                      f"python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extras {args}\n")
            out.write(f"# kernel: {launches[0][kcol]} ({len(launches)} captured launches, values per launch; "
                      f"under ncu launches are serialised, cold-cache, at ncu's clocks)\n")
            for h, u in zip(hdr, units):
                if h in KEEP:
                    i = hdr.index(h)
                    out.write(f"{h} [{u}] {[r[i] for r in launches]}\n")
            out.write(f"dram bytes per launch (read + write): {per_launch:.0f}  per frame: {per_launch / frames:.0f}\n")
    # BGRA runs the same kernel instantiation over the same layout as RGBA (the alpha byte is byte 3 in both)
    if "4k_packed_per_span_colours:RGBA" in table:
        e = dict(table["4k_packed_per_span_colours:RGBA"])
        e["source"] += "; BGRA: RGBA's capture (same kernel, same layout)"
        table["4k_packed_per_span_colours:BGRA"] = e
    with open(os.path.join(ROOT, "profiles", "r02_roofline_traffic.json"), "w") as f:
        json.dump(table, f, indent=1)
    for k, v in table.items():
        print(f"{k}: {v['dram_bytes_per_launch'] / 1e6:.1f} MB per launch, {v['ncu_duration_us']:.1f} us under ncu")


if __name__ == "__main__":
    main()

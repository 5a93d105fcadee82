#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` output into the markdown table kept under profiles/.

usage: ncu_summary.py raw.csv [launches.csv] > profiles/rNN_....md
Per kernel: duration, DRAM bytes (read+write = the `traffic` of bench.py's roofline object), achieved
DRAM GB/s and its fraction of the measured HBM peak, occupancy, issue utilisation, L1/L2 hit rates."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def to_unit(v, unit, want):
    scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * scale[unit] / scale[want]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    pk = peak()
    print(f"| kernel | grid | regs | time us | dram rd MB | dram wr MB | dram GB/s | frac of {pk:.0f} GB/s | warps active % | issue active % | L1 hit % | L2 hit % | LSU wavefront % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    tot_t = tot_b = 0.0
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("gsc::", "")
        t = to_unit(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]], "us")
        rd = to_unit(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]], "Mbyte")
        wr = to_unit(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]], "Mbyte")
        gbs = (rd + wr) * 1e6 / (t * 1e-6) / 1e9
        tot_t += t
        tot_b += rd + wr

        def g(k):
            return float(r[col[k]]) if k in col and r[col[k]] not in ("", "n/a") else float("nan")
        print(f"| {name} | {r[col['launch__grid_size']]} | {r[col['launch__registers_per_thread']]} | {t:.1f} | {rd:.1f} | {wr:.1f} | "
              f"{gbs:.0f} | {gbs / pk:.3f} | {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {g('l1tex__t_sector_hit_rate.pct'):.0f} | "
              f"{g('lts__t_sector_hit_rate.pct'):.0f} | {g('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed'):.0f} |")
    print(f"| **sum** | | | {tot_t:.1f} | | | {tot_b * 1e6 / (tot_t * 1e-6) / 1e9:.0f} | {tot_b * 1e6 / (tot_t * 1e-6) / 1e9 / pk:.3f} | | | | | |")
    if len(sys.argv) > 2:
        lrows = list(csv.DictReader(l for l in open(sys.argv[2]) if l.startswith('"')))
        agg = {}
        for r in lrows:
            n = r["Kernel Name"].split("(")[0].replace("gsc::", "").replace("<unnamed>::", "")
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += float(r["Metric Value"]) / 1e3
        total = sum(a[1] for a in agg.values())
        print("\nLaunch list (gpu__time_duration.sum, all launches of the command, cold-cache serialised times):\n")
        print("| kernel | launches | total us | share |")
        print("|---|---|---|---|")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| {n} | {a[0]} | {a[1]:.1f} | {a[1] / total:.3f} |")


if __name__ == "__main__":
    main()

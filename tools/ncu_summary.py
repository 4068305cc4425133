"""Summarise an ncu report (raw page CSV) per kernel: launches, mean duration, DRAM bytes and
achieved DRAM GB/s, plus the throughput percentages ncu reports.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_summary.py [hbm_peak_GBs]
"""
import csv
import json
import os
import sys
from collections import OrderedDict

WANT = OrderedDict([
    ("gpu__time_duration.sum", "ns"),
    ("dram__bytes_read.sum", "rd"),
    ("dram__bytes_write.sum", "wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def main():
    peak = float(sys.argv[1]) if len(sys.argv) > 1 else None
    if peak is None:
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6650.0
    rows = list(csv.reader(sys.stdin))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, units, data = r, rows[i + 1], rows[i + 2:]
            break
    if hdr is None:
        sys.exit("no ncu raw CSV header found")
    kn = hdr.index("Kernel Name")
    cols = {k: hdr.index(k) for k in WANT if k in hdr}
    agg = OrderedDict()
    for r in data:
        if len(r) <= kn:
            continue
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, dict(n=0, **{v: 0.0 for v in WANT.values()}))
        a["n"] += 1
        for k, c in cols.items():
            try:
                v = float(r[c].replace(",", ""))
            except ValueError:
                continue
            v *= UNIT_SCALE.get(units[c], 1.0)
            a[WANT[k]] += v
    print("%-22s %4s %10s %12s %12s %9s %7s %6s %6s %6s %5s %6s" %
          ("kernel", "n", "avg us", "dram rd B", "dram wr B", "GB/s", "of pk", "dram%", "l2%", "sm%", "occ%", "regs"))
    for name, a in agg.items():
        n = a["n"]
        us = a["ns"] / n / 1e3
        gbs = (a["rd"] + a["wr"]) / max(a["ns"], 1e-9)
        print("%-22s %4d %10.1f %12.0f %12.0f %9.1f %6.1f%% %6.1f %6.1f %6.1f %5.1f %6.0f" %
              (name[:22], n, us, a["rd"] / n, a["wr"] / n, gbs, 100 * gbs / peak, a["dram%"] / n, a["l2%"] / n, a["sm%"] / n,
               a["occ%"] / n, a["regs"] / n))
    print("(GB/s = DRAM bytes read+written / kernel duration under ncu, cold cache; 'of pk' against %.1f GB/s measured copy bandwidth)" % peak)


if __name__ == "__main__":
    main()

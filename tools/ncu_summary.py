#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into one line per (kernel, grid): duration, DRAM throughput (bytes/s and
% of ncu's peak), SM throughput, occupancy, registers, tensor-pipe activity.   python tools/ncu_summary.py raw.csv"""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, name, scale=None):
        if name not in col or r[col[name]] in ("", "n/a"):
            return float("nan")
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]]
        if scale:
            v *= scale.get(u, 1.0)
        return v

    last = {}
    for r in rows[2:]:
        last[(r[col["Kernel Name"]].split("(")[0][:44], r[col["launch__grid_size"]])] = r
    print("%-46s %7s %10s %10s %7s %6s %6s %5s %7s" % ("kernel", "grid", "dur_us", "dram_GB/s", "dram%", "sm%", "occ%", "regs", "tensor%"))
    for (name, grid), r in last.items():
        dur = num(r, "gpu__time_duration.sum", {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6})
        bw = num(r, "dram__bytes.sum.per_second", {"byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1, "Tbyte/s": 1e3})
        print("%-46s %7s %10.1f %10.1f %7.1f %6.1f %6.1f %5.0f %7.2f" % (
            name, grid, dur, bw, num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), num(r, "launch__registers_per_thread"),
            num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")))


if __name__ == "__main__":
    main(sys.argv[1])

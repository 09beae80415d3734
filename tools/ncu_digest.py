#!/usr/bin/env python
"""Digest of `ncu --set full` captures for profiles/: for every launch in the given .ncu-rep files the numbers the roofline
arithmetic needs (duration, DRAM bytes read + written, tensor-pipe and issue activity, L2 / L1 throughput, top stall reasons).

    python tools/ncu_digest.py gpurun_out/a.ncu-rep [b.ncu-rep ...] [--json profiles/r02_traffic.json --key NAME ...]

Writes a text table to stdout; with --json also merges {key: {"dram_bytes_per_launch": ...}} into that file (bench.py reads
`roofline.traffic` from it instead of carrying a literal)."""
import csv
import io
import json
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "xbar_rd"), ("launch__registers_per_thread", "regs")]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            d[h] = (v, u)
        yield d


def val(d, name):
    if name not in d or d[name][0] in ("", "n/a"):
        return float("nan")
    v, u = d[name]
    return float(v.replace(",", "")) * SCALE.get(u, 1.0)


def main(argv):
    reps = [a for a in argv if a.endswith(".ncu-rep")]
    jpath = argv[argv.index("--json") + 1] if "--json" in argv else None
    keys = argv[argv.index("--key") + 1:] if "--key" in argv else []
    sum_key = None
    if "--sum-key" in argv:          # one entry = the sum over every launch of the capture (e.g. the K12 launches of one update)
        sum_key = argv[argv.index("--sum-key") + 1]
        keys = []
    tot_bytes, tot_us, n_l = 0.0, 0.0, 0
    print("%-52s %9s %10s %10s %8s %7s %6s %6s %6s %10s" % ("kernel (grid)", "dur_us", "dram_rd_MB", "dram_wr_MB", "tensor%",
                                                          "issue%", "l2%", "l1%", "dram%", "xbar_rd_MB"))
    merged = {}
    i = 0
    for rep in reps:
        for d in rows_of(rep):
            name = d["Kernel Name"][0].split("(")[0].replace("void <unnamed>::", "")[:40] + " " + d["launch__grid_size"][0]
            stalls = sorted(((float(v[0]), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k, v in d.items()
                             if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v[0] not in ("", "n/a")
                             and "not_selected" not in k and "_selected" not in k), reverse=True)[:3]
            print("%-52s %9.1f %10.1f %10.1f %8.2f %7.1f %6.1f %6.1f %6.1f %10.1f   stalls: %s" % (
                name, val(d, WANT[0][0]), val(d, WANT[1][0]) / 1e6, val(d, WANT[2][0]) / 1e6, val(d, WANT[3][0]),
                val(d, WANT[4][0]), val(d, WANT[5][0]), val(d, WANT[6][0]), val(d, WANT[7][0]), val(d, WANT[8][0]) / 1e6,
                ", ".join("%s %.1f" % (n, p) for p, n in stalls)))
            tot_bytes += val(d, WANT[1][0]) + val(d, WANT[2][0])
            tot_us += val(d, WANT[0][0])
            n_l += 1
            if i < len(keys):
                merged[keys[i]] = {"dram_bytes_per_launch": val(d, WANT[1][0]) + val(d, WANT[2][0]), "duration_us": val(d, WANT[0][0]),
                                   "source": rep.split("/")[-1]}
            i += 1
    if sum_key:
        merged[sum_key] = {"dram_bytes_per_launch": tot_bytes, "duration_us": tot_us, "launches": n_l,
                           "source": ",".join(r.split("/")[-1] for r in reps)}
        print("sum over %d launches: %.1f MB DRAM, %.1f us" % (n_l, tot_bytes / 1e6, tot_us))
    if jpath and merged:
        try:
            cur = json.load(open(jpath))
        except Exception:
            cur = {}
        cur.update(merged)
        json.dump(cur, open(jpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])

#!/bin/bash
# everything profiles/ wants from ONE GPU box: per-kernel roofline table (all kernels), shard-size diagnostics, the launch list
# of the bench command under ncu, and the K12 update capture
mkdir -p gpurun_out
timeout 900 python tools/kernel_bench.py --reps 5 > gpurun_out/kernels_all.json 2> gpurun_out/kernels_all.err; echo "kernel_bench all rc=$?"; tail -2 gpurun_out/kernels_all.err
bash tools/shard_diag.sh 2>&1 | grep -v "^ \|x  " | head -20
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${XB_NCU_SKIP:-700} -c 3000 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "ncu bench launch list rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/bench_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", "")); us = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    k = r[ki].split("(")[0][:70]; tot[k] += us; cnt[k] += 1
T = sum(tot.values())
print("launches", sum(cnt.values()), "total us", round(T))
for k, v in tot.most_common(14): print("%6.1f%% %9.0f us x%5d %s" % (100 * v / T, v, cnt[k], k))
PY

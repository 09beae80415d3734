#!/bin/bash
# First GPU call of the next round for the experimental K12 tensor-core layers, everything under its
# own timeout so that a hung kernel costs one step, not the box:
#     gpurun --timeout 1500 -- 'bash tools/k12_bringup.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
export XB_EXPERIMENTAL_TC=1
mkdir -p gpurun_out
echo "== parity tests (forward default-on, backward / 3 planes / planes gather / tc PPO update behind the env var)"
timeout 400 python -m pytest tests/test_gpu_tc_conv.py -q > gpurun_out/k12_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/k12_tests.log
echo "== the same tests with the row-coalesced producer mapping (XB_K12_MAP=1: emulator-verified, never run on hardware)"
XB_K12_MAP=1 timeout 400 python -m pytest tests/test_gpu_tc_conv.py -q > gpurun_out/k12_tests_map1.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/k12_tests_map1.log
echo "== per-kernel timings: K12 layers at the PPO minibatch (both producer mappings), K3-P, K10, K11"
for MAPV in 0 1; do
  XB_K12_MAP=$MAPV timeout 400 python tools/kernel_bench.py --only k12,k3p,k10,k11 --reps 10 > gpurun_out/k12_kernels_map$MAPV.json 2> gpurun_out/k12_kernels_map$MAPV.err; echo "map=$MAPV rc=$?"
  python - "$MAPV" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/k12_kernels_map%s.json" % sys.argv[1]))
    for k in d["kernels"]:
        print("%-62s %-28s %9.1f us  %6.1f TF/s  hbm %.2f" % (k["kernel"][:62], k["shape"][:28], k["us"], k.get("TFLOPs", 0.0), k["frac_hbm"]))
except Exception as e:
    print("no kernel json:", e)
PY
done
echo "== headline bench through the tc encoder (3 planes = float32-grade, then 2 planes)"
for P in 3 2; do
  timeout 400 python bench.py --compute tc --tc-planes $P --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_p$P.json 2> gpurun_out/bench_tc_p$P.err
  echo "planes=$P rc=$?"; cat gpurun_out/bench_tc_p$P.json
done
echo "== ncu: the K12 kernels of conv2 / conv3 (report kept in /tmp: a .ncu-rep can exceed the 64 MiB copied back; the raw page is exported here)"
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy \
    --section LaunchStats --clock-control none -k regex:conv_tc_kernel -c 16 -f -o /tmp/k12 \
    python tools/kernel_bench.py --only k12 --reps 1 > /dev/null 2> gpurun_out/k12_ncu.err; echo "ncu rc=$?"
ncu -i /tmp/k12.ncu-rep --page raw --csv > gpurun_out/k12_ncu_raw.csv 2>> gpurun_out/k12_ncu.err
python tools/ncu_summary.py gpurun_out/k12_ncu_raw.csv > gpurun_out/k12_ncu_summary.txt 2>> gpurun_out/k12_ncu.err; cat gpurun_out/k12_ncu_summary.txt

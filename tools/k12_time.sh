#!/bin/bash
# K12: parity tests, per-layer timings (both producer mappings), the headline bench through the tc encoder
mkdir -p gpurun_out
echo "== parity tests"
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -q > gpurun_out/k12_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/k12_tests.log | tail -30
for MAPV in 1 0; do
  XB_K12_MAP=$MAPV timeout 300 python tools/kernel_bench.py --only k12,k3p --reps 5 > gpurun_out/k12_kernels_map$MAPV.json 2> gpurun_out/k12_kernels_map$MAPV.err; echo "map=$MAPV rc=$?"
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
for cfgs in "1 3" "1 2"; do
  set -- $cfgs
  XB_K12_MAP=$1 timeout 300 python bench.py --compute tc --tc-planes $2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_tc_map$1_p$2.json 2> gpurun_out/bench_tc_map$1_p$2.err
  echo "map=$1 planes=$2 rc=$?"; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_tc_map$1_p$2.json').read()); print(d['value'], d['ms_per_step'], d.get('last_info'))
except Exception as e: print('no json', e)"
  tail -3 gpurun_out/bench_tc_map$1_p$2.err
done

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_conv.py -q > gpurun_out/k12_tests.log 2>&1; echo "tests (TMA) rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/k12_tests.log | tail -20
XB_K12_TMA=0 timeout 300 python -m pytest tests/test_gpu_tc_conv.py -q > gpurun_out/k12_tests_notma.log 2>&1; echo "tests (cp.async only) rc=$?"; tail -2 gpurun_out/k12_tests_notma.log
timeout 300 python tools/kernel_bench.py --only k12 --reps 5 > gpurun_out/k12_kernels.json 2> gpurun_out/k12_kernels.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/k12_kernels.json"))
    tot = {}
    for k in d["kernels"]:
        print("%-62s %-28s %9.1f us  %6.1f TF/s" % (k["kernel"][:62], k["shape"][:28], k["us"], k.get("TFLOPs", 0.0)))
        key = "P=3" if ("PB=3" in k["kernel"] or "P=3" in k["kernel"]) else "P=2"
        tot[key] = tot.get(key, 0) + k["us"]
    print("sum per minibatch (us):", tot)
except Exception as e:
    print("no kernel json:", e)
PY
tail -5 gpurun_out/k12_kernels.err

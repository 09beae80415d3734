#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke, the default bench line (+ K12 per-layer timings)
mkdir -p gpurun_out
timeout 120 python tools/tma_probe.py > gpurun_out/tma_probe.log 2>&1; echo "probe rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/gpu_tests.log | tail -25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json; tail -6 gpurun_out/bench_default.err
timeout 300 python tools/kernel_bench.py --only k12,k12box --reps 5 > gpurun_out/k12_kernels.json 2> gpurun_out/k12_kernels.err; echo "kernel_bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/k12_kernels.json"))
    for k in d["kernels"]:
        if "PB=3" in k["kernel"] or "P=3" in k["kernel"]:
            print("%-62s %-28s %9.1f us  %6.1f TF/s" % (k["kernel"][:62], k["shape"][:28], k["us"], k.get("TFLOPs", 0.0)))
except Exception as e:
    print("no kernel json:", e)
PY
tail -3 gpurun_out/k12_kernels.err

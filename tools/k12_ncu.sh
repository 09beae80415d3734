#!/bin/bash
# ncu --set full on the K12 kernels of conv2 forward / fc forward (3 planes) at a reduced batch (ncu replays ~40x)
mkdir -p gpurun_out
export XB_K12_BATCH=${XB_K12_BATCH:-2048}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/k12_conv1 \
    python tools/kernel_bench.py --only k12 --reps 1 > /dev/null 2> gpurun_out/k12_ncu1.err; echo "ncu conv1 fwd rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 11 -c 1 -f -o gpurun_out/k12_conv2 \
    python tools/kernel_bench.py --only k12 --reps 1 > /dev/null 2> gpurun_out/k12_ncu2.err; echo "ncu conv2 fwd rc=$?"
ls -la gpurun_out/*.ncu-rep
timeout 300 python bench.py --compute tc --tc-planes 3 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_tc_p3.json 2> gpurun_out/bench_tc_p3.err; echo "bench rc=$?"; cat gpurun_out/bench_tc_p3.json | head -c 600; tail -3 gpurun_out/bench_tc_p3.err

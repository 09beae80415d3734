#!/bin/bash
# ncu --set full on the 14 K12 launches of ONE PPO update at the bench shape (8192 rows, 3 planes): forward conv1 / conv2 / conv3 /
# Linear, then Linear wgrad + dgrad, conv3 wgrad + dgrad, conv2 wgrad + 4 dgrad phases, conv1 wgrad.  The two warm-up updates
# (28 launches) are skipped.  Output: gpurun_out/k12_update.ncu-rep (+ a launch list with durations only).
mkdir -p gpurun_out
export XB_PROF_PLAIN=1 XB_PROF_T=${XB_PROF_T:-32}
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel --launch-skip 28 --launch-count 14 -f \
    -o gpurun_out/k12_update python tools/update_profile.py > gpurun_out/k12_ncu.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/k12_ncu.log
ls -la gpurun_out/*.ncu-rep
# the launch list of the bench command itself (graph kernel nodes are profiled one by one): per-launch times are serialised
# and cold-cache - only the kernels' SHARES of the step are comparable with the bench line
# (the first ~700 launches are model initialisation and the synthetic rollout; the list keeps the next 3000: warm-up + timed step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${XB_NCU_SKIP:-700} -c 3000 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "ncu bench launch list rc=$?"

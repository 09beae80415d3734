#!/bin/bash
# compute-sanitizer over the kernels that hand-roll mbarrier / bulk-async / tcgen05 / TMEM protocols (K3 gathers, K9-TC,
# K12) and the plain kernels next to them.  Run on a GPU box:   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Each tool x test subset gets its own timeout (the sanitizer slows kernels 10-100x); summaries land in gpurun_out/ and the
# ones to be judged are copied to profiles/.
mkdir -p gpurun_out
export XB_SANITIZE=1          # tests shrink their shapes when this is set
SAN=/usr/local/cuda/bin/compute-sanitizer
declare -A SUBSETS
SUBSETS[k3]="tests/test_gpu_rollout.py -k gather"
SUBSETS[k9tc]="tests/test_gpu_qmix.py -k tensor_core_mixer_forward"
SUBSETS[k12]="tests/test_gpu_tc_conv.py -k 'forward_conv_three_planes or raw_uint8 or data_gradient'"
SUBSETS[k12w]="tests/test_gpu_tc_conv.py -k 'weight_gradient_raw or gather_obs_planes or split_and_pack or pack_weights or padded_rows'"
SUBSETS[k12box]="tests/test_gpu_tc_conv.py -k 'box_convolutions'"
for tool in memcheck racecheck synccheck; do
  keys="k9tc k12 k12box"                       # the mbarrier / TMEM protocols: every tool
  [ $tool = memcheck ] && keys="k3 k9tc k12 k12w k12box"
  for key in $keys; do
    log=gpurun_out/sanitize_${tool}_${key}.log
    eval timeout ${XB_SAN_TIMEOUT:-240} $SAN --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest ${SUBSETS[$key]} -q -x -p no:cacheprovider > $log 2>&1
    rc=$?
    echo "== $tool $key rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $log | tr '\n' ' ')"
  done
done

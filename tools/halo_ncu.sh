#!/bin/bash
# one ncu --set full capture (with SASS-level stall sampling) of the halo-mode conv3 forward launch, at a reduced batch
mkdir -p gpurun_out
export XB_K12_BATCH=${XB_K12_BATCH:-2048}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel --launch-skip ${XB_NCU_SKIP:-30} --launch-count 1 -f \
    -o gpurun_out/k12_halo python tools/kernel_bench.py --only k12box --reps 1 > gpurun_out/halo_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/halo_ncu.log
ncu -i gpurun_out/k12_halo.ncu-rep --page source --csv > gpurun_out/k12_halo_source.csv 2> /dev/null; wc -l gpurun_out/k12_halo_source.csv

#!/bin/bash
# compute-side scaling bound on ONE GPU: one rank's shard of a G-GPU job timed alone (bench.py --shard-of G), and the per-kernel
# table of one update at the 8-GPU shard size (1024 rows)
mkdir -p gpurun_out
for g in 2 4 8; do
  timeout 300 python bench.py --shard-of $g --e2e-steps 0 > gpurun_out/shard_of_$g.json 2> gpurun_out/shard_of_$g.err; echo "shard-of $g rc=$?"
  tail -1 gpurun_out/shard_of_$g.json; grep -E "Error|error|Traceback" -A3 gpurun_out/shard_of_$g.err | head -12
done
XB_PROF_ENVS=32 XB_PROF_T=32 timeout 300 python tools/update_profile.py > gpurun_out/update_profile_1024.txt 2>&1; echo "profile rc=$?"
grep -v Warn gpurun_out/update_profile_1024.txt | head -64

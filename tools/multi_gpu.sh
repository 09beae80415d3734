#!/bin/bash
# usage: tools/multi_gpu.sh N   (run under `gpurun --gpus N`): the bench line at N GPUs exactly as the driver launches it,
# the secondary workloads (BASELINE configs 3 / 5: PER-DQN, SAC, QMIX) at N GPUs, and - for comparison on the same box - the
# N=1 bench line.  Everything lands in gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline ${XB_N1_FLAGS:-} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench N=1 rc=$?"
cut -c1-330 gpurun_out/bench_n1.json
timeout 900 bash -c "$(declare -f run); run $N 29511 bench.py --gpus $N --steps 10 --warmup 3" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
cut -c1-330 gpurun_out/bench_n$N.json; tail -4 gpurun_out/bench_n$N.err
timeout 900 bash -c "$(declare -f run); run $N 29512 tools/algo_bench.py --no-cpu --graph --iters 200" > gpurun_out/algo_n$N.json 2> gpurun_out/algo_n$N.err; echo "algo N=$N rc=$?"
cat gpurun_out/algo_n$N.json | cut -c1-1500; tail -4 gpurun_out/algo_n$N.err
python - <<PY
import json
for f in ("gpurun_out/bench_n1.json", "gpurun_out/bench_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f  ms/step %.2f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"] if d.get("e2e") else -1))
        print("   phases", json.dumps(d.get("phases")))
        print("   parity", json.dumps(d.get("parity_vs_1gpu")))
    except Exception as e:
        print(f, "unreadable:", e)
PY

#!/usr/bin/env python
"""Secondary measurements for BASELINE configs 3-5 (parity-test cases, not the headline): sample + update throughput of
DQN+PER (2^20-transition HBM replay, B=512), SAC (1M replay, B=1024) and QMIX (5 agents x 72-d, T=60, 32 episodes)
on one B200, next to the oracle port of the reference's torch-CPU path on the host.  Prints one JSON object.

    python tools/algo_bench.py [--iters 200] [--cpu-iters 5] [--no-cpu] [--graph]
    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/algo_bench.py --no-cpu --graph    (G GPUs)

Multi-GPU (BASELINE configs 3 and 5 are multi-GPU configurations): one process per GPU; the replay is sharded (envs / episodes
per rank), every update takes batch/G rows from every rank, the gradient bucket (+ logged statistics) is all-reduced once per
update inside the captured graph; the time of an update is the max over ranks (barrier + synchronize on both sides) and the
global batch is fixed (strong scaling).
"""
import argparse
import json
import os
import sys
import time
from argparse import Namespace

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
RANK, WORLD, LOCAL = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
DEV = "cuda:%d" % LOCAL


def timed(fn, iters, warm=10):
    import torch.distributed as dist
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 1e3 / iters], dtype=torch.float64, device=DEV)
    if WORLD > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cpu_timed(fn, iters):
    fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters


def bench_perdqn(args):
    from xuance_b200.common import PerOffPolicyBuffer, Box, Discrete, BaseCallback
    from xuance_b200.torch.rl_models import Basic_CNN, DeepQNetwork
    from xuance_b200.torch.learners import PerDQN_Learner
    N, S, B, A = 16 // WORLD, 65536, 512 // WORLD, 4          # this rank's envs and its share of the 512-row batch
    buf = PerOffPolicyBuffer(Box(0, 255, (84, 84, 4), np.uint8), Discrete(A), None, N, N * S, B, alpha=0.5, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(RANK)
    for t in range(64):   # real stores through K1 + K5 insert, then the rest of the ring is declared filled
        o = torch.randint(0, 256, (N, 84, 84, 4), dtype=torch.uint8, device=DEV, generator=g)
        buf.store(o, torch.randint(0, A, (N,), device=DEV), torch.randn(N, device=DEV), torch.zeros(N, device=DEV), o)
    buf.size = S
    leaves = torch.rand((N, S), device=DEV) + 0.01
    cap = buf._it_capacity
    buf._it_sum[:, cap:cap + S], buf._it_min[:, cap:cap + S] = leaves, leaves
    lvl = cap // 2
    while lvl >= 1:
        buf._it_sum[:, lvl:2 * lvl] = buf._it_sum[:, 2 * lvl:4 * lvl:2] + buf._it_sum[:, 2 * lvl + 1:4 * lvl:2]
        buf._it_min[:, lvl:2 * lvl] = torch.minimum(buf._it_min[:, 2 * lvl:4 * lvl:2], buf._it_min[:, 2 * lvl + 1:4 * lvl:2])
        lvl //= 2
    rep = Basic_CNN(input_shape=(84, 84, 4), kernels=[8, 4, 3], strides=[4, 2, 1], filters=[32, 64, 64], activation=nn.ReLU, device=DEV)
    model = DeepQNetwork(rep, [512], Discrete(A), None, None, nn.ReLU, DEV).to(DEV)
    cfg = Namespace(distributed_training=WORLD > 1, episode_length=1000, use_grad_clip=False, grad_clip_norm=0.5, device=DEV,
                    model_dir="/tmp/x", running_steps=10**7, parallels=N * WORLD, learning_rate=1e-4, gamma=0.99,
                    sync_frequency=500, start_training=0, training_frequency=1, compute=args.compute, tc_planes=3)
    if args.compute != "fp32":
        rep.tc_planes = 3
        rep.set_compute(args.compute)
    cfg.use_cuda_graph = args.graph
    lrn = PerDQN_Learner(cfg, model, BaseCallback())

    def step():
        s = buf.sample(0.4)
        td, _ = lrn.update(sync=False, **s)
        buf.update_priorities(s["step_choices"], td)

    dt = timed(step, args.iters)
    out = {"config": "DQN+PER, 2^20-transition uint8 HBM replay (59 GB), 16 envs, B=512 (global), encoder compute=%s" % args.compute,
           "gpu_updates_per_s": 1 / dt, "gpu_transitions_per_s": B * WORLD / dt, "gpu_ms_per_update": dt * 1e3}
    if not args.no_cpu:
        from oracle.replay import PerReplayOracle
        from oracle.nets import DeepQNetworkOracle
        from oracle.learners import DQNLearnerOracle
        Sc = 1024   # the reference stores float32 image replay: 2^20 would need 226 GB of host RAM
        ob = PerReplayOracle((84, 84, 4), (), N, N * Sc, B, alpha=0.5, obs_dtype=np.uint8)
        rng = np.random.default_rng(0)
        for t in range(Sc):
            o = rng.integers(0, 256, size=(N, 84, 84, 4), dtype=np.uint8)
            ob.store(o, rng.integers(0, A, N), rng.normal(size=N).astype(np.float32), np.zeros(N, bool), o)
        ol = DQNLearnerOracle(DeepQNetworkOracle(A), per=True)

        def cstep():
            s = ob.sample(0.4)
            td, _ = ol.update(**s)
            ob.update_priorities(s["step_choices"], td)

        cdt = cpu_timed(cstep, args.cpu_iters)
        out.update(cpu_updates_per_s=1 / cdt, cpu_ms_per_update=cdt * 1e3, cpu_replay_capacity=N * Sc,
                   cpu_threads=torch.get_num_threads(), speedup=cdt / dt)
    return out


def bench_sac(args):
    from copy import deepcopy
    from xuance_b200.common import DummyOffPolicyBuffer, Box, BaseCallback
    from xuance_b200.torch.rl_models import Basic_Identical, SAC_GaussianActor, TwinActionValueCritic, SoftActorCritic
    from xuance_b200.torch.learners import REGISTRY_Learners
    N, S, B, od, ad = 4, 250000 // WORLD, 1024 // WORLD, 17, 6
    aspace = Box(-1, 1, (ad,), np.float32)
    buf = DummyOffPolicyBuffer(Box(-10, 10, (od,), np.float32), aspace, None, N, N * S, B, device=DEV)
    buf._obs.normal_(), buf._next_obs.normal_(), buf._act_rows.uniform_(-1, 1), buf._fields.normal_()
    buf.size = S
    rep = Basic_Identical((od,), device=DEV)
    model = SoftActorCritic(SAC_GaussianActor(rep, [256, 256], aspace, None, None, nn.LeakyReLU, nn.Tanh, DEV),
                            TwinActionValueCritic(deepcopy(rep), aspace, [256, 256], None, None, nn.LeakyReLU, DEV)).to(DEV)
    cfg = Namespace(distributed_training=WORLD > 1, episode_length=1000, use_grad_clip=False, grad_clip_norm=0.5, device=DEV,
                    model_dir="/tmp/x", running_steps=10**6, parallels=N, start_training=0, training_frequency=1,
                    learning_rate_actor=1e-3, learning_rate_critic=1e-3, tau=0.005, gamma=0.99, alpha=0.2,
                    use_automatic_entropy_tuning=True)
    cfg.use_cuda_graph = args.graph
    lrn = REGISTRY_Learners["SAC_Learner"](cfg, model, BaseCallback())
    dt = timed(lambda: lrn.update(sync=False, **buf.sample()), args.iters)
    out = {"config": "SAC, 17-d obs / 6-d act, 1M replay, B=1024 (global), MLP 256-256", "gpu_updates_per_s": 1 / dt,
           "gpu_transitions_per_s": B * WORLD / dt, "gpu_ms_per_update": dt * 1e3}
    if not args.no_cpu:
        from oracle.replay import UniformReplayOracle
        from oracle.sac import SACModelOracle, SACLearnerOracle
        ob = UniformReplayOracle((od,), (ad,), N, N * S, B)
        ob.observations[...] = np.random.default_rng(0).normal(size=ob.observations.shape)
        ob.size = S
        ol = SACLearnerOracle(SACModelOracle(od, ad))
        cdt = cpu_timed(lambda: ol.update(torch.randn(B, ad), torch.randn(B, ad), **ob.sample()), max(args.cpu_iters, 20))
        out.update(cpu_updates_per_s=1 / cdt, cpu_ms_per_update=cdt * 1e3, cpu_threads=torch.get_num_threads(), speedup=cdt / dt)
    return out


def bench_qmix(args, Be=32):
    from helpers import qmix_episode_stream
    from test_gpu_qmix import _product_model, _buffers
    from xuance_b200.common import BaseCallback
    from xuance_b200.torch.learners import REGISTRY_Learners
    n, od, A, S, T = 5, 72, 12, 98, 60
    Be_global, Be = Be, Be // WORLD                              # this rank's share of the episodes of an update
    keys, grouping, model = _product_model(n, od, A, S, device=DEV)
    n_envs, C = 8, max(256, 2 * Be)
    prod, ob = _buffers(keys, od, A, S, n_envs, C, Be, T, device=DEV)
    for ev in qmix_episode_stream(np.random.default_rng(RANK), keys, n_envs, T, od, A, S, C // n_envs):
        if ev[0] == 'store':
            prod.store(**ev[1]), ob.store(**ev[1])
        else:
            prod.finish_path(ev[1], **ev[2]), ob.finish_path(ev[1], **ev[2])
    cfg = Namespace(distributed_training=WORLD > 1, episode_length=T, use_grad_clip=False, grad_clip_norm=10.0, device=DEV,
                    model_dir="/tmp/x", running_steps=10**7, parallels=n_envs, use_parameter_sharing=True, use_rnn=True,
                    use_actions_mask=False, learning_rate=7e-4, sync_frequency=200, double_q=True, n_epochs=1,
                    start_training=0, gamma=0.99)
    cfg.use_cuda_graph = args.graph
    lrn = REGISTRY_Learners["QMIX_Learner"](cfg, grouping, model, BaseCallback())
    dt = timed(lambda: lrn.update(prod.sample(), sync=False), args.iters)
    out = {"config": f"QMIX, 5 agents x 72-d, S=98, A=12, T=60, {Be_global} episodes/update (global), GRU 64 + mixer 32/32",
           "gpu_updates_per_s": 1 / dt, "gpu_agent_steps_per_s": Be_global * T * n / dt, "gpu_ms_per_update": dt * 1e3}
    if not args.no_cpu:
        from oracle.qmix import QMIXModelOracle, QMIXLearnerOracle
        ol = QMIXLearnerOracle(QMIXModelOracle(n, od, A, S), keys, detach_q_eval=False)
        cdt = cpu_timed(lambda: ol.update(ob.sample()), max(args.cpu_iters, 10))
        out.update(cpu_updates_per_s=1 / cdt, cpu_ms_per_update=cdt * 1e3, cpu_threads=torch.get_num_threads(), speedup=cdt / dt)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--cpu-iters", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--graph", action="store_true", help="capture the learners' device update in a CUDA graph")
    ap.add_argument("--compute", default="fp32", choices=["fp32", "tc"], help="pixel encoder of the PER-DQN network")
    args = ap.parse_args()
    torch.cuda.set_device(LOCAL)
    torch.manual_seed(0)          # identical initial weights on every rank
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(DEV))
        args.no_cpu = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    only = set(x for x in args.only.split(",") if x)
    res = {}
    if not only or "perdqn" in only:
        res["perdqn"] = bench_perdqn(args)
    if not only or "sac" in only:
        res["sac"] = bench_sac(args)
    if not only or "qmix" in only:
        res["qmix"] = bench_qmix(args)
        res["qmix_4096"] = bench_qmix(Namespace(**{**vars(args), "iters": 20, "no_cpu": True}), Be=1024)
    res["cuda_graph"] = bool(args.graph)
    res["n_gpus"] = WORLD
    if RANK == 0:
        print(json.dumps(res))
    if WORLD > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

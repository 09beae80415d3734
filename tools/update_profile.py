"""Per-kernel GPU time of ONE eager PPO epoch (4 updates of 8192 rows) of the bench workload: torch.profiler (CUPTI) table,
all kernels - ours and the torch / cuBLAS ones around them.  Run on a GPU box; not a bench value (profiler attached)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xuance_b200.common import Box, Discrete
from xuance_b200.torch.agents import PPO_Agent


def main():
    dev = "cuda:0"
    torch.cuda.set_device(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n_envs, T = int(os.environ.get("XB_PROF_ENVS", "256")), int(os.environ.get("XB_PROF_T", "32"))
    cfg = bench.ppo_namespace(dev, n_envs, False, "tc")
    cfg.tc_planes = 3
    cfg.parallels = n_envs
    cfg.horizon_size = T
    cfg.buffer_size = n_envs * T
    cfg.n_minibatch = max(1, n_envs * T // 8192)
    cfg.use_cuda_graph = False
    agent = PPO_Agent(cfg, envs=None, observation_space=Box(0, 255, bench.OBS_SHAPE, np.uint8), action_space=Discrete(bench.N_ACTIONS))
    rng = np.random.default_rng(0)
    acts, rews, vals, terms, logp, boot = bench.synth_scalars(rng, T, n_envs)
    mem = agent.memory
    for t in range(T):
        frame = torch.randint(0, 256, (n_envs,) + bench.OBS_SHAPE, dtype=torch.uint8, device=dev)
        mem.store(frame, torch.from_numpy(acts[t]).to(dev), torch.from_numpy(rews[t]).to(dev), torch.from_numpy(vals[t]).to(dev),
                  torch.from_numpy(terms[t]).to(dev), {"old_logp": torch.from_numpy(logp[t]).to(dev)})
    for i in range(n_envs):
        mem.finish_path(float(boot[i]), i)
    for _ in range(2):
        agent.train_epochs(1)
    torch.cuda.synchronize()
    if os.environ.get("XB_PROF_PLAIN", "0") == "1":      # under ncu (CUPTI is taken): just run the one update
        agent.train_epochs(1)
        torch.cuda.synchronize()
        print("plain update done")
        return
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        agent.train_epochs(1)
        torch.cuda.synchronize()
    n_upd = cfg.n_minibatch
    rows = []
    for e in prof.key_averages():
        dt = getattr(e, "device_time_total", None)
        if dt is None:
            dt = getattr(e, "cuda_time_total", 0.0)
        if e.device_type.name == "CUDA" or (dt and e.self_device_time_total > 0):
            rows.append((e.self_device_time_total / n_upd, e.count / n_upd, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print("GPU kernel time per update: %.1f us over %d updates" % (tot, n_upd))
    for us, cnt, key in rows[:70]:
        print("%9.1f us  x%5.1f  %s" % (us, cnt, key[:150]))


if __name__ == "__main__":
    main()

"""Process-group plumbing (reference: xuance/torch/utils/operations.py:11-28 ``init_distributed_mode``).

One process per GPU, launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment);
NCCL on CUDA, gloo when CUDA is absent (CPU tests of the host-side sharding logic).  The data path uses exactly
one collective per update: a sum all-reduce of the learner's flat gradient bucket (plus a 3-float all-reduce for
the global advantage statistics of a sharded PPO minibatch)."""
import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_distributed_mode(master_port=None, backend=None):
    """Idempotent init from the torchrun environment; returns (rank, world_size, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world_size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if master_port is not None:
            os.environ.setdefault("MASTER_PORT", str(master_port))
        use_cuda = torch.cuda.is_available()
        if use_cuda:
            torch.cuda.set_device(local_rank)
        kw = {"device_id": torch.device("cuda", local_rank)} if use_cuda else {}
        dist.init_process_group(backend=backend or ("nccl" if use_cuda else "gloo"), rank=rank,
                                world_size=world_size, **kw)
    return rank, world_size, local_rank


def allreduce_sum_(t):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        from xuance_b200 import _lib
        prof = _lib.profile
        if prof is not None and "nccl_all_reduce" in prof:      # bench.py's per-phase breakdown: CUDA events around the collective
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            e1.record()
            prof["nccl_all_reduce"].append((e0, e1))
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_bounds(n, rank, world_size):
    """Contiguous env shard of rank ``rank``: [lo, hi)."""
    assert n % world_size == 0, "n_envs must be divisible by the world size"
    per = n // world_size
    return rank * per, (rank + 1) * per


def stratified_minibatches(n_local, n_minibatch, rng):
    """Rank-local shuffle split into n_minibatch equal slices (DESIGN.md "Multi-GPU"): global minibatch m is the
    union over ranks of local slice m, so every rank contributes exactly B/world rows to every update."""
    perm = rng.permutation(n_local)
    per = n_local // n_minibatch
    return [perm[i * per:(i + 1) * per] for i in range(n_minibatch)]

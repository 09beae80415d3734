"""GPU: checkpoint of the HBM buffers, ring cursors and PER trees - a restored buffer continues
bit-identically to the one that was saved."""
import random
from argparse import Namespace

import numpy as np
import pytest
import torch

from helpers import synth_rollout, qmix_episode_stream

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _roundtrip(sd, tmp_path, name):
    p = str(tmp_path / name)
    torch.save(sd, p)
    return torch.load(p, map_location="cpu", weights_only=False)


def test_onpolicy_buffer_checkpoint(tmp_path):
    from xuance_b200.common import DummyOnPolicyBuffer_Atari, Box, Discrete
    N, T, shape = 4, 12, (12, 12, 4)
    mk = lambda: DummyOnPolicyBuffer_Atari(Box(0, 255, shape, np.uint8), Discrete(4), {"old_logp": ()}, N, T, device=DEV)
    ro = synth_rollout(np.random.default_rng(0), N, T, shape, p_term=0.1)
    a = mk()
    step = lambda buf, t: buf.store(ro["obs"][t], ro["acts"][t], ro["rews"][t], ro["vals"][t], ro["terms"][t],
                                    {"old_logp": ro["logp"][t]})
    for t in range(7):
        step(a, t)
    a.finish_path(0.25, 1)
    b = mk()
    b.load_state_dict(_roundtrip(a.state_dict(), tmp_path, "onpolicy.buffer"))
    assert (b.ptr, b.size) == (a.ptr, a.size) and np.array_equal(a.start_ids, b.start_ids)
    for buf in (a, b):
        for t in range(7, T):
            step(buf, t)
        for i in range(N):
            buf.finish_path(float(ro["boot"][i]), i)
    idx = np.random.default_rng(1).permutation(N * T)
    sa, sb = a.sample(idx), b.sample(idx)
    for k in ("obs", "actions", "returns", "values", "advantages"):
        assert torch.equal(sa[k], sb[k]), k
    assert torch.equal(sa["aux_batch"]["old_logp"], sb["aux_batch"]["old_logp"])
    with pytest.raises(ValueError):
        DummyOnPolicyBuffer_Atari(Box(0, 255, shape, np.uint8), Discrete(4), {"old_logp": ()}, N, T + 1,
                                  device=DEV).load_state_dict(a.state_dict())


def test_per_buffer_checkpoint(tmp_path):
    from xuance_b200.common import PerOffPolicyBuffer, Box, Discrete
    N, S, B = 4, 32, 16
    mk = lambda: PerOffPolicyBuffer(Box(-9, 9, (5,), np.float32), Discrete(3), None, N, N * S, B, alpha=0.6, device=DEV)
    rng = np.random.default_rng(2)
    a = mk()

    def feed(buf, n, r):
        for _ in range(n):
            buf.store(r.normal(size=(N, 5)).astype(np.float32), r.integers(0, 3, N), r.normal(size=N).astype(np.float32),
                      r.random(N) < 0.1, r.normal(size=(N, 5)).astype(np.float32))

    feed(a, 40, rng)            # wraps the ring
    s = a.sample(0.4, uniforms=rng.random((N, B // N)))
    a.update_priorities(s["step_choices"], np.abs(rng.normal(size=B)).astype(np.float32))
    b = mk()
    b.load_state_dict(_roundtrip(a.state_dict(), tmp_path, "per.buffer"))
    assert torch.equal(a._it_sum, b._it_sum) and torch.equal(a._it_min, b._it_min)
    assert torch.equal(a._max_priority, b._max_priority) and (a.ptr, a.size) == (b.ptr, b.size)
    r1, r2 = np.random.default_rng(5), np.random.default_rng(5)
    feed(a, 5, r1), feed(b, 5, r2)
    u = rng.random((N, B // N))
    sa, sb = a.sample(0.5, uniforms=u), b.sample(0.5, uniforms=u)
    for k in ("obs", "actions", "obs_next", "rewards", "terminals", "weights", "step_choices"):
        assert torch.equal(sa[k], sb[k]), k


def test_episode_replay_checkpoint(tmp_path):
    from xuance_b200.common import MARL_OffPolicyBuffer_RNN, Box, Discrete
    n, obs_dim, A, S, T, n_envs, C, Be = 3, 10, 5, 12, 8, 2, 6, 4
    keys = [f"agent_{i}" for i in range(n)]
    mk = lambda: MARL_OffPolicyBuffer_RNN(agent_keys=keys, state_space=Box(-1, 1, (S,)),
                                          obs_space={k: Box(-1, 1, (obs_dim,)) for k in keys},
                                          act_space={k: Discrete(A) for k in keys}, n_envs=n_envs, buffer_size=C,
                                          batch_size=Be, max_episode_steps=T, device=DEV)
    a = mk()
    events = list(qmix_episode_stream(np.random.default_rng(3), keys, n_envs, T, obs_dim, A, S, 5))
    cut = len(events) * 2 // 3

    def play(buf, evs):
        for ev in evs:
            if ev[0] == 'store':
                buf.store(**ev[1])
            else:
                buf.finish_path(ev[1], **ev[2])

    play(a, events[:cut])
    b = mk()
    b.load_state_dict(_roundtrip(a.state_dict(), tmp_path, "episodes.buffer"))
    play(a, events[cut:]), play(b, events[cut:])      # the restored buffer also carries the unfinished episodes
    assert (a.ptr, a.size) == (b.ptr, b.size)
    np.random.seed(0)
    sa = a.sample()
    np.random.seed(0)
    sb = b.sample()
    for k, v in sa["_stacked"].items():
        assert torch.equal(v, sb["_stacked"][k]), k


def test_agent_saves_and_restores_replay(tmp_path):
    from xuance_b200.common.common_tools import get_arguments
    from xuance_b200.environment import make_envs
    from xuance_b200.torch.agents import REGISTRY_Agents
    cfg = get_arguments("perdqn", "atari", "atari", parser_args=Namespace(
        device=DEV, parallels=2, buffer_size=2 * 64, batch_size=16, start_training=32, running_steps=1000,
        sync_frequency=10, model_dir=str(tmp_path / "models")))
    envs = make_envs(cfg)
    agent = REGISTRY_Agents[cfg.agent](cfg, envs)
    agent.model_dir_save = str(tmp_path / "models" / "seed_1")
    agent.train(40)
    agent.save_model("m.pth", save_buffer=True)
    agent2 = REGISTRY_Agents[cfg.agent](cfg, envs)
    assert agent2.memory.size == 0
    agent2.load_model(str(tmp_path / "models" / "seed_1" / "m.pth"), load_buffer=True)
    assert agent2.memory.size == agent.memory.size == 40
    assert torch.equal(agent2.memory._it_sum, agent.memory._it_sum)
    assert torch.equal(agent2.memory.observations, agent.memory.observations)
    for k, v in agent.model.state_dict().items():
        assert torch.equal(v, agent2.model.state_dict()[k]), k
    envs.close()

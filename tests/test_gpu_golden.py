"""GPU parity against the COMMITTED reference outputs (tests/golden/*.npz, written by tests/golden/make_golden.py from
the unmodified reference): the CUDA path is compared with what agi-brain/xuance itself produced, without the oracle's
arithmetic in between.  (The oracle's network classes are used only to rebuild the seeded initial weights of the two
fixtures that store digests of them instead of the weights.)

Tolerances: bytes / indices / PER trees and weights bit-exact; fp32 scans and learner scalars as stated per assert."""
import os
import random
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import fill_buffers, build_product_ppo_model, ppo_config, qmix_episode_stream

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _check_digest(k, v, g, rtol):
    """sum |w| of the whole tensor: relative tolerance plus 2e-5 per element (zero-initialised biases end at ~lr per
    element, where only an absolute bound is meaningful)."""
    flat = v.detach().reshape(-1).double().cpu()
    np.testing.assert_allclose(flat.abs().sum().item(), g[f"final_digest/{k}"][1], rtol=rtol, atol=2e-5 * flat.numel(),
                               err_msg=k)


def _check_final(model, g, head_tol, digest_rtol):
    for k, v in model.state_dict().items():
        ref = g[f"final_head/{k}"]
        np.testing.assert_allclose(v.detach().reshape(-1)[: ref.shape[0]].cpu().numpy(), ref, err_msg=k, **head_tol)
        _check_digest(k, v, g, digest_rtol)


@pytest.mark.parametrize("case,shape,atari,use_gae", [("vec_gae", (5,), False, True), ("vec_nstep", (5,), False, False),
                                                       ("atari_gae", (8, 8, 4), True, True)])
def test_onpolicy_buffer_vs_reference_fixture(case, shape, atari, use_gae):
    """K1 store + K2 scan + K3 gathers vs DummyOnPolicyBuffer[_Atari] of the reference (memory_tools.py:163-287)."""
    from xuance_b200.common import DummyOnPolicyBuffer, DummyOnPolicyBuffer_Atari, Box, Discrete
    g = _load("onpolicy.npz")
    ro = {k: g[f"{case}/in/{k}"] for k in ("obs", "acts", "rews", "vals", "terms", "logp", "boot")}
    T, N = ro["rews"].shape
    cls = DummyOnPolicyBuffer_Atari if atari else DummyOnPolicyBuffer
    space = Box(0, 255, shape, np.uint8) if atari else Box(-10, 10, shape, np.float32)
    prod = cls(space, Discrete(4), {"old_logp": ()}, N, T, use_gae=use_gae, use_advnorm=True, device=DEV)
    mids = [(int(t), int(e), float(v) if py else np.float32(v)) for t, e, v, py in g[f"{case}/in/mids"]]
    fill_buffers([prod], ro, mids)
    tol = dict(rtol=1e-5, atol=2e-6)    # the warp scan reorders fp32 operations relative to the reference's loop
    np.testing.assert_allclose(prod.returns.cpu().numpy(), g[f"{case}/returns"], **tol)
    np.testing.assert_allclose(prod.advantages.cpu().numpy(), g[f"{case}/advantages"], **tol)
    assert np.array_equal(prod.start_ids, g[f"{case}/start_ids"])
    s = prod.sample(g[f"{case}/in/idx"])
    for k in ("obs", "actions", "values"):                      # gathered bytes: bit-exact, same dtype
        ref = g[f"{case}/sample/{k}"]
        got = s[k].cpu().numpy()
        assert np.array_equal(got, ref) and got.dtype == ref.dtype, k
    assert np.array_equal(s["aux_batch"]["old_logp"].cpu().numpy(), g[f"{case}/sample/old_logp"])
    np.testing.assert_allclose(s["returns"].cpu().numpy(), g[f"{case}/sample/returns"], **tol)
    np.testing.assert_allclose(s["advantages"].cpu().numpy(), g[f"{case}/sample/advantages"], rtol=2e-5, atol=5e-6)


def test_per_replay_vs_reference_fixture():
    """K5 vs PerOffPolicyBuffer + SumSegmentTree / MinSegmentTree of the reference: every index, weight and tree node
    bit-exact (memory_tools.py:518-598, segtree_tool.py)."""
    from xuance_b200.common import PerOffPolicyBuffer, Box, Discrete
    g = _load("per.npz")
    N, S, B, alpha = g["cfg"]
    N, S, B = int(N), int(S), int(B)
    prod = PerOffPolicyBuffer(Box(-9, 9, (3,), np.float32), Discrete(4), None, N, N * S, B, alpha=float(alpha),
                              device=DEV)
    ins = [g[f"in/{k}"] for k in ("obs", "acts", "rews", "terms", "next_obs")]
    j = 0
    for t in range(ins[0].shape[0]):
        prod.store(*[a[t] for a in ins])
        if j < int(g["n_samples"]) and int(g[f"s{j}/t"]) == t:
            s = prod.sample(0.4, uniforms=g[f"s{j}/u"])
            assert np.array_equal(s["step_choices"].cpu().numpy(), g[f"s{j}/step_choices"]), j
            assert np.array_equal(s["weights"].cpu().numpy(), g[f"s{j}/weights"]), j
            assert np.array_equal(s["obs"].cpu().numpy(), g[f"s{j}/obs"])
            assert np.array_equal(s["rewards"].cpu().numpy(), g[f"s{j}/rewards"])
            random.seed(t)                                      # default path: random.random() in the reference's order
            s2 = prod.sample(0.4)
            assert np.array_equal(s2["step_choices"].cpu().numpy(), g[f"s{j}/step_choices"]), j
            prod.update_priorities(s["step_choices"], g[f"s{j}/td"])
            ps = prod._it_sum.cpu().numpy().astype(np.float64)
            pm = prod._it_min.cpu().numpy().astype(np.float64)
            for i in range(N):
                assert np.array_equal(ps[i], g[f"s{j}/sum_tree"][i]), (j, i)
                assert np.array_equal(pm[i], g[f"s{j}/min_tree"][i]), (j, i)
            assert np.array_equal(prod._max_priority.cpu().numpy().astype(np.float64), g[f"s{j}/max_priority"])
            j += 1
    assert j == int(g["n_samples"])


def test_ppo_update_vs_reference_fixture():
    """PPO_Learner.update (K3-obs formats, cuDNN fp32, K4, K7) vs two updates recorded from the reference learner."""
    from oracle.nets import SharedActorCriticOracle        # seeded construction only: rebuilds the fixture's init weights
    from xuance_b200.torch.learners import PPO_Learner
    from xuance_b200.common import BaseCallback
    _no_tf32()
    g = _load("ppo_update.npz")
    A, B = 6, 24
    torch.manual_seed(3)
    init = SharedActorCriticOracle(A).state_dict()
    model = build_product_ppo_model(A, DEV)
    model.load_state_dict(init)
    cfg = ppo_config(DEV, running_steps=4096 * 10, parallels=32)
    learner = PPO_Learner(cfg, model, BaseCallback())
    assert learner.total_iters == int(g["total_iters"])
    rng = np.random.default_rng(9)
    for it in range(2):
        s = {"obs": torch.from_numpy(rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8)).cuda(),
             "actions": rng.integers(0, A, size=B).astype(np.float32),
             "returns": rng.normal(size=B).astype(np.float32), "advantages": rng.normal(size=B).astype(np.float32),
             "aux_batch": {"old_logp": (rng.normal(size=B) * 0.05 - np.log(A)).astype(np.float32)}}
        info = learner.update(**s)
        got = [info["actor_loss"], info["critic_loss"], info["entropy"], info["learning_rate"], info["predict_value"]]
        np.testing.assert_allclose(got, g["infos"][it][:5], rtol=2e-4, atol=1e-5)      # fp32 conv/GEMM summation order
        np.testing.assert_allclose(float(info["clip_ratio"]), g["infos"][it][5], atol=1.0 / B + 1e-7)
    _check_final(model, g, dict(rtol=1e-3, atol=1e-4), 1e-4)       # Adam: |step| ~ lr, sign-sensitive near zero gradients


def test_perdqn_update_vs_reference_fixture():
    """PerDQN_Learner.update (K6 + K7 + target sync) vs three updates recorded from the reference learner."""
    from oracle.nets import DeepQNetworkOracle               # seeded construction only
    from xuance_b200.common import Discrete, BaseCallback
    from xuance_b200.torch.rl_models import Basic_CNN, DeepQNetwork
    from xuance_b200.torch.learners import PerDQN_Learner
    _no_tf32()
    g = _load("dqn_update.npz")
    A, B = 5, 16
    torch.manual_seed(4)
    init = DeepQNetworkOracle(A).state_dict()
    rep = Basic_CNN(input_shape=(84, 84, 4), kernels=[8, 4, 3], strides=[4, 2, 1], filters=[32, 64, 64],
                    activation=nn.ReLU, device=DEV)
    model = DeepQNetwork(rep, [512], Discrete(A), None, None, nn.ReLU, DEV).to(DEV)
    model.load_state_dict(init)
    cfg = Namespace(distributed_training=False, episode_length=1000, use_grad_clip=False, grad_clip_norm=0.5,
                    device=DEV, model_dir="/tmp/xb", running_steps=4096 * 10, parallels=32, learning_rate=1e-4,
                    end_factor_lr_decay=0.5, gamma=0.99, sync_frequency=2, start_training=0, training_frequency=1)
    lrn = PerDQN_Learner(cfg, model, BaseCallback())
    rng = np.random.default_rng(10)
    for it in range(3):
        s = {"obs": torch.from_numpy(rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8)).cuda(),
             "actions": rng.integers(0, A, size=B).astype(np.float32),
             "obs_next": torch.from_numpy(rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8)).cuda(),
             "rewards": rng.normal(size=B).astype(np.float32), "terminals": (rng.random(B) < 0.2).astype(np.float32)}
        td, info = lrn.update(**s)
        np.testing.assert_allclose([info["Qloss"], info["learning_rate"], info["predictQ"]], g["infos"][it],
                                   rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(td.cpu().numpy(), g["abs_td"][it], rtol=2e-3, atol=2e-5)
    for k, v in model.state_dict().items():
        _check_digest(k, v, g, 1e-4)


def test_sac_update_vs_reference_fixture():
    """SAC_Learner.update (K8 losses, three K7 steps, Polyak) vs three updates recorded from the reference learner; the
    fixture holds the initial weights and the noise the reference drew."""
    from copy import deepcopy
    from xuance_b200.common import Box, BaseCallback
    from xuance_b200.torch.rl_models import Basic_Identical, SAC_GaussianActor, TwinActionValueCritic, SoftActorCritic
    from xuance_b200.torch.learners.sac_learner import SAC_Learner
    _no_tf32()
    g = _load("sac_update.npz")
    obs_dim, act_dim = 17, 6
    aspace = Box(-1, 1, (act_dim,), np.float32)
    rep = Basic_Identical((obs_dim,), device=DEV)
    actor = SAC_GaussianActor(rep, [64, 64], aspace, None, None, nn.LeakyReLU, nn.Tanh, DEV)
    critic = TwinActionValueCritic(deepcopy(rep), aspace, [64, 64], None, None, nn.LeakyReLU, DEV)
    model = SoftActorCritic(actor, critic).to(DEV)
    model.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init/")}, strict=True)
    cfg = Namespace(distributed_training=False, episode_length=1000, use_grad_clip=False, grad_clip_norm=0.5,
                    device=DEV, model_dir="/tmp/xb", running_steps=100000, parallels=4, start_training=0,
                    training_frequency=1, learning_rate_actor=1e-3, learning_rate_critic=1e-3, tau=0.005, gamma=0.99,
                    alpha=0.2, use_automatic_entropy_tuning=True, end_factor_lr_decay=0.7, use_cuda_graph=False)
    lrn = SAC_Learner(cfg, model, BaseCallback())
    assert lrn.total_iters == int(g["total_iters"])
    for it in range(3):
        s = {k: torch.from_numpy(g[f"in{it}/{k}"]).cuda() for k in ("obs", "actions", "obs_next", "rewards", "terminals")}
        info = lrn.update(noise_pi=torch.from_numpy(g[f"noise_pi/{it}"]).cuda(),
                          noise_next=torch.from_numpy(g[f"noise_next/{it}"]).cuda(), **s)
        ref = g["infos"][it]        # Qloss, Ploss, Qvalue, alpha_loss, alpha, actor_lr
        np.testing.assert_allclose([info["Qloss"], info["Ploss"], info["Qvalue"]], ref[:3], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(info["alpha_loss"], ref[3], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(info["alpha"], ref[4], rtol=1e-5)
        assert info["actor_lr"] == ref[5]
    _check_final(model, g, dict(rtol=2e-3, atol=2e-4), 1e-3)


def test_qmix_replay_and_update_vs_reference_fixture():
    """MARL_OffPolicyBuffer_RNN (bit-exact) and QMIX_Learner.update (K9 kernels, K7) vs the reference's recorded episode
    replay and three updates; qmix_rnn_detach_q_eval=True selects the reference's as-is behaviour."""
    from xuance_b200.common import (AgentGrouping, Discrete, Box, BaseCallback, MARL_OffPolicyBuffer_RNN)
    from xuance_b200.torch.rl_models import (Basic_RNN, AgentFeatureEncoder, DiscreteActionValueCritic, QMIX_Mixer,
                                             MixingQNetwork)
    from xuance_b200.torch.learners.qmix_learner import QMIX_Learner
    _no_tf32()
    g = _load("qmix_update.npz")
    n, obs_dim, A, S, T, n_envs, C, Be = (int(x) for x in g["cfg"])
    keys = [f"agent_{i}" for i in range(n)]
    prod = MARL_OffPolicyBuffer_RNN(agent_keys=keys, state_space=Box(-1, 1, (S,)),
                                    obs_space={k: Box(-1, 1, (obs_dim,)) for k in keys},
                                    act_space={k: Discrete(A) for k in keys}, n_envs=n_envs, buffer_size=C,
                                    batch_size=Be, max_episode_steps=T, device=DEV)
    for ev in qmix_episode_stream(np.random.default_rng(21), keys, n_envs, T, obs_dim, A, S, 5):
        if ev[0] == 'store':
            prod.store(**ev[1])
        else:
            prod.finish_path(ev[1], **ev[2])
    assert [prod.ptr, int(prod.size)] == list(g["buffer/ptr_size"])
    # whole-buffer contents through the public sample(): draw every slot once
    rep = Basic_RNN(input_shape=(obs_dim,), hidden_sizes=None, initialize=nn.init.orthogonal_, activation=nn.ReLU,
                    device=DEV, fc_hidden_sizes=[64], recurrent_hidden_size=64, N_recurrent_layers=1, dropout=0, rnn='GRU')
    q = nn.ModuleDict({'shared': DiscreteActionValueCritic(AgentFeatureEncoder(rep), Discrete(A), [64], None,
                                                           nn.init.orthogonal_, nn.ReLU, DEV)})
    grouping = AgentGrouping.shared(keys)
    model = MixingQNetwork(grouping, q, QMIX_Mixer(S, 32, 32, n, DEV), use_rnn=True, device=DEV).to(DEV)
    init = {k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init/")}
    model.load_state_dict(init, strict=True)
    cfg = Namespace(distributed_training=False, episode_length=T, use_grad_clip=False, grad_clip_norm=10.0, device=DEV,
                    model_dir="/tmp/x", running_steps=100000, parallels=4, use_parameter_sharing=True, use_rnn=True,
                    use_actions_mask=False, learning_rate=7e-4, sync_frequency=2, double_q=True, n_epochs=1,
                    start_training=0, gamma=0.99, end_factor_lr_decay=0.5, qmix_rnn_detach_q_eval=True,
                    use_cuda_graph=False)
    lrn = QMIX_Learner(cfg, grouping, model, BaseCallback())
    assert lrn.total_iters == int(g["total_iters"])
    for it in range(3):
        np.random.seed(it)
        sample = prod.sample()
        if it == 0:     # the sampled rows are the reference's rows, bit for bit
            np.random.seed(0)
            rows = np.random.choice(int(prod.size), Be)
            for kname in ('obs', 'actions', 'rewards', 'terminals', 'agent_mask'):
                for i, a in enumerate(keys):
                    ref = g[f"buffer/{kname}"][rows, i]
                    assert np.array_equal(sample[kname][a].cpu().numpy(), ref), (kname, a)
            assert np.array_equal(sample['filled'].cpu().numpy(), g["buffer/filled"][rows])
            assert np.array_equal(sample['state'].cpu().numpy(), g["buffer/state"][rows])
        info = lrn.update(sample)
        np.testing.assert_allclose([info["loss_Q"], info["predictQ"]], g["infos"][it][:2], rtol=5e-4, atol=1e-5)
        assert info["learning_rate"] == g["infos"][it][2]
    _check_final(model, g, dict(rtol=2e-3, atol=3e-4), 1e-3)
    sd = model.state_dict()
    for k in sd:    # the reference run left its agent networks untouched; so does the as-is mode
        if k.startswith("individual_q_networks"):
            assert torch.equal(sd[k].cpu(), init[k]), k

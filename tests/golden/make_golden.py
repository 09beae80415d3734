"""Generate tests/golden/*.npz by running the UNMODIFIED reference (agi-brain/xuance v1.4.4 @ 4f0b05b).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference is imported read-only with two import stubs (oracle/ref_stubs: gymnasium, pyglet).  Each fixture
stores the seeded inputs that cannot be re-derived cheaply plus the reference's outputs; the CPU tests re-run
the oracle on the same inputs and compare.  Environment of record is written into every file."""
import os
import random
import sys
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_loader import import_reference  # noqa: E402

import_reference()
from gymnasium.spaces import Box, Discrete  # noqa: E402
from xuance.common import BaseCallback  # noqa: E402
from xuance.common.common_tools import discount_cumsum  # noqa: E402
from xuance.common.memory_tools import (DummyOnPolicyBuffer, DummyOnPolicyBuffer_Atari, PerOffPolicyBuffer,  # noqa
                                        DummyOffPolicyBuffer)
from xuance.torch.learners import PPO_Learner, DQN_Learner, PerDQN_Learner  # noqa: E402
from xuance.torch.rl_models.representations import AC_CNN_Atari, Basic_CNN  # noqa: E402
from xuance.torch.rl_models.heads import CategoricalActorHead, ValueHead  # noqa: E402
from xuance.torch.rl_models.architectures.single_agent.actor_critic import SharedActorCritic  # noqa: E402
from xuance.torch.rl_models.architectures.single_agent.deep_q_network import DeepQNetwork  # noqa: E402
from helpers import synth_rollout, fill_buffers  # noqa: E402

ENV = "torch %s / numpy %s / reference 4f0b05b" % (torch.__version__, np.__version__)


def golden_onpolicy():
    out = {"env": ENV}
    cases = [("vec_gae", 6, 24, (5,), False, True), ("vec_nstep", 6, 24, (5,), False, False),
             ("atari_gae", 4, 16, (8, 8, 4), True, True)]
    for name, N, T, shape, atari, use_gae in cases:
        rng = np.random.default_rng(hash(name) % 1000 if False else len(name) * 7 + N)
        ro = synth_rollout(rng, N, T, shape, obs_dtype=np.uint8 if atari else np.float32, p_term=0.08)
        cls = DummyOnPolicyBuffer_Atari if atari else DummyOnPolicyBuffer
        space = Box(0, 255, shape, np.uint8) if atari else Box(-10, 10, shape, np.float32)
        ref = cls(space, Discrete(4), {"old_logp": ()}, N, T, use_gae=use_gae)
        mids = [(T // 3, 1, np.float32(0.37)), (T // 2, 0, 0.0), (T // 2 + 1, 0, np.float32(-1.5))]
        fill_buffers([ref], ro, mids)
        idx = rng.permutation(N * T)[: N * T // 2]
        s = ref.sample(idx)
        for k, v in ro.items():
            out[f"{name}/in/{k}"] = v
        # 4th column: 1 if the bootstrap was a Python float (np.append then promotes the path to float64, as
        # happens for the reference's finish_path(0.0, i) calls), 0 if it was an np.float32 (model output)
        out[f"{name}/in/mids"] = np.array([(t, e, float(v), float(isinstance(v, float))) for t, e, v in mids])
        out[f"{name}/in/idx"] = idx
        out[f"{name}/returns"], out[f"{name}/advantages"] = ref.returns, ref.advantages
        out[f"{name}/start_ids"] = ref.start_ids
        for k in ("obs", "actions", "returns", "values", "advantages"):
            out[f"{name}/sample/{k}"] = s[k]
        out[f"{name}/sample/old_logp"] = s["aux_batch"]["old_logp"]
    out["kat/discount_cumsum"] = np.asarray(discount_cumsum(np.array([0, 1, 2, 2]), 0.99))
    np.savez_compressed(os.path.join(HERE, "onpolicy.npz"), **out)


def golden_per():
    """PER under the canonical float32 rule: the reference's float64 ``_max_priority`` array is replaced by a float32
    one right after construction (state change in the test harness, no reference source change) - see oracle/replay.py."""
    out = {"env": ENV}
    N, S, B, alpha = 4, 48, 32, 0.5
    rng = np.random.default_rng(42)
    ref = PerOffPolicyBuffer(Box(-1, 1, (3,)), Discrete(4), None, N, N * S, B, alpha=alpha)
    ref._max_priority = np.ones(N, np.float32)
    steps, samples = [], []
    for t in range(70):
        st = (rng.normal(size=(N, 3)).astype(np.float32), rng.integers(0, 4, N), rng.normal(size=N).astype(np.float32),
              rng.random(N) < 0.1, rng.normal(size=(N, 3)).astype(np.float32))
        ref.store(*st)
        steps.append(st)
        if t > 5 and t % 6 == 0:
            random.seed(t)
            u = np.array([[random.random() for _ in range(B // N)] for _ in range(N)])
            random.seed(t)
            s = ref.sample(0.4)
            td = np.abs(rng.normal(size=B)).astype(np.float32)
            td[2] = 0.0
            ref.update_priorities(s["step_choices"], td)
            samples.append(dict(t=t, u=u, step_choices=s["step_choices"], weights=s["weights"], td=td,
                                obs=s["obs"], rewards=s["rewards"],
                                sum_tree=np.array([np.array(tr._value, dtype=np.float64) for tr in ref._it_sum]),
                                min_tree=np.array([np.array(tr._value, dtype=np.float64) for tr in ref._it_min]),
                                max_priority=ref._max_priority.astype(np.float64).copy()))
    for i, name in enumerate(("obs", "acts", "rews", "terms", "next_obs")):
        out[f"in/{name}"] = np.stack([s[i] for s in steps])
    out["n_samples"] = len(samples)
    for j, s in enumerate(samples):
        for k, v in s.items():
            out[f"s{j}/{k}"] = v
    out["cfg"] = np.array([N, S, B, alpha])
    # the default (mixed float64/float32) behaviour of the reference under NumPy>=2, recorded for DESIGN.md
    ref2 = PerOffPolicyBuffer(Box(-1, 1, (3,)), Discrete(4), None, 2, 8, 2, alpha=alpha)
    ref2.store(np.zeros((2, 3)), np.zeros(2), np.zeros(2), np.zeros(2), np.zeros((2, 3)))
    ref2.store(np.zeros((2, 3)), np.zeros(2), np.zeros(2), np.zeros(2), np.zeros((2, 3)))
    ref2.update_priorities(np.array([[0], [1]]), np.array([0.3, 0.7], np.float32))
    out["note/default_leaf_dtypes"] = np.array([type(ref2._it_sum[0]._value[4]).__name__,
                                                type(ref2._it_sum[0]._value[5]).__name__])
    np.savez_compressed(os.path.join(HERE, "per.npz"), **out)


def _learner_cfg(**kw):
    cfg = dict(distributed_training=False, episode_length=1000, use_grad_clip=True, grad_clip_norm=0.5, device="cpu",
               model_dir="/tmp/xb_golden", running_steps=4096 * 10, parallels=32, learning_rate=2.5e-4,
               use_linear_lr_decay=True, end_factor_lr_decay=0.5, horizon_size=128, n_epochs=4, n_minibatch=4,
               vf_coef=0.25, ent_coef=0.01, clip_range=0.2, gamma=0.99, sync_frequency=2, start_training=0,
               training_frequency=1)
    cfg.update(kw)
    return Namespace(**cfg)


def golden_ppo():
    """Two PPO updates of the reference learner on the NatureCNN actor-critic at 84x84x4, B=24.  Initial weights are
    reproduced from torch.manual_seed (construction order is part of the oracle's contract); the fixture keeps
    inputs, info dicts and per-tensor digests of the parameters after the updates."""
    out = {"env": ENV}
    A, B = 6, 24
    torch.manual_seed(3)
    rep = AC_CNN_Atari(input_shape=(84, 84, 4), kernels=[8, 4, 3], strides=[4, 2, 1], filters=[32, 64, 64],
                       activation=torch.nn.ReLU, device="cpu", fc_hidden_sizes=[512])
    actor = CategoricalActorHead(512, [], A, None, torch.nn.init.orthogonal_, torch.nn.ReLU, "cpu")
    critic = ValueHead(512, [], None, torch.nn.init.orthogonal_, torch.nn.ReLU, "cpu")
    model = SharedActorCritic(rep, actor, critic)
    init = {k: v.clone() for k, v in model.state_dict().items()}
    learner = PPO_Learner(_learner_cfg(), model, BaseCallback())
    rng = np.random.default_rng(9)
    infos = []
    for it in range(2):
        s = {"obs": rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8),
             "actions": rng.integers(0, A, size=B).astype(np.float32),
             "returns": rng.normal(size=B).astype(np.float32), "advantages": rng.normal(size=B).astype(np.float32),
             "aux_batch": {"old_logp": (rng.normal(size=B) * 0.05 - np.log(A)).astype(np.float32)}}
        info = learner.update(**s)
        infos.append([info["actor_loss"], info["critic_loss"], info["entropy"], info["learning_rate"],
                      info["predict_value"], float(info["clip_ratio"])])
    out["infos"] = np.array(infos, dtype=np.float64)
    out["total_iters"] = learner.total_iters
    for k, v in model.state_dict().items():
        out[f"init_digest/{k}"] = np.array([v_.double().sum().item() for v_ in (init[k], init[k].abs())])
        flat = v.detach().reshape(-1)
        out[f"final_head/{k}"] = flat[:64].numpy().copy()
        out[f"final_digest/{k}"] = np.array([flat.double().sum().item(), flat.double().abs().sum().item()])
    np.savez_compressed(os.path.join(HERE, "ppo_update.npz"), **out)


def golden_dqn():
    out = {"env": ENV}
    A, B = 5, 16
    torch.manual_seed(4)
    rep = Basic_CNN(input_shape=(84, 84, 4), kernels=[8, 4, 3], strides=[4, 2, 1], filters=[32, 64, 64],
                    activation=torch.nn.ReLU, device="cpu")
    model = DeepQNetwork(rep, [512], Discrete(A), None, None, torch.nn.ReLU, "cpu")
    learner = PerDQN_Learner(_learner_cfg(learning_rate=1e-4, use_grad_clip=False), model, BaseCallback())
    rng = np.random.default_rng(10)
    infos, tds = [], []
    for it in range(3):
        s = {"obs": rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8).astype(np.float32),
             "actions": rng.integers(0, A, size=B).astype(np.float32),
             "obs_next": rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8).astype(np.float32),
             "rewards": rng.normal(size=B).astype(np.float32), "terminals": (rng.random(B) < 0.2).astype(np.float32)}
        td, info = learner.update(**s)
        infos.append([info["Qloss"], info["learning_rate"], info["predictQ"]])
        tds.append(td)
    out["infos"], out["abs_td"] = np.array(infos, dtype=np.float64), np.stack(tds)
    for k, v in model.state_dict().items():
        flat = v.detach().reshape(-1)
        out[f"final_digest/{k}"] = np.array([flat.double().sum().item(), flat.double().abs().sum().item()])
    np.savez_compressed(os.path.join(HERE, "dqn_update.npz"), **out)


def golden_sac():
    """Three SAC updates of the reference learner (17-d obs / 6-d action, MLP 64-64, automatic entropy tuning).  The
    reference draws its reparameterisation noise from torch's global RNG; the fixture stores the noise it drew (the stream
    is replayed: randn(B, act) for the actor step, then randn(B, act) for the target step)."""
    from copy import deepcopy
    from xuance.torch.rl_models.representations.mlp import Basic_Identical
    from xuance.torch.rl_models.actors.gaussian_actors import SAC_GaussianActor
    from xuance.torch.rl_models.critics.twin_critics import TwinActionValueCritic
    from xuance.torch.rl_models.architectures.single_agent.actor_critic import SoftActorCritic
    from xuance.torch.learners import SAC_Learner
    out = {"env": ENV}
    torch.manual_seed(7)
    obs_dim, act_dim, B = 17, 6, 32
    aspace = Box(-1, 1, (act_dim,), np.float32)
    rep = Basic_Identical((obs_dim,), device='cpu')
    actor = SAC_GaussianActor(rep, [64, 64], aspace, None, None, torch.nn.LeakyReLU, torch.nn.Tanh, 'cpu')
    critic = TwinActionValueCritic(deepcopy(rep), aspace, [64, 64], None, None, torch.nn.LeakyReLU, 'cpu')
    model = SoftActorCritic(actor, critic)
    for k, v in model.state_dict().items():
        out[f"init/{k}"] = v.numpy().copy()
    cfg = _learner_cfg(learning_rate_actor=1e-3, learning_rate_critic=1e-3, tau=0.005, alpha=0.2, use_grad_clip=False,
                       use_automatic_entropy_tuning=True, end_factor_lr_decay=0.7, parallels=4, running_steps=100000)
    lrn = SAC_Learner(cfg, model, BaseCallback())
    out["total_iters"] = lrn.total_iters
    rng = np.random.default_rng(12)
    infos = []
    for it in range(3):
        s = {"obs": rng.normal(size=(B, obs_dim)).astype(np.float32),
             "actions": rng.uniform(-1, 1, size=(B, act_dim)).astype(np.float32),
             "obs_next": rng.normal(size=(B, obs_dim)).astype(np.float32),
             "rewards": rng.normal(size=B).astype(np.float32), "terminals": (rng.random(B) < 0.1).astype(np.float32)}
        torch.manual_seed(500 + it)
        out[f"noise_pi/{it}"] = torch.randn(B, act_dim).numpy()
        out[f"noise_next/{it}"] = torch.randn(B, act_dim).numpy()
        torch.manual_seed(500 + it)
        info = lrn.update(**s)
        for k, v in s.items():
            out[f"in{it}/{k}"] = v
        infos.append([info["Qloss"], info["Ploss"], info["Qvalue"], info["alpha_loss"], info["alpha"], info["actor_lr"]])
    out["infos"] = np.array(infos, dtype=np.float64)
    for k, v in model.state_dict().items():
        flat = v.detach().reshape(-1)
        out[f"final_head/{k}"] = flat[:32].numpy().copy()
        out[f"final_digest/{k}"] = np.array([flat.double().sum().item(), flat.double().abs().sum().item()])
    np.savez_compressed(os.path.join(HERE, "sac_update.npz"), **out)


def golden_qmix():
    """Episode replay + three QMIX updates of the reference (5 agents x 24-d obs, 7 actions, 30-d state, T=10, GRU 64,
    mixer 32/32, double-Q, use_actions_mask=False).  Pins BOTH reference behaviours recorded in DESIGN.md: the values and
    the fact that the agent networks do not move (q_eval is sliced under no_grad)."""
    from xuance.common import AgentGrouping
    from xuance.common.memory_tools_marl import MARL_OffPolicyBuffer_RNN
    from xuance.torch.rl_models.representations.rnn import Basic_RNN
    from xuance.torch.rl_models.representations.agent_feature import AgentFeatureEncoder
    from xuance.torch.rl_models.modules.identity_encoder import build_identity_encoder, IdentityFeatureFusion
    from xuance.torch.rl_models.critics.base_critics import DiscreteActionValueCritic
    from xuance.torch.rl_models.heads.q_mix_head import QMIX_Mixer
    from xuance.torch.rl_models.architectures.multi_agent.value_factorization import MixingQNetwork
    from xuance.torch.learners.multi_agent_rl.qmix_learner import QMIX_Learner
    from helpers import qmix_episode_stream
    out = {"env": ENV}
    torch.manual_seed(8)
    n, obs_dim, A, S, T = 5, 24, 7, 30, 10
    keys = [f"agent_{i}" for i in range(n)]
    rep = Basic_RNN(input_shape=(obs_dim,), hidden_sizes=None, initialize=torch.nn.init.orthogonal_,
                    activation=torch.nn.ReLU, device='cpu', fc_hidden_sizes=[64], recurrent_hidden_size=64,
                    N_recurrent_layers=1, dropout=0, rnn='GRU')
    enc = AgentFeatureEncoder(rep, build_identity_encoder(n, 'none', None, 'cpu'), IdentityFeatureFusion(64, 0, 'concat'))
    q = torch.nn.ModuleDict({'shared': DiscreteActionValueCritic(enc, Discrete(A), [64], None, torch.nn.init.orthogonal_,
                                                                   torch.nn.ReLU, 'cpu')})
    grouping = AgentGrouping.shared(keys)
    model = MixingQNetwork(grouping, q, QMIX_Mixer(S, 32, 32, n, 'cpu'), use_rnn=True, device='cpu')
    for k, v in model.state_dict().items():
        out[f"init/{k}"] = v.numpy().copy()
    cfg = _learner_cfg(episode_length=T, use_grad_clip=False, parallels=4, running_steps=100000,
                       use_parameter_sharing=True, use_rnn=True, use_actions_mask=False, learning_rate=7e-4,
                       sync_frequency=2, double_q=True, n_epochs=1, end_factor_lr_decay=0.5)
    lrn = QMIX_Learner(cfg, grouping, model, BaseCallback())
    out["total_iters"] = lrn.total_iters
    n_envs, C, Be = 3, 12, 6
    rb = MARL_OffPolicyBuffer_RNN(agent_keys=keys, state_space=Box(-1, 1, (S,)),
                                  obs_space={k: Box(-1, 1, (obs_dim,)) for k in keys},
                                  act_space={k: Discrete(A) for k in keys}, n_envs=n_envs, buffer_size=C, batch_size=Be,
                                  max_episode_steps=T, use_actions_mask=False)
    for ev in qmix_episode_stream(np.random.default_rng(21), keys, n_envs, T, obs_dim, A, S, 5):
        if ev[0] == 'store':
            rb.store(**ev[1])
        else:
            rb.finish_path(ev[1], **ev[2])
    out["cfg"] = np.array([n, obs_dim, A, S, T, n_envs, C, Be])
    out["buffer/ptr_size"] = np.array([rb.ptr, int(rb.size)])
    out["buffer/filled"], out["buffer/state"] = rb.data['filled'], rb.data['state']
    for kname in ('obs', 'actions', 'rewards', 'terminals', 'agent_mask'):
        out[f"buffer/{kname}"] = np.stack([rb.data[kname][a] for a in keys], axis=1)
    infos = []
    for it in range(3):
        np.random.seed(it)
        s = rb.sample()
        info = lrn.update(s)
        infos.append([info["loss_Q"], info["predictQ"], info["learning_rate"]])
    out["infos"] = np.array(infos, dtype=np.float64)
    for k, v in model.state_dict().items():
        flat = v.detach().reshape(-1)
        out[f"final_head/{k}"] = flat[:32].numpy().copy()
        out[f"final_digest/{k}"] = np.array([flat.double().sum().item(), flat.double().abs().sum().item()])
    np.savez_compressed(os.path.join(HERE, "qmix_update.npz"), **out)


def golden_rollout_glue():
    """RunningMeanStd + Agent._process_observation arithmetic (statistic_tools.py:117-185, agent.py:273-276) and
    CategoricalDistribution log_prob / entropy / deterministic_sample (distributions.py:128-162) of the reference."""
    from xuance.common.statistic_tools import RunningMeanStd
    from xuance.common.common_tools import EPS
    from xuance.torch.rl_models.modules.distributions import CategoricalDistribution
    out = {"env": ENV}
    rng = np.random.default_rng(33)
    for name, N, D in (("cartpole", 8, 4), ("wide", 256, 17)):
        rms = RunningMeanStd(shape=(D,))
        for it in range(6):
            x = (rng.normal(size=(N, D)) * (1.0 + 3.0 * it) + 0.5 * it).astype(np.float32)
            rms.update(x)
            y = np.clip((x - rms.mean) / (rms.std + EPS), -5, 5)
            out[f"rms/{name}/x{it}"], out[f"rms/{name}/y{it}"] = x, y
            out[f"rms/{name}/mean{it}"], out[f"rms/{name}/var{it}"] = rms.mean.copy(), rms.var.copy()
            out[f"rms/{name}/count{it}"] = np.float64(rms.count)
    for A in (2, 4, 18):
        logits = (rng.normal(size=(64, A)) * 2.0).astype(np.float32)
        acts = rng.integers(0, A, size=64)
        d = CategoricalDistribution(A)
        d.set_param(logits=torch.from_numpy(logits))
        out[f"cat/{A}/logits"], out[f"cat/{A}/actions"] = logits, acts
        out[f"cat/{A}/log_prob"] = d.log_prob(torch.from_numpy(acts)).numpy()
        out[f"cat/{A}/entropy"] = d.entropy().numpy()
        out[f"cat/{A}/argmax"] = d.deterministic_sample().numpy()
        out[f"cat/{A}/probs"] = d.probs.numpy()
    np.savez_compressed(os.path.join(HERE, "rollout_glue.npz"), **out)


if __name__ == "__main__":
    golden_onpolicy()
    golden_per()
    golden_ppo()
    golden_dqn()
    golden_sac()
    golden_qmix()
    golden_rollout_glue()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))

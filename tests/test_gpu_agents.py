"""GPU: the reference-style end-to-end runs (the reference's own tests are exactly this - train for N steps, no
exception - SURVEY.md section 4), plus a learning check on BASELINE config 1 (PPO-Clip CartPole-v1, 8 envs)."""
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ppo_cartpole_learns():
    from xuance_b200 import get_runner
    runner = get_runner("ppo", "classic_control", "CartPole-v1",
                        parser_args=Namespace(device="cuda:0", running_steps=40960, model_dir="/tmp/xb_models/ppo"))
    agent = runner.agent
    assert agent.n_envs == 8 and agent.memory.n_size == 256
    info = runner.run("train", running_steps=40960)        # 20 rollouts of 8 x 256
    assert {"actor_loss", "critic_loss", "entropy", "learning_rate", "predict_value", "clip_ratio"} <= set(agent.logged)
    scores = agent.test(test_episodes=8)
    runner.finish()
    assert np.mean(scores) > 60.0, scores                  # a random policy scores ~22


def test_perdqn_and_sac_run_end_to_end():
    from xuance_b200.common import Box
    from xuance_b200.common.common_tools import get_arguments
    from xuance_b200.environment import make_envs
    from xuance_b200.torch.agents import REGISTRY_Agents
    # PER-DQN on the Atari-shaped synthetic env, small replay
    cfg = get_arguments("perdqn", "atari", "atari", parser_args=Namespace(
        device="cuda:0", parallels=4, buffer_size=4 * 256, batch_size=32, start_training=64, running_steps=2000,
        sync_frequency=10, model_dir="/tmp/xb_models/perdqn"))
    envs = make_envs(cfg)
    agent = REGISTRY_Agents[cfg.agent](cfg, envs)
    info = agent.train(60)
    assert "Qloss" in info and np.isfinite(info["Qloss"])
    assert agent.memory.size == 60 and float(agent.memory._max_priority.max()) >= 1.0
    envs.close()

    # SAC on a MuJoCo-shaped synthetic env
    class _Env:
        max_episode_steps = 50

        def __init__(self, seed=0):
            self.observation_space = Box(-10, 10, (17,), np.float32)
            self.action_space = Box(-1, 1, (6,), np.float32)
            self.rng, self.t = np.random.default_rng(seed), 0

        def reset(self, **kw):
            self.t = 0
            return self.rng.normal(size=17).astype(np.float32), {}

        def step(self, a):
            self.t += 1
            return self.rng.normal(size=17).astype(np.float32), float(-np.square(a).sum()), False, self.t >= 50, {}

        def close(self):
            pass

    from xuance_b200.environment import DummyVecEnv, XuanCeEnvWrapper
    cfg = get_arguments("sac", "mujoco", "mujoco", parser_args=Namespace(
        device="cuda:0", parallels=4, buffer_size=4000, batch_size=64, start_training=40, running_steps=2000,
        model_dir="/tmp/xb_models/sac"))
    envs = DummyVecEnv([lambda i=i: XuanCeEnvWrapper(_Env(i)) for i in range(4)])
    agent = REGISTRY_Agents["SAC"](cfg, envs)
    info = agent.train(80)
    assert np.isfinite(info["Qloss"]) and np.isfinite(info["Ploss"]) and info["alpha"] > 0
    envs.close()


def test_checkpoint_roundtrip_reference_format(tmp_path):
    from helpers import build_product_ppo_model, ppo_config
    from xuance_b200.common import BaseCallback
    from xuance_b200.torch.learners import PPO_Learner
    model = build_product_ppo_model(4, "cuda:0")
    lrn = PPO_Learner(ppo_config("cuda:0"), model, BaseCallback())
    rng = np.random.default_rng(0)
    B = 32
    s = {"obs": torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device="cuda:0"),
         "actions": rng.integers(0, 4, size=B).astype(np.float32), "returns": rng.normal(size=B).astype(np.float32),
         "advantages": rng.normal(size=B).astype(np.float32),
         "aux_batch": {"old_logp": (rng.normal(size=B) * 0.05 - 1.4).astype(np.float32)}}
    lrn.update(**s)
    path = str(tmp_path / "seed_1" / "final_train_model.pth")
    lrn.save_model(path)
    ckpt = torch.load(path, weights_only=True)
    assert set(ckpt) == {"policy", "optimizer", "rng_state", "cuda_rng_state"}          # drl_learner.py:64-93
    assert "representation.model.0.weight" in ckpt["policy"] and "exp_avg" in ckpt["optimizer"]["state"][0]
    model2 = build_product_ppo_model(4, "cuda:0")
    lrn2 = PPO_Learner(ppo_config("cuda:0"), model2, BaseCallback())
    lrn2.load_model(str(tmp_path))
    for a, b in zip(model.state_dict().values(), model2.state_dict().values()):
        assert torch.equal(a, b)
    assert torch.equal(lrn.optimizer.exp_avg, lrn2.optimizer.exp_avg) and lrn2.optimizer.step_count == 1
    i1, i2 = lrn.update(**s), lrn2.update(**s)
    assert abs(i1["actor_loss"] - i2["actor_loss"]) < 1e-6


def test_ppo_train_epochs_cuda_graph_matches_eager():
    """train_epochs with the minibatch update captured in a CUDA graph == the eager path (same kernels, same order)."""
    from xuance_b200.common import Box, Discrete
    from xuance_b200.torch.agents import PPO_Agent
    from helpers import synth_rollout, fill_buffers
    import copy
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    base = dict(agent="PPO", learner="PPO_Learner", representation="AC_CNN_Atari", env_name="Atari",
                distributed_training=False, device="cuda:0", seed=3, parallels=8, running_steps=100000, horizon_size=16,
                n_epochs=2, n_minibatch=2, learning_rate=2.5e-4, vf_coef=0.25, ent_coef=0.01, clip_range=0.2, gamma=0.99,
                use_gae=True, gae_lambda=0.95, use_advnorm=True, use_grad_clip=True, grad_clip_norm=0.5, use_obsnorm=False,
                use_rewnorm=False, activation="relu", filters=[32, 64, 64], kernels=[8, 4, 3], strides=[4, 2, 1],
                fc_hidden_sizes=[512], actor_hidden_size=[], critic_hidden_size=[], model_dir="/tmp/xb_models/g",
                logger=None, compute="fp32", episode_length=None)
    ro = synth_rollout(np.random.default_rng(0), 8, 16, (84, 84, 4))
    results = []
    for graph in (False, True):
        cfg = Namespace(**dict(base, use_cuda_graph=graph))
        agent = PPO_Agent(cfg, None, Box(0, 255, (84, 84, 4), np.uint8), Discrete(4))
        fill_buffers([agent.memory], ro)
        np.random.seed(11)
        info = agent.train_epochs(2)
        info2 = agent.train_epochs(2)          # second call replays the captured graph
        results.append((info2, {k: v.clone() for k, v in agent.model.state_dict().items()}))
    (i0, p0), (i1, p1) = results
    for k in ("actor_loss", "critic_loss", "entropy", "predict_value", "clip_ratio"):
        np.testing.assert_allclose(i1[k], i0[k], rtol=1e-4, atol=1e-6, err_msg=k)
    for k in p0:
        np.testing.assert_allclose(p1[k].cpu().numpy(), p0[k].cpu().numpy(), rtol=1e-3, atol=5e-5, err_msg=k)  # cuDNN wgrad atomics


def test_tensor_env_wrapper_feeds_the_hbm_buffer():
    """E4 (tensor_env.py:38-51): device-facing env wrapper -> buffer.store with CUDA tensors, uint8 frames stay uint8."""
    from xuance_b200.environment import make_envs, TensorEnvWrapper
    from xuance_b200.common import DummyOnPolicyBuffer_Atari
    envs = TensorEnvWrapper(make_envs(Namespace(env_id="SyntheticAtari", vectorize="Dummy_Atari", parallels=4, env_seed=2)),
                            "cuda:0")
    obs, _ = envs.reset()
    assert obs.is_cuda and obs.dtype == torch.uint8 and tuple(obs.shape) == (4, 84, 84, 4)
    buf = DummyOnPolicyBuffer_Atari(envs.observation_space, envs.action_space, {"old_logp": ()}, 4, 8, device="cuda:0")
    seen = []
    for t in range(8):
        acts = torch.randint(0, 4, (4,), device="cuda:0")
        nxt, rew, term, trunc, infos = envs.step(acts)
        buf.store(obs, acts, rew, torch.zeros(4, device="cuda:0"), term, {"old_logp": torch.zeros(4, device="cuda:0")})
        seen.append(obs.cpu())
        obs = nxt
    assert buf.full
    for t in range(8):
        assert torch.equal(buf.observations[:, t].cpu(), seen[t])
    envs.close()

"""GPU tests of the QMIX Agent path: action-masked selection (K9 select with ``avail``) vs the oracle's masked variant,
the multi-agent vector env contract, and QMIX learning end to end through ``REGISTRY_Agents["QMIX"]`` on the
SMAC-shaped synthetic environment (use_actions_mask=True, the shipped SC2 configuration)."""
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle.qmix import QMIXModelOracle, QMIXLearnerOracle
from test_gpu_qmix import _product_model

pytestmark = pytest.mark.gpu


def _random_sample(rng, keys, B, T, obs_dim, A, S):
    L = rng.integers(max(2, T // 3), T + 1, size=B)
    filled = (np.arange(T)[None, :] < L[:, None])
    avail = {k: (rng.random((B, T + 1, A)) < 0.6) for k in keys}
    acts = {}
    for k in keys:
        avail[k][..., 1] = True                                        # at least one available action everywhere
        logits = rng.random((B, T, A)) + 10.0 * avail[k][:, :T]
        acts[k] = logits.argmax(-1).astype(np.float32)                 # taken actions are available ones
    return {"obs": {k: rng.normal(size=(B, T + 1, obs_dim)).astype(np.float32) for k in keys},
            "actions": acts,
            "rewards": {k: rng.normal(size=(B, T)).astype(np.float32) for k in keys},
            "terminals": {k: (rng.random((B, T)) < 0.05) for k in keys},
            "agent_mask": {k: (rng.random((B, T)) < 0.9) for k in keys},
            "avail_actions": avail, "filled": filled, "state": rng.normal(size=(B, T + 1, S)).astype(np.float32),
            "batch_size": B, "sequence_length": T}


@pytest.mark.parametrize("double_q", [True, False])
def test_qmix_update_with_action_mask_matches_oracle(double_q):
    from xuance_b200.common import BaseCallback
    from xuance_b200.torch.learners.qmix_learner import QMIX_Learner
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    n, obs_dim, A, S, T, B = 5, 72, 12, 98, 20, 8
    keys, grouping, model = _product_model(n, obs_dim, A, S)
    om = QMIXModelOracle(n, obs_dim, A, S)
    model.load_state_dict(om.state_dict(), strict=True)
    cfg = Namespace(distributed_training=False, episode_length=T, use_grad_clip=True, grad_clip_norm=10.0,
                    device="cuda:0", model_dir="/tmp/x", running_steps=100000, parallels=4, use_parameter_sharing=True,
                    use_rnn=True, use_actions_mask=True, learning_rate=7e-4, sync_frequency=2, double_q=double_q,
                    n_epochs=1, start_training=0, gamma=0.99, end_factor_lr_decay=0.5, qmix_rnn_detach_q_eval=False,
                    use_cuda_graph=False)
    lrn = QMIX_Learner(cfg, grouping, model, BaseCallback())
    orc = QMIXLearnerOracle(om, keys, learning_rate=7e-4, sync_frequency=2, double_q=double_q, use_grad_clip=True,
                            grad_clip_norm=10.0, end_factor_lr_decay=0.5, total_iters=lrn.total_iters,
                            detach_q_eval=False, use_actions_mask=True)
    rng = np.random.default_rng(7)
    for it in range(3):
        s = _random_sample(rng, keys, B, T, obs_dim, A, S)
        ip, io = lrn.update(s), orc.update(s)
        # the masked target contains -1e10 * (agents whose every next action is unavailable never occur here) - values stay O(1)
        np.testing.assert_allclose(ip["loss_Q"], io["loss_Q"], rtol=5e-4, atol=1e-6, err_msg=f"it{it}")
        np.testing.assert_allclose(ip["predictQ"], io["predictQ"], rtol=5e-4, atol=1e-5)
    # and the mask matters: the same batch without it gives a different loss
    orc2 = QMIXLearnerOracle(QMIXModelOracle(n, obs_dim, A, S), keys, use_actions_mask=False, detach_q_eval=False)
    orc2.model.load_state_dict(om.state_dict())
    orc3 = QMIXLearnerOracle(QMIXModelOracle(n, obs_dim, A, S), keys, use_actions_mask=True, detach_q_eval=False)
    orc3.model.load_state_dict(om.state_dict())
    s = _random_sample(rng, keys, B, T, obs_dim, A, S)
    assert abs(orc2.update(s)["loss_Q"] - orc3.update(s)["loss_Q"]) > 1e-6


def _qmix_config(**over):
    cfg = dict(agent="QMIX", env_name="StarCraft2", env_id="5m_vs_6m", env_seed=1, learner="QMIX_Learner",
               policy="Mixing_Q_network", representation="Basic_RNN", vectorize="Dummy_StarCraft2", use_rnn=True, rnn="GRU",
               N_recurrent_layers=1, fc_hidden_sizes=[64], recurrent_hidden_size=64, dropout=0, q_hidden_size=[64],
               activation="relu", use_parameter_sharing=True, use_actions_mask=True, hidden_dim_mixing_net=32,
               hidden_dim_hyper_net=32, seed=1, parallels=8, buffer_size=64, batch_size=16, learning_rate=2e-3, gamma=0.5,
               double_q=True, start_greedy=1.0, end_greedy=0.05, decay_step_greedy=12000, start_training=200,
               running_steps=40000, n_epochs=4, sync_frequency=20, use_grad_clip=False, grad_clip_norm=0.5,
               device="cuda:0", model_dir="/tmp/xb200_qmix_models", log_dir="/tmp/xb200_qmix_logs", episode_limit=20,
               p_death=0.02, distributed_training=False, use_cuda_graph=False)
    cfg.update(over)
    return Namespace(**cfg)


def test_multi_agent_vector_env_contract():
    from xuance_b200.environment import make_envs, DummyVecMultiAgentEnv
    from xuance_b200.environment.vector_envs import AlreadySteppingError, NotSteppingError
    envs = make_envs(_qmix_config(parallels=3))
    assert isinstance(envs, DummyVecMultiAgentEnv) and envs.num_envs == 3 and envs.num_agents == 5
    assert envs.state_space.shape == (98,) and envs.observation_space["agent_0"].shape == (72,)
    obs, info = envs.reset()
    assert set(obs[0]) == set(envs.agents) and info[0]["episode_step"] == 0
    with pytest.raises(NotSteppingError):
        envs.step_wait()
    done = 0
    for t in range(25):
        acts = [{a: int(np.flatnonzero(envs.buf_avail_actions[e][a])[0]) for a in envs.agents} for e in range(3)]
        envs.step_async(acts)
        with pytest.raises(AlreadySteppingError):
            envs.step_async(acts)
        obs, rew, term, trunc, info = envs.step_wait()
        for e in range(3):
            assert np.array_equal(envs.buf_state[e], info[e]["state"])
            if all(term[e].values()) or trunc[e]:
                done += 1
                assert {"reset_obs", "reset_state", "reset_avail_actions"} <= set(info[e])
                assert info[e]["episode_step"] <= 20
    assert done >= 3
    envs.close()


def test_qmix_learns_through_the_agent_registry():
    """REGISTRY_Agents["QMIX"] on the cue-following SMAC-shaped env: the greedy team score rises well above chance
    (chance with ~7.7 available actions per agent is ~0.13 per step; the trained policy reads the cue)."""
    from xuance_b200.environment import make_envs
    from xuance_b200.torch.agents import REGISTRY_Agents
    cfg = _qmix_config()
    envs = make_envs(cfg)
    agent = REGISTRY_Agents[cfg.agent](cfg, envs)
    assert agent.use_actions_mask and agent.memory.use_actions_mask
    before = float(np.mean(agent.test(16, make_envs(cfg))))
    info = agent.train(cfg.running_steps // cfg.parallels)
    assert "loss_Q" in info and np.isfinite(info["loss_Q"])
    assert agent.memory.size > 0 and agent.e_greedy < 0.2
    after = float(np.mean(agent.test(32, make_envs(cfg))))
    # episode score = sum over <=20 steps of the fraction of living agents that hit their cue
    assert after > before + 4.0 and after > 8.0, (before, after)
    agent.finish()


def test_vdn_agent_trains_through_the_registry():
    """REGISTRY_Agents["VDN"]: the sum mixer (no parameters) through the same rollout / replay / learner path."""
    from xuance_b200.environment import make_envs
    from xuance_b200.torch.agents import REGISTRY_Agents
    from xuance_b200.torch.rl_models import VDN_mixer
    cfg = _qmix_config(agent="VDN", learner="VDN_Learner", running_steps=20000, decay_step_greedy=6000)
    agent = REGISTRY_Agents["VDN"](cfg, make_envs(cfg))
    assert isinstance(agent.model.eval_Qtot, VDN_mixer) and not list(agent.model.eval_Qtot.parameters())
    before = float(np.mean(agent.test(16, make_envs(cfg))))
    info = agent.train(cfg.running_steps // cfg.parallels)
    assert np.isfinite(info["loss_Q"])
    after = float(np.mean(agent.test(32, make_envs(cfg))))
    assert after > before + 2.0, (before, after)
    agent.finish()

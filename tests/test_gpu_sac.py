"""GPU parity: SAC learner (K8 loss kernels, three flat-bucket Adam steps, soft update) vs the torch-CPU oracle."""
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle.sac import SACModelOracle, SACLearnerOracle

pytestmark = pytest.mark.gpu


def _build(obs_dim, act_dim, device):
    from xuance_b200.common import Box
    from xuance_b200.torch.rl_models import Basic_Identical, SAC_GaussianActor, TwinActionValueCritic, SoftActorCritic
    from copy import deepcopy
    aspace = Box(-1, 1, (act_dim,), np.float32)
    rep = Basic_Identical((obs_dim,), device=device)
    actor = SAC_GaussianActor(rep, [256, 256], aspace, None, None, nn.LeakyReLU, nn.Tanh, device)
    critic = TwinActionValueCritic(deepcopy(rep), aspace, [256, 256], None, None, nn.LeakyReLU, device)
    return SoftActorCritic(actor, critic).to(device)


@pytest.mark.parametrize("auto_alpha,graph", [(True, False), (False, False), (True, True)])
def test_sac_learner_matches_oracle(auto_alpha, graph):
    from xuance_b200.common import BaseCallback
    from xuance_b200.torch.learners.sac_learner import SAC_Learner
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    obs_dim, act_dim, B = 17, 6, 256
    om = SACModelOracle(obs_dim, act_dim)
    model = _build(obs_dim, act_dim, "cuda:0")
    missing = model.load_state_dict(om.state_dict(), strict=True)
    cfg = Namespace(distributed_training=False, episode_length=1000, use_grad_clip=False, grad_clip_norm=0.5,
                    device="cuda:0", model_dir="/tmp/xb", running_steps=100000, parallels=4, start_training=0,
                    training_frequency=1, learning_rate_actor=1e-3, learning_rate_critic=1e-3, tau=0.005, gamma=0.99,
                    alpha=0.2, use_automatic_entropy_tuning=auto_alpha, end_factor_lr_decay=0.7, use_cuda_graph=graph)
    lrn = SAC_Learner(cfg, model, BaseCallback())
    orc = SACLearnerOracle(om, auto_alpha=auto_alpha, end_factor_lr_decay=0.7, total_iters=lrn.total_iters)
    rng = np.random.default_rng(1)
    for it in range(4):
        s = {"obs": rng.normal(size=(B, obs_dim)).astype(np.float32),
             "actions": rng.uniform(-1, 1, size=(B, act_dim)).astype(np.float32),
             "obs_next": rng.normal(size=(B, obs_dim)).astype(np.float32),
             "rewards": rng.normal(size=B).astype(np.float32), "terminals": (rng.random(B) < 0.1).astype(np.float32)}
        n1 = torch.from_numpy(rng.normal(size=(B, act_dim)).astype(np.float32))
        n2 = torch.from_numpy(rng.normal(size=(B, act_dim)).astype(np.float32))
        io = orc.update(n1, n2, **s)
        ip = lrn.update(noise_pi=n1.cuda(), noise_next=n2.cuda(), **{k: torch.from_numpy(v).cuda() for k, v in s.items()})
        for k in ("Qloss", "Ploss", "Qvalue"):
            np.testing.assert_allclose(ip[k], io[k], rtol=2e-4, atol=2e-5, err_msg=f"{k} it{it}")
        assert ip["actor_lr"] == io["actor_lr"] and ip["critic_lr"] == io["critic_lr"]
        if auto_alpha:
            np.testing.assert_allclose(ip["alpha"], io["alpha"], rtol=1e-5)
            np.testing.assert_allclose(ip["alpha_loss"], io["alpha_loss"], rtol=1e-3, atol=1e-6)
    so, sp = om.state_dict(), model.state_dict()
    for k in so:
        np.testing.assert_allclose(sp[k].cpu().numpy(), so[k].numpy(), rtol=2e-3, atol=2e-4, err_msg=k)
    # target critic moved by the Polyak update and still differs from the critic
    assert not torch.equal(sp["target_critic.critic_head_1.values.0.weight"], sp["critic.critic_head_1.values.0.weight"])

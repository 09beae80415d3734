"""Torch-CPU restatement of the reference learners' update() on the hot path.  TEST INFRASTRUCTURE.

  * PPO-Clip ........ xuance/torch/learners/policy_gradient/ppo_learner.py:13-95
  * DQN / PER-DQN ... xuance/torch/learners/qlearning_family/dqn_learner.py:13-75, perdqn_learner.py:16-80
  * optimiser recipe  Adam(lr, eps=1e-5) + LinearLR(1 -> end_factor over total_iters) stepped every update,
                      clip_grad_norm_(max_norm) before the step  (ppo_learner.py:18-22,61-67)
The reference's arithmetic here is PyTorch autograd + torch.optim, so this uses the same torch calls; what is
restated is the loss algebra, its operand order, the returned info keys and the optimiser recipe.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.distributions import Categorical


def ppo_clip_terms(logits, v_pred, act, ret, adv, old_logp, clip_range):
    """ppo_learner.py:46-59 - returns (a_loss, c_loss, e_loss, ratio)."""
    dist = Categorical(logits=logits)
    log_prob = dist.log_prob(act)
    ratio = (log_prob - old_logp).exp().float()
    surrogate1 = ratio.clamp(1.0 - clip_range, 1.0 + clip_range) * adv
    surrogate2 = adv * ratio
    a_loss = -torch.minimum(surrogate1, surrogate2).mean()
    c_loss = nn.functional.mse_loss(v_pred, ret.detach())
    e_loss = dist.entropy().mean()
    return a_loss, c_loss, e_loss, ratio


class PPOLearnerOracle:
    kind = "ppo"   # "a2c": a2c_learner.py:46-52 ; "pg": pg_learner.py:44-47

    def __init__(self, model, learning_rate=2.5e-4, vf_coef=0.25, ent_coef=0.01, clip_range=0.2,
                 use_grad_clip=True, grad_clip_norm=0.5, end_factor_lr_decay=1.0, total_iters=1, kind="ppo"):
        self.kind = kind
        self.model = model
        self.optimizer = torch.optim.Adam(model.parameters(), learning_rate, eps=1e-5)
        self.scheduler = torch.optim.lr_scheduler.LinearLR(self.optimizer, start_factor=1.0,
                                                           end_factor=end_factor_lr_decay, total_iters=total_iters)
        self.vf_coef, self.ent_coef, self.clip_range = vf_coef, ent_coef, clip_range
        self.use_grad_clip, self.grad_clip_norm = use_grad_clip, grad_clip_norm
        self.iterations = 0

    def update(self, **samples):
        self.iterations += 1
        obs = torch.as_tensor(samples['obs'])
        act = torch.as_tensor(samples['actions'])
        ret = torch.as_tensor(samples['returns'])
        adv = torch.as_tensor(samples['advantages'])
        old_logp = torch.as_tensor(samples['aux_batch']['old_logp']) if self.kind == "ppo" else None
        logits, v_pred = self.model(obs)
        if self.kind == "ppo":
            a_loss, c_loss, e_loss, ratio = ppo_clip_terms(logits, v_pred, act, ret, adv, old_logp, self.clip_range)
            loss = a_loss - self.ent_coef * e_loss + self.vf_coef * c_loss
        else:
            dist = Categorical(logits=logits)
            log_prob = dist.log_prob(act)
            e_loss = dist.entropy().mean()
            if self.kind == "a2c":
                a_loss = -(adv * log_prob).mean()
                c_loss = nn.functional.mse_loss(v_pred, ret)
                loss = a_loss - self.ent_coef * e_loss + self.vf_coef * c_loss
            else:
                a_loss = -(ret * log_prob).mean()
                c_loss = torch.zeros(())
                loss = a_loss - self.ent_coef * e_loss
            ratio = torch.ones_like(log_prob)
        self.optimizer.zero_grad()
        loss.backward()
        if self.use_grad_clip:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clip_norm)
        self.optimizer.step()
        self.scheduler.step()
        lr = self.optimizer.state_dict()['param_groups'][0]['lr']
        cr = ((ratio < 1 - self.clip_range).sum() + (ratio > 1 + self.clip_range).sum()) / ratio.shape[0]
        return {"actor_loss": a_loss.item(), "critic_loss": c_loss.item(), "entropy": e_loss.item(),
                "learning_rate": lr, "predict_value": v_pred.mean().item(), "clip_ratio": cr}


class DQNLearnerOracle:
    """DQN_Learner / PerDQN_Learner (``per=True`` also returns |td| as float32 ndarray, perdqn_learner.py:80)."""

    def __init__(self, model, learning_rate=1e-4, gamma=0.99, sync_frequency=500, use_grad_clip=False,
                 grad_clip_norm=0.5, end_factor_lr_decay=1.0, total_iters=1, per=False, double_q=False):
        self.double_q = double_q   # ddqn_learner.py:39-44
        self.model = model
        self.optimizer = torch.optim.Adam(model.parameters(), learning_rate, eps=1e-5)
        self.scheduler = torch.optim.lr_scheduler.LinearLR(self.optimizer, start_factor=1.0,
                                                           end_factor=end_factor_lr_decay, total_iters=total_iters)
        self.gamma, self.sync_frequency = gamma, sync_frequency
        self.use_grad_clip, self.grad_clip_norm = use_grad_clip, grad_clip_norm
        self.per = per
        self.iterations = 0

    def update(self, **samples):
        self.iterations += 1
        obs = torch.as_tensor(samples['obs'])
        act = torch.as_tensor(samples['actions'], dtype=torch.int64)
        nxt = torch.as_tensor(samples['obs_next'])
        rew = torch.as_tensor(samples['rewards'])
        ter = torch.as_tensor(samples['terminals'], dtype=torch.float)
        evalQ = self.model(obs)
        targetQ = self.model.target(nxt)
        predictQ = evalQ.gather(-1, act.unsqueeze(-1)).squeeze(-1)
        if self.double_q:
            targetA = self.model(nxt).argmax(dim=-1)
            targetQ = targetQ.gather(-1, targetA.unsqueeze(-1)).squeeze(-1)
        else:
            targetQ = targetQ.max(dim=-1).values
        targetQ = rew + self.gamma * (1 - ter) * targetQ
        td_error = targetQ - predictQ
        loss = nn.functional.mse_loss(predictQ, targetQ.detach())
        self.optimizer.zero_grad()
        loss.backward()
        if self.use_grad_clip:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clip_norm)
        self.optimizer.step()
        self.scheduler.step()
        if self.iterations % self.sync_frequency == 0:
            self.model.copy_target()
        lr = self.optimizer.state_dict()['param_groups'][0]['lr']
        info = {"Qloss": loss.item(), "learning_rate": lr, "predictQ": predictQ.mean().item()}
        if self.per:
            return np.abs(td_error.cpu().detach().numpy()), info
        return info

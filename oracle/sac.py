"""Torch-CPU restatement of the reference SAC model + learner (BASELINE config 4).  TEST INFRASTRUCTURE.

  * Basic_Identical representation ............. xuance/torch/rl_models/representations/mlp.py:10-24
  * SAC_GaussianActorHead (clamp log_std) ...... heads/actor_head.py:75-105
  * tanh-squashed Gaussian log-prob ............ modules/distributions.py:200-218
  * TwinActionValueCritic / ValueHead .......... critics/twin_critics.py:11-62, heads/critic_head.py:9-30
  * SoftActorCritic (Qpolicy/Qtarget/Qaction/soft_update) ... architectures/single_agent/actor_critic.py:107-159
  * SAC_Learner.update ......................... learners/policy_gradient/sac_learner.py:14-126
The reference draws the reparameterisation noise with Normal.rsample(); for reproducible parity the restatement
takes the two standard-normal noise tensors (actor step, target step) explicitly - mu + std*noise is exactly what
rsample computes."""
import copy
import numpy as np
import torch
import torch.nn as nn
from torch.distributions import Normal
from torch.nn.functional import softplus

from .nets import mlp


class SACModelOracle(nn.Module):
    def __init__(self, obs_dim, act_dim, hidden=(256, 256), act=nn.LeakyReLU, init=None):
        super().__init__()
        self.actor = nn.Module()
        self.actor.actor_head = nn.Module()
        self.actor.actor_head.output = mlp([obs_dim, *hidden], act=act, last_act=True, init=init)
        self.actor.actor_head.out_mu = nn.Linear(hidden[-1], act_dim)
        self.actor.actor_head.out_log_std = nn.Linear(hidden[-1], act_dim)
        self.critic = nn.Module()
        self.critic.critic_head_1 = nn.Module()
        self.critic.critic_head_1.values = mlp([obs_dim + act_dim, *hidden, 1], act=act, init=init)
        self.critic.critic_head_2 = nn.Module()
        self.critic.critic_head_2.values = mlp([obs_dim + act_dim, *hidden, 1], act=act, init=init)
        self.target_critic = copy.deepcopy(self.critic)
        self.act_dim = act_dim

    def _dist(self, obs):
        h = self.actor.actor_head.output(obs)
        mu = self.actor.actor_head.out_mu(h)
        log_std = torch.clamp(self.actor.actor_head.out_log_std(h), -20, 2)
        return mu, log_std.exp()

    def sample_and_logprob(self, obs, noise):
        mu, std = self._dist(obs)
        pre = mu + std * noise
        act = torch.tanh(pre)
        log_prob = Normal(mu, std).log_prob(pre)
        log_prob = log_prob + (-2. * (torch.log(torch.tensor([2.0])) - pre - softplus(-2. * pre)))
        return act, log_prob.sum(-1)

    @staticmethod
    def _q(critic, obs, act):
        x = torch.concat([obs, act], dim=-1)
        return critic.critic_head_1.values(x).squeeze(-1), critic.critic_head_2.values(x).squeeze(-1)

    def Qpolicy(self, obs, noise):
        a, lp = self.sample_and_logprob(obs, noise)
        q1, q2 = self._q(self.critic, obs, a)
        return lp, q1, q2

    def Qtarget(self, obs, noise):
        a, lp = self.sample_and_logprob(obs, noise)
        q1, q2 = self._q(self.target_critic, obs, a)
        return lp, torch.min(q1, q2)

    def Qaction(self, obs, act):
        return self._q(self.critic, obs, act)

    def soft_update(self, tau):
        for ep, tp in zip(self.critic.parameters(), self.target_critic.parameters()):
            tp.data.mul_(1 - tau)
            tp.data.add_(tau * ep.data)


class SACLearnerOracle:
    def __init__(self, model, lr_actor=1e-3, lr_critic=1e-3, tau=0.005, gamma=0.99, alpha=0.2, auto_alpha=True,
                 use_grad_clip=False, grad_clip_norm=0.5, end_factor_lr_decay=1.0, total_iters=1):
        self.model = model
        self.opt_a = torch.optim.Adam(model.actor.parameters(), lr_actor)
        self.opt_c = torch.optim.Adam(model.critic.parameters(), lr_critic)
        mk = lambda o: torch.optim.lr_scheduler.LinearLR(o, start_factor=1.0, end_factor=end_factor_lr_decay,
                                                         total_iters=total_iters)
        self.sch_a, self.sch_c = mk(self.opt_a), mk(self.opt_c)
        self.tau, self.gamma, self.alpha = tau, gamma, alpha
        self.auto = auto_alpha
        self.use_grad_clip, self.grad_clip_norm = use_grad_clip, grad_clip_norm
        if auto_alpha:
            self.target_entropy = -float(model.act_dim)
            self.log_alpha = nn.Parameter(torch.zeros(1, requires_grad=True))
            self.alpha = self.log_alpha.exp()
            self.opt_alpha = torch.optim.Adam([self.log_alpha], lr=lr_actor)

    def update(self, noise_pi, noise_next, **s):
        obs, act = torch.as_tensor(s['obs']), torch.as_tensor(s['actions'])
        nxt, rew = torch.as_tensor(s['obs_next']), torch.as_tensor(s['rewards'])
        ter = torch.as_tensor(s['terminals'], dtype=torch.float)
        log_pi, q1, q2 = self.model.Qpolicy(obs, noise_pi)
        model_q = torch.min(q1, q2).reshape([-1])
        p_loss = (self.alpha * log_pi.reshape([-1]) - model_q).mean()
        self.opt_a.zero_grad()
        p_loss.backward()
        if self.use_grad_clip:
            torch.nn.utils.clip_grad_norm_(self.model.actor.parameters(), self.grad_clip_norm)
        self.opt_a.step()
        aq1, aq2 = self.model.Qaction(obs, act)
        log_pi_next, target_q = self.model.Qtarget(nxt, noise_next)
        target_value = target_q - self.alpha * log_pi_next.reshape([-1])
        backup = rew + (1 - ter) * self.gamma * target_value
        q_loss = nn.functional.mse_loss(aq1, backup.detach()) + nn.functional.mse_loss(aq2, backup.detach())
        self.opt_c.zero_grad()
        q_loss.backward()
        if self.use_grad_clip:
            torch.nn.utils.clip_grad_norm_(self.model.critic.parameters(), self.grad_clip_norm)
        self.opt_c.step()
        info = {"Qloss": q_loss.item(), "Ploss": p_loss.item(), "Qvalue": model_q.mean().item()}
        if self.auto:
            alpha_loss = -(self.log_alpha * (log_pi + self.target_entropy).detach()).mean()
            self.opt_alpha.zero_grad()
            alpha_loss.backward()
            self.opt_alpha.step()
            self.alpha = self.log_alpha.exp()
            info.update(alpha_loss=alpha_loss.item(), alpha=self.alpha.item())
        self.sch_a.step()
        self.sch_c.step()
        self.model.soft_update(self.tau)
        info.update(actor_lr=self.opt_a.param_groups[0]['lr'], critic_lr=self.opt_c.param_groups[0]['lr'])
        return info

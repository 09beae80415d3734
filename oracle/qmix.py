"""Torch-CPU restatement of the reference QMIX path (BASELINE config 5): episode replay, GRU agents, hypernetwork
mixer, QMIX_Learner.update.  TEST INFRASTRUCTURE (see oracle/__init__).

  * MARL_OffPolicyBuffer_RNN (episode-major store/finish_path/sample) ... xuance/common/memory_tools_marl.py:770-996
  * Basic_RNN (Linear+ReLU -> GRU, batch_first) .......................... torch/rl_models/representations/rnn.py:9-99
  * DiscreteActionValueCritic / QValueHead ............................... critics/base_critics.py:91-135, heads/q_head.py:11-39
  * QMIX_Mixer ........................................................... heads/q_mix_head.py:28-95
  * MixingQNetwork forward / Qtarget / Q_tot / copy_target ............... architectures/multi_agent/value_factorization.py:17-174
  * build_training_data + _forward_transitions + QMIX_Learner.update ..... learners/base/marl_learner.py:319-408,
                                                 learners/multi_agent_rl/iql_learner.py:37-83, qmix_learner.py:24-112
Scope: one parameter-sharing group ("shared"), use_rnn=True; use_actions_mask=False as the reference runs it, and the masked
variant the reference evidently intends (its RNN+mask path crashes on one slice, SURVEY.md headline 6), pinned against the
live reference with that slice corrected at test time.  Module / parameter names mirror the reference so its state_dict
loads 1:1."""
import copy
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .nets import mlp


class EpisodeReplayOracle:
    """memory_tools_marl.py:770-996 for a fixed agent list; arrays keyed like the reference's ``data`` dict."""

    def __init__(self, agent_keys, obs_dim, state_dim, n_envs, buffer_size, batch_size, max_episode_steps):
        self.agent_keys, self.n_envs = list(agent_keys), n_envs
        self.buffer_size, self.batch_size, self.T = buffer_size, batch_size, max_episode_steps
        self.obs_dim, self.state_dim = obs_dim, state_dim
        self.data = self._alloc(buffer_size)
        self.episode_data = self._alloc(n_envs)
        self.ptr = self.size = 0

    def _alloc(self, n):
        T, K = self.T, self.agent_keys
        return {'obs': {k: np.zeros((n, T + 1, self.obs_dim), np.float32) for k in K},
                'actions': {k: np.zeros((n, T), np.float32) for k in K},
                'rewards': {k: np.zeros((n, T), np.float32) for k in K},
                'terminals': {k: np.zeros((n, T), np.bool_) for k in K},
                'agent_mask': {k: np.zeros((n, T), np.bool_) for k in K},
                'filled': np.zeros((n, T), np.bool_),
                'state': np.zeros((n, T + 1, self.state_dim), np.float32)}

    def store(self, **step):  # :912-929
        t = step['episode_steps']
        e = range(self.n_envs)
        self.episode_data['filled'][e, t] = True
        for key, val in step.items():
            if key not in self.episode_data:
                continue
            if key == 'state':
                self.episode_data['state'][e, t] = val
                continue
            for a in self.agent_keys:
                self.episode_data[key][a][e, t] = val[a]

    def finish_path(self, i_env, **terminal):  # :952-969 + store_episodes :931-950
        t = terminal['episode_step']
        self.episode_data['state'][i_env, t] = terminal['state']
        for a in self.agent_keys:
            self.episode_data['obs'][a][i_env, t] = terminal['obs'][a]
        for key in self.data:
            if key in ('filled', 'state'):
                self.data[key][self.ptr] = self.episode_data[key][i_env].copy()
            else:
                for a in self.agent_keys:
                    self.data[key][a][self.ptr] = self.episode_data[key][a][i_env].copy()
        self.ptr = (self.ptr + 1) % self.buffer_size
        self.size = min(self.size + 1, self.buffer_size)
        self.episode_data['filled'][i_env] = np.zeros(self.T, np.bool_)

    def gather(self, episodes):
        out = {}
        for key in self.data:
            if key in ('filled', 'state'):
                out[key] = self.data[key][episodes]
            else:
                out[key] = {a: self.data[key][a][episodes] for a in self.agent_keys}
        out['batch_size'], out['sequence_length'] = len(episodes), self.T
        return out

    def sample(self, batch_size=None):  # :971-996
        bs = self.batch_size if batch_size is None else batch_size
        return self.gather(np.random.choice(self.size, bs))


class _RNNRep(nn.Module):
    def __init__(self, obs_dim, fc_hidden, rnn_hidden, init=nn.init.orthogonal_):
        super().__init__()
        self.mlp = mlp([obs_dim, fc_hidden], last_act=True, init=init)
        self.rnn = nn.GRU(fc_hidden, rnn_hidden, 1, batch_first=True)
        if init is not None:
            for wl in self.rnn.all_weights:
                for w in wl:
                    init(w) if len(w.shape) > 1 else nn.init.constant_(w, 0)

    def forward(self, x, h0):
        return self.rnn(self.mlp(x), h0)[0]


class _AgentQ(nn.Module):
    def __init__(self, obs_dim, n_actions, fc_hidden=64, rnn_hidden=64, q_hidden=(64,), init=nn.init.orthogonal_):
        super().__init__()
        self.representation = nn.Module()
        self.representation.obs_representation = _RNNRep(obs_dim, fc_hidden, rnn_hidden, init)
        self.critic_head = nn.Module()
        self.critic_head.q_value = mlp([rnn_hidden, *q_hidden, n_actions], init=init)
        self.rnn_hidden = rnn_hidden

    def forward(self, obs):  # obs [B*n, T+1, obs_dim]
        h0 = torch.zeros(1, obs.shape[0], self.rnn_hidden)
        return self.critic_head.q_value(self.representation.obs_representation(obs, h0))


class MixerOracle(nn.Module):
    """q_mix_head.py:28-95."""

    def __init__(self, dim_state, dim_hidden, dim_hypernet_hidden, n_agents):
        super().__init__()
        self.S, self.H, self.n = dim_state, dim_hidden, n_agents
        hh = dim_hypernet_hidden
        self.hyper_w_1 = nn.Sequential(nn.Linear(dim_state, hh), nn.ReLU(), nn.Linear(hh, dim_hidden * n_agents))
        self.hyper_w_2 = nn.Sequential(nn.Linear(dim_state, hh), nn.ReLU(), nn.Linear(hh, dim_hidden))
        self.hyper_b_1 = nn.Linear(dim_state, dim_hidden)
        self.hyper_b_2 = nn.Sequential(nn.Linear(dim_state, hh), nn.ReLU(), nn.Linear(hh, 1))

    def forward(self, values_n, states):
        states = torch.as_tensor(states, dtype=torch.float32).reshape(-1, self.S)
        agent_qs = values_n.reshape(-1, 1, self.n)
        w_1 = torch.abs(self.hyper_w_1(states)).view(-1, self.n, self.H)
        b_1 = self.hyper_b_1(states).view(-1, 1, self.H)
        hidden = F.elu(torch.bmm(agent_qs, w_1) + b_1)
        w_2 = torch.abs(self.hyper_w_2(states)).view(-1, self.H, 1)
        b_2 = self.hyper_b_2(states).view(-1, 1, 1)
        return (torch.bmm(hidden, w_2) + b_2).view(-1, 1)


class QMIXModelOracle(nn.Module):
    def __init__(self, n_agents, obs_dim, n_actions, state_dim, fc_hidden=64, rnn_hidden=64, q_hidden=(64,),
                 mix_hidden=32, hyper_hidden=32):
        super().__init__()
        self.n_agents = n_agents
        self.individual_q_networks = nn.ModuleDict({'shared': _AgentQ(obs_dim, n_actions, fc_hidden, rnn_hidden, q_hidden)})
        self.target_individual_q_networks = copy.deepcopy(self.individual_q_networks)
        self.eval_Qtot = MixerOracle(state_dim, mix_hidden, hyper_hidden, n_agents)
        self.target_Qtot = copy.deepcopy(self.eval_Qtot)

    def parameters_model(self):
        return list(self.individual_q_networks.parameters()) + list(self.eval_Qtot.parameters())

    def copy_target(self):
        for e, t in zip(self.individual_q_networks.parameters(), self.target_individual_q_networks.parameters()):
            t.data.copy_(e)
        for e, t in zip(self.eval_Qtot.parameters(), self.target_Qtot.parameters()):
            t.data.copy_(e)


class QMIXLearnerOracle:
    def __init__(self, model, agent_keys, learning_rate=7e-4, gamma=0.99, sync_frequency=200, double_q=True,
                 use_grad_clip=False, grad_clip_norm=10.0, end_factor_lr_decay=1.0, total_iters=1,
                 detach_q_eval=True, use_actions_mask=False):
        """``use_actions_mask=True`` is the INTENDED masking of iql_learner.py:60-81 / value_factorization.py:86-89 for
        use_rnn (the double-Q arg-max sees unavailable actions at -1e10; target values of unavailable actions at step t+1
        become -1e10) with the time axis sliced as ``[:, :, 1:]``.  The reference at 4f0b05b slices the AGENT axis there
        (``[:, 1:]``) and raises a shape error, so this variant cannot be pinned to the reference as it is; it is pinned to
        the live reference learner with that ONE expression corrected in memory at test time
        (tests/test_oracle_vs_reference.py::test_qmix_action_mask_variant_against_the_reference_with_its_slice_corrected).
        ``detach_q_eval=True`` reproduces the reference AS IS: iql_learner.py:57-59 slices ``q_eval`` to
        ``[:, :, :-1]`` INSIDE ``torch.no_grad()``, so with use_rnn=True the sliced tensor carries no graph and the
        agent networks receive no gradient - only the mixer trains (verified against the live reference:
        tests/test_oracle_vs_reference.py).  ``False`` is the evidently intended computation (slice outside no_grad)."""
        self.detach_q_eval = detach_q_eval
        self.use_actions_mask = use_actions_mask
        self.model, self.agent_keys = model, list(agent_keys)
        # LearnerMAS.build_optimizer (marl_learner.py:64-76): ONE Adam over model.parameters() (targets get no grads)
        self.optimizer = torch.optim.Adam(model.parameters(), lr=learning_rate, eps=1e-5, weight_decay=0.0)
        self.scheduler = torch.optim.lr_scheduler.LinearLR(self.optimizer, start_factor=1.0,
                                                           end_factor=end_factor_lr_decay, total_iters=total_iters)
        self.gamma, self.sync_frequency, self.double_q = gamma, sync_frequency, double_q
        self.use_grad_clip, self.grad_clip_norm = use_grad_clip, grad_clip_norm
        self.iterations = 0

    def update(self, sample):
        self.iterations += 1
        K, B, T = self.agent_keys, sample['batch_size'], sample['sequence_length']
        n = len(K)
        st = lambda d, dt=None: torch.stack([torch.as_tensor(d[k], dtype=dt) for k in K], dim=1)
        obs = st(sample['obs'])                                    # [B, n, T+1, obs]
        actions = st(sample['actions'])                            # [B, n, T]
        rewards, terminals = st(sample['rewards']), st(sample['terminals'], torch.float32)
        agent_mask = st(sample['agent_mask'], torch.float32)
        filled = torch.as_tensor(sample['filled'], dtype=torch.float32)       # [B, T]
        state = torch.as_tensor(sample['state'])                   # [B, T+1, S]
        rewards_tot = rewards.mean(dim=1)
        terminals_tot = terminals.all(dim=1).float() if terminals.dtype == torch.bool else (terminals != 0).all(dim=1).float()
        packed = obs.flatten(0, 1)                                 # [B*n, T+1, obs]
        q_all = self.model.individual_q_networks['shared'](packed).reshape(B, n, T + 1, -1)
        avail = None
        if self.use_actions_mask:
            avail = st(sample['avail_actions'], torch.float32)     # [B, n, T+1, A]
        with torch.no_grad():
            if avail is not None:                                  # value_factorization.py:86-89
                q_det = q_all.clone().detach()
                q_det[avail == 0] = -1e10
                actions_next = q_det.argmax(dim=-1)[:, :, 1:]
            else:
                actions_next = q_all.argmax(dim=-1)[:, :, 1:]
            q_next = self.model.target_individual_q_networks['shared'](packed).reshape(B, n, T + 1, -1)[:, :, 1:]
            if avail is not None:                                  # iql_learner.py:75-81, time axis
                q_next = q_next.clone()
                q_next[avail[:, :, 1:] == 0] = -1e10
        q_eval = q_all[:, :, :-1]
        if self.detach_q_eval:
            q_eval = q_eval.detach()
        mask = agent_mask * filled.unsqueeze(1).repeat(1, n, 1)
        q_eval_taken = q_eval.gather(-1, actions.long().unsqueeze(-1)).reshape(B, n, T)
        if self.double_q:
            q_next_taken = q_next.gather(-1, actions_next.long().unsqueeze(-1)).reshape(B, n, T)
        else:
            q_next_taken = q_next.max(dim=-1, keepdim=True).values.reshape(B, n, T)
        q_eval_taken = q_eval_taken * mask
        q_next_taken = q_next_taken * mask
        cat = lambda x: torch.concat([x[:, i].reshape(-1, 1) for i in range(n)], dim=-1)     # Q_tot :137-142
        q_tot_eval = self.model.eval_Qtot(cat(q_eval_taken), state[:, :-1]).reshape(-1)
        q_tot_next = self.model.target_Qtot(cat(q_next_taken), state[:, 1:]).reshape(-1)
        q_tot_target = rewards_tot.reshape(-1) + (1 - terminals_tot.reshape(-1)) * self.gamma * q_tot_next
        f = filled.reshape(-1)
        td = (q_tot_eval - q_tot_target.detach()) * f
        loss = (td ** 2).sum() / f.sum()
        self.optimizer.zero_grad()
        loss.backward()
        if self.use_grad_clip:
            torch.nn.utils.clip_grad_norm_(self.model.parameters_model(), self.grad_clip_norm)
        self.optimizer.step()
        self.scheduler.step()
        info = {"learning_rate": self.optimizer.param_groups[0]['lr'], "loss_Q": loss.item(),
                "predictQ": q_tot_eval.mean().item()}
        if self.iterations % self.sync_frequency == 0:
            self.model.copy_target()
        return info

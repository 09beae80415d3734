"""Torch-CPU restatement of the reference's networks on the hot path.  TEST INFRASTRUCTURE (see oracle/__init__).

The reference's arithmetic here IS PyTorch (SURVEY.md 8c "third-party arithmetic"): nn.Conv2d / nn.Linear /
torch.distributions.Categorical.  The restatement therefore uses the same torch modules, wired as the reference
wires them, with parameter names chosen so that a reference ``state_dict`` loads 1:1 (checked in tests).

  * conv / linear blocks, padding = (k - s)//2 ........... xuance/torch/rl_models/modules/layers.py:16-65
  * AC_CNN_Atari (NatureCNN w/ padding -> 6400 flat) ....... representations/cnn.py:53-102
  * Basic_CNN (+AdaptiveMaxPool -> 64 feat) ................ representations/cnn.py:11-50
  * CategoricalActorHead / ValueHead / QValueHead .......... heads/actor_head.py:14-42, critic_head.py:9-30, q_head.py:11-39
  * SharedActorCritic ...................................... architectures/single_agent/actor_critic.py:40-59
  * DeepQNetwork (eval + deep-copied target) ............... architectures/single_agent/deep_q_network.py:19-99
"""
import copy
import numpy as np
import torch
import torch.nn as nn


def _ortho(layer, gain=float(np.sqrt(2.0))):
    nn.init.orthogonal_(layer.weight, gain=gain)
    nn.init.constant_(layer.bias, 0.0)
    return layer


def conv_stack(in_hwc, kernels, strides, filters, init=None):
    """layers.py:33-65: Conv2d(C->f, k, s, padding=(k-s)//2) + ReLU per stage; returns (modules, (C,H,W))."""
    h, w, c = in_hwc
    mods = []
    for k, s, f in zip(kernels, strides, filters):
        pad = int((k - s) // 2)
        conv = nn.Conv2d(c, f, k, s, padding=pad)
        if init is not None:
            init(conv)
        mods += [conv, nn.ReLU()]
        c = f
        h = int((h + 2 * pad - (k - 1) - 1) / s + 1)
        w = int((w + 2 * pad - (k - 1) - 1) / s + 1)
    return mods, (c, h, w)


def prep_pixels(obs):
    """cnn.py:98-101: obs/255.0 (true division) -> float32 -> NHWC->NCHW."""
    obs = torch.as_tensor(obs)
    obs = obs / 255.0
    return torch.as_tensor(obs, dtype=torch.float32).permute((0, 3, 1, 2))


class NatureEncoderAC(nn.Module):
    """AC_CNN_Atari: 3 conv(+ReLU), Flatten, Linear(6400->512)+ReLU, orthogonal(sqrt2)/zero-bias init."""

    def __init__(self, in_hwc=(84, 84, 4), kernels=(8, 4, 3), strides=(4, 2, 1), filters=(32, 64, 64),
                 fc_hidden_sizes=(512,)):
        super().__init__()
        mods, chw = conv_stack(in_hwc, kernels, strides, filters, init=_ortho)
        mods.append(nn.Flatten())
        d = int(np.prod(chw))
        for hsz in fc_hidden_sizes:
            mods += [_ortho(nn.Linear(d, hsz)), nn.ReLU()]
            d = hsz
        self.model = nn.Sequential(*mods)
        self.out_dim = d

    def forward(self, obs):
        return self.model(prep_pixels(obs))


class PoolEncoderQ(nn.Module):
    """Basic_CNN: convs (default torch init unless ``initialize``), AdaptiveMaxPool2d(1,1), Flatten."""

    def __init__(self, in_hwc=(84, 84, 4), kernels=(8, 4, 3), strides=(4, 2, 1), filters=(32, 64, 64)):
        super().__init__()
        mods, chw = conv_stack(in_hwc, kernels, strides, filters)
        mods += [nn.AdaptiveMaxPool2d((1, 1)), nn.Flatten()]
        self.model = nn.Sequential(*mods)
        self.out_dim = filters[-1]

    def forward(self, obs):
        return self.model(prep_pixels(obs))


def mlp(sizes, act=nn.ReLU, last_act=False, init=None):
    mods = []
    for i in range(len(sizes) - 1):
        lin = nn.Linear(sizes[i], sizes[i + 1])
        if init is not None:
            init(lin.weight)
            nn.init.constant_(lin.bias, 0)
        mods.append(lin)
        if i < len(sizes) - 2 or last_act:
            mods.append(act())
    return nn.Sequential(*mods)


class SharedActorCriticOracle(nn.Module):
    """representation -> {Linear(512->A) logits, Linear(512->1).squeeze(-1)}; parameter names mirror the
    reference (representation.model.*, actor.logits.*, critic.values.*) so state_dicts interchange."""

    def __init__(self, n_actions, init_heads=nn.init.orthogonal_, **enc_kw):
        super().__init__()
        self.representation = NatureEncoderAC(**enc_kw)
        self.actor = nn.Module()
        self.actor.logits = mlp([self.representation.out_dim, n_actions], init=init_heads)
        self.critic = nn.Module()
        self.critic.values = mlp([self.representation.out_dim, 1], init=init_heads)

    def forward(self, obs):
        z = self.representation(obs)
        return self.actor.logits(z), self.critic.values(z).squeeze(-1)


class DeepQNetworkOracle(nn.Module):
    def __init__(self, n_actions, q_hidden=(512,), init=None, **enc_kw):
        super().__init__()
        self.n_actions = n_actions
        self.representation = PoolEncoderQ(**enc_kw)
        self.eval_Q_head = nn.Module()
        self.eval_Q_head.q_value = mlp([self.representation.out_dim, *q_hidden, n_actions], init=init)
        self.target_representation = copy.deepcopy(self.representation)
        self.target_Q_head = copy.deepcopy(self.eval_Q_head)

    def forward(self, obs):
        return self.eval_Q_head.q_value(self.representation(obs))

    def target(self, obs):
        return self.target_Q_head.q_value(self.target_representation(obs))

    def copy_target(self):
        for e, t in zip(self.representation.parameters(), self.target_representation.parameters()):
            t.data.copy_(e)
        for e, t in zip(self.eval_Q_head.parameters(), self.target_Q_head.parameters()):
            t.data.copy_(e)

"""Policy distributions (reference: xuance/torch/rl_models/modules/distributions.py:96-218).  Same method
surface; the arithmetic is torch.distributions, as in the reference."""
import math

import torch
from torch.distributions import Categorical, Normal
from torch.nn.functional import softplus


class CategoricalDistribution:
    def __init__(self, action_dim):
        self.action_dim = action_dim
        self.distribution = None
        self.probs, self.logits = None, None

    def set_param(self, probs=None, logits=None):
        if probs is None and logits is None:
            raise RuntimeError("Failed to setup distributions without given probs or logits.")
        self.distribution = Categorical(probs=probs, logits=logits, validate_args=False)
        self.probs, self.logits = self.distribution.probs, self.distribution.logits

    def get_param(self):
        return self.logits

    def log_prob(self, x):
        return self.distribution.log_prob(x)

    def entropy(self):
        return self.distribution.entropy()

    def stochastic_sample(self):
        return self.distribution.sample()

    def deterministic_sample(self):
        return torch.argmax(self.distribution.probs, dim=-1)


class DiagGaussianDistribution:
    def __init__(self, action_dim):
        self.action_dim = action_dim
        self.mu, self.std, self.distribution = None, None, None

    def set_param(self, mu, std):
        self.mu, self.std = mu, std
        # validate_args=False: torch's argument validation does a host-synchronising `.all()` per construction and per
        # log_prob (not CUDA-graph capturable); the arithmetic is unchanged
        self.distribution = Normal(mu, std, validate_args=False)

    def get_param(self):
        return self.mu, self.std

    def log_prob(self, x):
        return self.distribution.log_prob(x).sum(-1)

    def entropy(self):
        return self.distribution.entropy().sum(-1)

    def stochastic_sample(self):
        return self.distribution.sample()

    def rsample(self):
        return self.distribution.rsample()

    def deterministic_sample(self):
        return self.mu


class ActivatedDiagGaussianDistribution(DiagGaussianDistribution):
    """tanh-squashed Gaussian of SAC (distributions.py:200-218): correction -2(log2 - u - softplus(-2u))."""

    def __init__(self, action_dim, activation_action, device):
        super().__init__(action_dim)
        self.activation_fn = activation_action()
        self.device = device

    def activated_rsample(self):
        return self.activation_fn(self.rsample())

    def activated_deterministic_sample(self):
        return self.activation_fn(self.deterministic_sample())

    def activated_rsample_and_logprob(self, noise=None):
        """``noise`` (standard normal, same shape as mu) may be supplied for reproducible parity tests; the
        default draws it exactly as Normal.rsample does."""
        if noise is None:
            pre = self.rsample()
        else:
            pre = self.mu + self.std * noise
        act = self.activation_fn(pre)
        log_prob = self.distribution.log_prob(pre)
        # log(2) as a constant (the reference builds Tensor([2.0]).log() on the host every call - not graph-capturable)
        correction = -2. * (math.log(2.0) - pre - softplus(-2. * pre))
        log_prob = log_prob + correction
        return act, log_prob.sum(-1)

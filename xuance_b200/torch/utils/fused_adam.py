"""Flat-bucket Adam with fused global-norm clipping (K7, include/xb200.h).

Replaces ``torch.nn.utils.clip_grad_norm_`` + ``torch.optim.Adam.step`` as the reference learners call them
(e.g. xuance/torch/learners/policy_gradient/ppo_learner.py:18-22, 61-65): same update rule (Adam, betas
(0.9, 0.999), ``eps`` as given, no weight decay / amsgrad), same clip rule
(coef = min(1, max_norm / (total_norm + 1e-6))).  It IS a ``torch.optim.Optimizer`` so the reference's
``LinearLR`` scheduler, ``optimizer.state_dict()['param_groups'][0]['lr']`` logging and checkpoint format keep
working.  Two kernel launches per step, no host synchronisation."""
import math

import torch

from ... import _lib
from .flat_bucket import FlatBucket


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, front=()):
        params = list(params)
        if weight_decay != 0.0:
            raise NotImplementedError("weight_decay is not on the hot path (reference configs use 0)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedAdam drives exactly one parameter group")
        self.bucket = FlatBucket(self.param_groups[0]["params"], front=front)
        dev = self.bucket.flat.device
        self.exp_avg = torch.zeros_like(self.bucket.flat)
        self.exp_avg_sq = torch.zeros_like(self.bucket.flat)
        self.step_count = 0
        self._hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        self._hyper_host = torch.zeros(4, dtype=torch.float32)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scratch = _lib.scratch(dev)
        self._step_t = torch.zeros((), dtype=torch.float32)
        for p, m, v in zip(self.bucket.params, self.bucket.views(self.exp_avg), self.bucket.views(self.exp_avg_sq)):
            self.state[p] = {"step": self._step_t, "exp_avg": m, "exp_avg_sq": v}

    @property
    def grad_norm(self):
        """Device scalar: the (pre-clip, post-averaging) global gradient L2 norm of the last clipped step."""
        return self._norm

    def zero_grad(self, set_to_none=False):
        self.bucket.zero_grad()

    def prepare(self):
        """Host half of a step: advance the step count and upload (step_size, bc2_sqrt, lr) for the kernels.  Kept apart
        from ``launch`` so that a CUDA graph can capture the device half only."""
        group = self.param_groups[0]
        self._opt_called = True     # what LRScheduler.step() looks at to tell "optimizer stepped first"
        self.step_count += 1
        self._step_t.fill_(float(self.step_count))
        b1, b2 = group["betas"]
        t = self.step_count
        self._hyper_host[0] = group["lr"] / (1.0 - b1 ** t)
        self._hyper_host[1] = math.sqrt(1.0 - b2 ** t)
        self._hyper_host[2] = group["lr"]
        self._hyper.copy_(self._hyper_host)     # 16-byte pageable H2D: the driver stages it before returning

    @torch.no_grad()
    def launch(self, max_norm=None, grad_scale=1.0, write_back_grad=False):
        """Device half: [global-norm reduction] + clip + Adam over the flat bucket (graph-capturable)."""
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        n = self.bucket.numel
        clip = float(max_norm) if max_norm is not None else -1.0
        if clip > 0:
            _lib.call("xb_grad_sumsq", _lib.ptr(self.bucket.grad), n, float(grad_scale), _lib.ptr(self._norm),
                      _lib.ptr(self._scratch))
        _lib.call("xb_adam_step", _lib.ptr(self.bucket.flat), _lib.ptr(self.bucket.grad), _lib.ptr(self.exp_avg),
                  _lib.ptr(self.exp_avg_sq), n, _lib.ptr(self._hyper), float(b1), float(b2), float(group["eps"]),
                  clip, _lib.ptr(self._norm), float(grad_scale), 1 if write_back_grad else 0)

    @torch.no_grad()
    def step(self, closure=None, max_norm=None, grad_scale=1.0, write_back_grad=False):
        """One Adam step over the bucket.  ``max_norm`` (float or None) applies clip_grad_norm_ semantics first;
        ``grad_scale`` multiplies the gradient (1/world_size after a sum all-reduce)."""
        self.prepare()
        self.launch(max_norm, grad_scale, write_back_grad)
        return None

    def snapshot(self):
        return (self.bucket.flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_count)

    def restore(self, snap):
        self.bucket.flat.copy_(snap[0])
        self.exp_avg.copy_(snap[1])
        self.exp_avg_sq.copy_(snap[2])
        self.step_count = snap[3]
        self._step_t.fill_(float(self.step_count))

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam state_dict (reference checkpoints) and copies it into the flat buffers."""
        groups = state_dict["param_groups"]
        for k in ("lr", "betas", "eps"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
        if "initial_lr" in groups[0]:
            self.param_groups[0]["initial_lr"] = groups[0]["initial_lr"]
        st = state_dict.get("state", {})
        # the saved ids enumerate the param group position by position (torch.optim.Optimizer.state_dict); the bucket holds
        # only the trainable, de-duplicated parameters, so map position -> parameter object -> bucket slot
        group_params = self.param_groups[0]["params"]
        saved_ids = groups[0]["params"]
        if len(saved_ids) != len(group_params):
            raise ValueError("optimizer checkpoint has %d parameters, this optimizer %d" % (len(saved_ids), len(group_params)))
        slot = {id(p): i for i, p in enumerate(self.bucket.params)}
        ms, vs = self.bucket.views(self.exp_avg), self.bucket.views(self.exp_avg_sq)
        for pos, pid in enumerate(saved_ids):
            if pid not in st:
                continue
            i = slot.get(id(group_params[pos]))
            if i is None:                     # frozen / duplicate parameter: no moments on this side
                continue
            m, v = st[pid]["exp_avg"], st[pid]["exp_avg_sq"]
            if tuple(m.shape) != tuple(ms[i].shape):
                raise ValueError("optimizer checkpoint: moment of parameter %d has shape %s, expected %s"
                                 % (pos, tuple(m.shape), tuple(ms[i].shape)))
            ms[i].copy_(m)
            vs[i].copy_(v)
            self.step_count = int(float(st[pid]["step"]))
        self._step_t.fill_(float(self.step_count))

"""Arithmetic model of the split-bf16 tensor-core layers (K9-TC, K12).  TEST INFRASTRUCTURE (see oracle/__init__).

The reference computes its layers in float32 (``nn.Conv2d`` / ``nn.Linear`` of xuance/torch/rl_models/representations/
cnn.py:45-50, 84-101 and heads/q_mix_head.py:52-95).  The tensor pipe multiplies bf16, so the kernels feed every float32
operand as P bf16 "planes" whose sum is the operand, and add up the plane-by-plane products whose plane indices sum to
less than P, accumulating in float32:

    P = 2 : x = hi + lo            products hi.hi + hi.lo + lo.hi                      |x - sum| <= 2^-16 |x|
    P = 3 : x = hi + mid + lo      products hh + hm + mh + hl + lh + mm                |x - sum| <= 2^-24 |x|

These functions restate that arithmetic in NumPy / torch-CPU (float64 accumulation) so tests can bound what the kernels
may differ from the float32 reference by, independent of any layout question.
"""
import numpy as np
import torch


def split_planes(x, planes=2):
    """float tensor -> float32 [planes, ...]: plane q = bfloat16(residual left by planes < q), as xb_split_bf16 and the
    kernels' epilogues compute it (round-to-nearest-even at every step)."""
    r = torch.as_tensor(x).float().clone()
    out = []
    for _ in range(planes):
        h = r.bfloat16().float()
        out.append(h)
        r = r - h
    return torch.stack(out).contiguous()


def split_residual_bound(planes):
    """Relative bound of |x - sum(planes)|: bf16 keeps 8 significant bits, so each plane leaves at most 2^-8 of what it
    rounded: 2^-16 after two planes (2^-17 typical), 2^-24 after three - float32's own precision."""
    return max(2.0 ** (-8 * planes), 2.0 ** -24)


def split_matmul(a, w, planes=2):
    """a [M, K] . w [N, K]^T through the plane products the kernels issue (pa + pb < planes), float64 accumulation."""
    ap, wp = split_planes(a, planes).double(), split_planes(w, planes).double()
    out = torch.zeros(a.shape[0], w.shape[0], dtype=torch.float64)
    for pa in range(planes):
        for pb in range(planes - pa):
            out += ap[pa] @ wp[pb].t()
    return out


def dropped_terms_bound(planes):
    """Relative size of the largest dropped product (plane indices summing to `planes`): 2^-8 per plane index."""
    return 2.0 ** (-8 * planes)

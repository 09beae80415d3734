"""EXPERIMENTAL host side of K12 (``xb_gemm_gather_tc``, include/xb200.h): geometry builders that express the NatureCNN
layers (xuance/torch/rl_models/representations/cnn.py:45-50, 84-101; modules/layers.py:16-65) and their data gradients
as the gathered-operand GEMM of ``xuance_b200/csrc/conv_index.h``, plus thin launch wrappers.

Nothing on a default path imports this module.  The geometry and the kernel's index arithmetic are verified on the host
(tests/test_conv_index.py emulates the shared-memory staging and the descriptor reads and compares with
``torch.nn.functional.conv2d`` / autograd); the kernel itself has not run on hardware yet.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from ... import _lib

KC = 64          # XB_CONV_KC
MAX_TAPS = 64    # XB_CONV_MAX_TAPS


@dataclass
class GatherGeometry:
    """One ``xb_gemm_gather_tc`` call: sites [B, OY, OX] over an NHWC input [B, IH, IW, C]; tap t reads
    (y*sy + dy[t], x*sx + dx[t]); the site's result row lands at (y*oys + oy0, x*oxs + ox0) of a [B, out_H, out_W] grid.
    ``fold`` > 1: ``fold`` horizontally adjacent pixels were folded into the channel axis (input viewed as
    [B, IH, W/fold, C*fold]) so that every 16-byte unit is 8 contiguous bf16 values."""
    B: int
    IH: int
    IW: int
    C: int
    OY: int
    OX: int
    sy: int
    sx: int
    dy: List[int]
    dx: List[int]
    out_H: int
    out_W: int
    oys: int = 1
    oxs: int = 1
    oy0: int = 0
    ox0: int = 0
    fold: int = 1
    taps: List[tuple] = field(default_factory=list)    # (kh, kw) of the ORIGINAL kernel per tap (before folding)

    @property
    def T(self):
        return len(self.dy)

    @property
    def K(self):
        return self.T * self.C

    @property
    def M(self):
        return self.B * self.OY * self.OX

    def check(self):
        assert self.C % 8 == 0, "channels per tap must be a multiple of 8 (one 16-byte unit)"
        assert self.K % KC == 0, "K = taps*channels must be a multiple of %d" % KC
        assert self.T <= MAX_TAPS
        assert all(-128 <= v <= 127 for v in self.dy + self.dx)
        assert (self.OY - 1) * self.oys + self.oy0 < self.out_H and (self.OX - 1) * self.oxs + self.ox0 < self.out_W
        return self


def conv_out(size, k, stride, pad):
    return (size + 2 * pad - k) // stride + 1


def conv_forward_geometry(B, H, W, C, KH, KW, stride, pad):
    """y = conv2d(x, w, stride, pad) with x NHWC [B,H,W,C] and the weight packed as [N, (kh, kw, c)]."""
    OY, OX = conv_out(H, KH, stride, pad), conv_out(W, KW, stride, pad)
    fold = 1
    if C % 8 != 0:
        assert 8 % C == 0, "channel counts below 8 must divide 8"
        fold = 8 // C
        assert W % fold == 0 and KW % fold == 0 and stride % fold == 0 and pad % fold == 0, \
            "folding %d pixels into the channel axis needs W, KW, stride and pad divisible by %d" % (fold, fold)
    dy, dx, taps = [], [], []
    for kh in range(KH):
        for kw2 in range(KW // fold):
            dy.append(kh - pad)
            dx.append(kw2 - pad // fold)
            taps.append((kh, kw2 * fold))
    return GatherGeometry(B=B, IH=H, IW=W // fold, C=C * fold, OY=OY, OX=OX, sy=stride, sx=stride // fold, dy=dy, dx=dx,
                          out_H=OY, out_W=OX, fold=fold, taps=taps).check()


def linear_geometry(B, in_features):
    """y = x @ W^T : one tap, one site per row."""
    return GatherGeometry(B=B, IH=1, IW=1, C=in_features, OY=1, OX=1, sy=1, sx=1, dy=[0], dx=[0], out_H=1, out_W=1,
                          taps=[(0, 0)]).check()


def conv_dgrad_geometries(B, H, W, C, KH, KW, stride, pad, N):
    """dx = conv2d_backward_input(dy): one GEMM per stride phase (iy % s, ix % s) over the output gradient
    dy NHWC [B, OY, OX, N]; each phase uses the kernel taps with kh = (py + pad) mod s, and writes the input-gradient
    pixels of its phase.  Returns [(geometry, [(kh, kw), ...])]; the matching weight matrix of a phase is
    [C, (tap, n)] = w[n, c, kh, kw]."""
    OY, OX = conv_out(H, KH, stride, pad), conv_out(W, KW, stride, pad)
    out = []
    for py in range(stride):
        for px in range(stride):
            ny, nx = (H - py + stride - 1) // stride, (W - px + stride - 1) // stride
            if ny <= 0 or nx <= 0:
                continue
            dy, dx, taps = [], [], []
            for kh in range(KH):
                if (py + pad - kh) % stride:
                    continue
                for kw in range(KW):
                    if (px + pad - kw) % stride:
                        continue
                    dy.append((py + pad - kh) // stride)
                    dx.append((px + pad - kw) // stride)
                    taps.append((kh, kw))
            if not taps:
                continue
            g = GatherGeometry(B=B, IH=OY, IW=OX, C=N, OY=ny, OX=nx, sy=1, sx=1, dy=dy, dx=dx, out_H=H, out_W=W,
                               oys=stride, oxs=stride, oy0=py, ox0=px, taps=taps).check()
            out.append((g, taps))
    return out


def dgrad_weight_matrix(w, taps):
    """torch conv weight [N, C, KH, KW] -> [C, (tap, n)] float32 for the taps of one dgrad phase."""
    cols = [w[:, :, kh, kw].t() for kh, kw in taps]          # each [C, N]
    return torch.cat(cols, dim=1).contiguous()


# ------------------------------------------------------------------------------------------------ device wrappers
def split_bf16(x):
    """float32 CUDA tensor -> (hi, lo) bfloat16 tensors of the same shape, x ~= hi + lo to 2^-17 relative."""
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    _lib.call("xb_split_bf16", _lib.ptr(x), x.numel(), _lib.ptr(hi), _lib.ptr(lo))
    return hi, lo


def pack_conv_weight(w):
    """[N, C, KH, KW] float32 CUDA -> (hi, lo) bfloat16 [N, KH*KW*C] in (kh, kw, c) column order."""
    w = w.contiguous()
    N, C, KH, KW = w.shape
    hi = torch.empty((N, KH * KW * C), dtype=torch.bfloat16, device=w.device)
    lo = torch.empty_like(hi)
    _lib.call("xb_pack_conv_weight", _lib.ptr(w), N, C, KH, KW, _lib.ptr(hi), _lib.ptr(lo))
    return hi, lo


def gemm_gather(in_hi, in_lo, w_hi, w_lo, geom, bias=None, relu=False, out_f32=None, out_hi=None, out_lo=None,
                out_ld=None, out_c0=0):
    """One K12 launch.  ``w_hi/w_lo`` [N, K]; outputs are caller-allocated matrices with ``out_ld`` elements per row
    (default N)."""
    N, K = w_hi.shape
    assert K == geom.K, (K, geom.K)
    out_ld = N if out_ld is None else out_ld
    dy = torch.tensor(geom.dy, dtype=torch.int8)
    dx = torch.tensor(geom.dx, dtype=torch.int8)
    _lib.call("xb_gemm_gather_tc", _lib.ptr(in_hi), _lib.ptr(in_lo), _lib.ptr(w_hi), _lib.ptr(w_lo),
              _lib.ptr(bias) if bias is not None else None, geom.B, geom.IH, geom.IW, geom.C, geom.OY, geom.OX, geom.sy,
              geom.sx, geom.T, dy.data_ptr(), dx.data_ptr(), N, 1 if relu else 0,
              _lib.ptr(out_hi) if out_hi is not None else None, _lib.ptr(out_lo) if out_lo is not None else None,
              _lib.ptr(out_f32) if out_f32 is not None else None, geom.out_H, geom.out_W, geom.oys, geom.oxs, geom.oy0,
              geom.ox0, out_ld, out_c0)
    return out_f32, out_hi, out_lo

"""Host side of K12 (``xb_gemm_gather_tc`` / ``xb_wgrad_gather_tc``, include/xb200.h): geometry builders that express the
NatureCNN layers (xuance/torch/rl_models/representations/cnn.py:45-50, 84-101; modules/layers.py:16-65), their data
gradients and weight gradients as the gathered-operand GEMM of ``xuance_b200/csrc/conv_index.h``, thin launch wrappers, and
``TensorCoreNatureCNN`` - the forward / backward orchestration behind ``compute="tc"`` of the pixel encoders.

The geometry and the kernel's index arithmetic are verified on the host (tests/test_conv_index.py emulates the
shared-memory staging, the descriptor reads and the per-plane MMA issue and compares with ``torch.nn.functional.conv2d`` /
autograd); tests/test_gpu_tc_conv.py pins every mode against float64 on B200.
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
# Operands are "plane tensors": bfloat16 [P, ...] with x = planes.sum(0); plane q = bf16 of the residual the planes before
# it leave.  P = 1 is exact for raw uint8 pixels (integers <= 256), P = 2 is exact to 2^-16, P = 3 to 2^-24.  A K12 launch
# multiplies an A operand of PA planes with a B operand of PB >= PA planes and keeps the products whose plane indices sum
# to < PB, each order of magnitude in its own float32 accumulator (added smallest first in the epilogue - the tensor core
# truncates when it adds into an accumulator).
def split_bf16(x, planes=2):
    """float32 CUDA tensor -> bfloat16 [planes, *x.shape]."""
    x = x.contiguous()
    out = torch.empty((planes,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    _lib.call("xb_split_bf16", _lib.ptr(x), x.numel(), planes, _lib.ptr(out))
    return out


def pack_conv_weight(w, planes=2, scale=1.0):
    """[N, C, KH, KW] float32 CUDA -> bfloat16 [planes, N, KH*KW*C] in (kh, kw, c) column order, times ``scale``."""
    w = w.contiguous()
    N, C, KH, KW = w.shape
    out = torch.empty((planes, N, KH * KW * C), dtype=torch.bfloat16, device=w.device)
    _lib.call("xb_pack_conv_weight", _lib.ptr(w), N, C, KH, KW, planes, float(scale), _lib.ptr(out))
    return out


def n_tile_for(N, planes_b):
    """Columns per work item: the largest multiple of 32 that divides N with planes_b * n_tile <= 256 (one tcgen05.mma
    spans the planes_b adjacent weight planes of a tile)."""
    t = (256 // planes_b) // 32 * 32
    while t >= 32:
        if N % t == 0:
            return t
        t -= 32
    raise ValueError("N = %d has no column tile that is a multiple of 32" % N)


def _plane_arg(t):
    """(pointer of plane 0, plane stride in elements) of a plane tensor whose planes are each contiguous."""
    assert t[0].is_contiguous(), "each plane must be contiguous"
    return _lib.ptr(t[0]), t.stride(0)


_TAPS = {}


def _taps(geom):
    """int8 host arrays of the tap offsets (kept alive per geometry: the ABI copies them at launch time)."""
    key = (tuple(geom.dy), tuple(geom.dx))
    if key not in _TAPS:
        _TAPS[key] = (torch.tensor(geom.dy, dtype=torch.int8), torch.tensor(geom.dx, dtype=torch.int8))
    return _TAPS[key]


def gemm_gather(x_pl, w_pl, geom, bias=None, relu=False, out_f32=None, out_pl=None, out_ld=None, out_c0=0, relu_mask=None,
                n_tile=None):
    """One K12 launch.  ``x_pl`` [PA, B, IH, IW, C] (any shape with that element order), ``w_pl`` [PB, N, K]; outputs are
    caller-allocated: ``out_f32`` [rows, out_ld] and / or ``out_pl`` [P_out, rows, out_ld]."""
    PB, N, K = w_pl.shape
    PA = x_pl.shape[0]
    assert K == geom.K and PA <= PB, (K, geom.K, PA, PB)
    out_ld = N if out_ld is None else out_ld
    n_tile = n_tile_for(N, PB) if n_tile is None else n_tile
    dy, dx = _taps(geom)
    xp, xs = _plane_arg(x_pl)
    wp, ws = _plane_arg(w_pl)
    op, os_ = _plane_arg(out_pl) if out_pl is not None else (None, 0)
    _lib.call("xb_gemm_gather_tc", PA, PB, xp, xs, wp, ws, _lib.ptr(bias) if bias is not None else None,
              _lib.ptr(relu_mask) if relu_mask is not None else None, geom.B, geom.IH, geom.IW, geom.C, geom.OY, geom.OX,
              geom.sy, geom.sx, geom.T, dy.data_ptr(), dx.data_ptr(), N, n_tile, 1 if relu else 0, op, os_,
              out_pl.shape[0] if out_pl is not None else 0, _lib.ptr(out_f32) if out_f32 is not None else None, geom.out_H,
              geom.out_W, geom.oys, geom.oxs, geom.oy0, geom.ox0, out_ld, out_c0)
    return out_f32, out_pl


def wgrad_gather(x_pl, g_pl, geom, splits, n_tile=None):
    """Partial weight gradients [splits, K, N] (float32) of the gathered GEMM ``geom`` for the output gradient planes
    ``g_pl`` [PB, sites, N]."""
    PB, _, N = g_pl.shape
    PA = x_pl.shape[0]
    assert PA <= PB and g_pl.stride(2) == 1
    n_tile = n_tile_for(N, PB) if n_tile is None else n_tile
    partials = torch.empty((splits, geom.K, N), dtype=torch.float32, device=g_pl.device)
    dy, dx = _taps(geom)
    xp, xs = _plane_arg(x_pl)
    _lib.call("xb_wgrad_gather_tc", PA, PB, xp, xs, _lib.ptr(g_pl), g_pl.stride(0), g_pl.stride(1), geom.B, geom.IH, geom.IW,
              geom.C, geom.OY, geom.OX, geom.sy, geom.sx, geom.T, dy.data_ptr(), dx.data_ptr(), N, n_tile, splits,
              _lib.ptr(partials))
    return partials


def wgrad_reduce(partials, N, C, KH, KW, out=None, accumulate=False, scale=1.0):
    """[splits, (kh,kw,c), N] partials -> torch-layout weight gradient [N, C, KH, KW] (times ``scale``)."""
    out = torch.empty((N, C, KH, KW), dtype=torch.float32, device=partials.device) if out is None else out
    _lib.call("xb_wgrad_reduce", _lib.ptr(partials), partials.shape[0], N, C, KH, KW, float(scale), _lib.ptr(out),
              1 if accumulate else 0)
    return out


def wgrad_splits(M, K, n_tiles=1, sm_count=148):
    """Split count for the weight gradient: about two work items per SM, at least 8 chunks of 64 sites per split, at most
    4096 sites per split (the tensor core truncates on every addition into the accumulator: a chain of n K-steps carries a
    bias of up to n * 2^-24 relative, so long reductions are cut and the pieces added in float32 by xb_wgrad_reduce), and
    the rule of ``xb_wgrad_sites_per_split`` (no empty split)."""
    m_tiles = -(-K // 128) * n_tiles
    s = max(1, min((2 * sm_count) // m_tiles, M // (8 * KC)))
    s = max(s, -(-M // 4096))
    while s > 1:
        per = -(-(-(-M // s)) // KC) * KC
        if (s - 1) * per < M:
            break
        s -= 1
    return s


class CudaBackend:
    """The product backend of ``TensorCoreNatureCNN``: every method is one or a few C-ABI launches (K12).  The host test
    substitutes an emulated backend with the same methods (tests/test_conv_index.py) to check the orchestration."""

    def __init__(self, planes=3):
        self.planes = planes

    def split(self, x):
        return split_bf16(x, self.planes)

    def pack_weight(self, w4d, scale=1.0):
        return pack_conv_weight(w4d, self.planes, scale)

    def empty_planes(self, shape, like):
        return torch.empty((self.planes,) + tuple(shape), dtype=torch.bfloat16, device=like.device)

    def empty_f32(self, shape, like):
        return torch.empty(shape, dtype=torch.float32, device=like.device)

    def gemm(self, x_pl, w_pl, geom, bias=None, relu=False, out_f32=None, out_pl=None, out_ld=None, out_c0=0, mask=None):
        gemm_gather(x_pl, w_pl, geom, bias=bias, relu=relu, out_f32=out_f32, out_pl=out_pl, out_ld=out_ld, out_c0=out_c0,
                    relu_mask=mask)

    def wgrad(self, x_pl, g_pl, geom, N, C, KH, KW, scale=1.0):
        nt = N // n_tile_for(N, g_pl.shape[0])
        splits = wgrad_splits(geom.M, geom.K, nt)
        return wgrad_reduce(wgrad_gather(x_pl, g_pl, geom, splits), N, C, KH, KW, scale=scale)

    def colsum(self, g_pl):
        # bias gradient: reduce the bf16 planes straight into float32 (no float32 copy of the [P, sites, N] tensor)
        return g_pl.sum(dim=1, dtype=torch.float32).sum(0)

    def to_float(self, pl):
        return pl.float().sum(0)


class TensorCoreNatureCNN:
    """The convolution stack + hidden layer of AC_CNN_Atari / Basic_CNN's conv part (cnn.py:45-50, 84-101) as K12 launches,
    forward AND backward, for one batch size.  Activations live as bf16 plane tensors (NHWC) between layers; the ReLU
    derivative is applied inside the data-gradient GEMM's epilogue from plane 0 of the saved activation.

    The input is either ``planes`` planes of the normalised observation x/255, or ONE plane holding the raw uint8 pixel
    values (exact in bf16): then the first layer's packed weights and its weight gradient carry the 1/255, the layer's A
    operand costs a third of the traffic and a third of the tensor work, and the minibatch gather writes one plane.

    ``convs``: nn.Conv2d modules (padding (k - s)//2 as layers.py:46 builds them), each followed by ReLU;
    ``fc``: nn.Linear over the NCHW-flattened last feature map, followed by ReLU."""

    def __init__(self, convs, fc, in_hwc, backend=None):
        self.convs, self.fc = list(convs), fc
        self.in_hwc = tuple(in_hwc)
        self.be = backend if backend is not None else CudaBackend()
        self._plans = {}

    def parameters(self):
        ps = []
        for c in self.convs:
            ps += [c.weight, c.bias]
        if self.fc is not None:
            ps += [self.fc.weight, self.fc.bias]
        return ps

    def _plan(self, B):
        if B in self._plans:
            return self._plans[B]
        H, W, C = self.in_hwc
        layers = []
        for conv in self.convs:
            N, Cw, KH, KW = conv.weight.shape
            s, p = conv.stride[0], conv.padding[0]
            assert Cw == C and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
            fwd = conv_forward_geometry(B, H, W, C, KH, KW, s, p)
            dgrad = conv_dgrad_geometries(B, H, W, C, KH, KW, s, p, N) if layers else None   # no gradient w.r.t. pixels
            layers.append(dict(kind="conv", N=N, C=C, KH=KH, KW=KW, fwd=fwd, dgrad=dgrad))
            H, W, C = fwd.OY, fwd.OX, N
        if self.fc is not None:
            N, K = self.fc.weight.shape
            assert K == H * W * C
            layers.append(dict(kind="fc", N=N, C=C, KH=H, KW=W, fwd=linear_geometry(B, K), dgrad=linear_geometry(B, N)))
        self._plans[B] = layers
        return layers

    # ---- forward: returns float32 [sites of the last layer, features]; keeps what backward needs
    def forward(self, x_pl, B):
        be, plan = self.be, self._plan(B)
        raw = x_pl.shape[0] == 1 and be.planes > 1          # one exact plane of uint8 values: 1/255 goes into the weights
        saved, cur, out_f32 = [], x_pl, None
        for li, L in enumerate(plan):
            mod = self.convs[li] if L["kind"] == "conv" else self.fc
            N, g = L["N"], L["fwd"]
            w4 = (mod.weight if L["kind"] == "conv" else mod.weight.reshape(N, L["C"], L["KH"], L["KW"])).detach()
            scale = 1.0 / 255.0 if (raw and li == 0) else 1.0
            w_pl = be.pack_weight(w4, scale)
            out_pl = be.empty_planes((g.M, N), cur)
            out_f32 = be.empty_f32((g.M, N), cur) if li == len(plan) - 1 else None
            be.gemm(cur, w_pl, g, bias=mod.bias.detach(), relu=True, out_f32=out_f32, out_pl=out_pl, out_ld=N)
            saved.append(dict(x=cur, y=out_pl, w4=w4, scale=scale))
            cur = out_pl
        self._saved = (B, saved)
        return out_f32

    # ---- backward: dz float32 = gradient w.r.t. the array forward() returned; returns gradients in parameters() order
    def backward(self, dz):
        be = self.be
        B, saved = self._saved
        plan = self._plan(B)
        grads = [None] * (2 * len(plan))
        y_last = saved[-1]["y"]
        g_pl = be.split(dz * (y_last[0] > 0).to(dz.dtype))            # ReLU derivative of the last layer (plane 0 > 0)
        for li in range(len(plan) - 1, -1, -1):
            L, sv = plan[li], saved[li]
            N, C, KH, KW = L["N"], L["C"], L["KH"], L["KW"]
            # -- weight and bias gradients (the weight gradient gathers exactly like the forward)
            dw = be.wgrad(sv["x"], g_pl, L["fwd"], N, C, KH, KW, sv["scale"])
            grads[2 * li] = dw if L["kind"] == "conv" else dw.reshape(N, C * KH * KW)
            grads[2 * li + 1] = be.colsum(g_pl)
            if li == 0:
                break
            # -- data gradient, masked by the previous activation's ReLU derivative inside the GEMM epilogue
            prev_y = saved[li - 1]["y"]                   # [P, sites_prev, C] == NHWC input planes of this layer
            out_pl = be.empty_planes(tuple(prev_y.shape[1:]), prev_y)
            if L["kind"] == "conv":
                for geom, taps in L["dgrad"]:
                    w_pl = be.split(dgrad_weight_matrix(sv["w4"], taps))          # [P, C, (tap, n)]
                    be.gemm(g_pl, w_pl, geom, out_pl=out_pl, out_ld=C, mask=prev_y[0])
            else:
                K = C * KH * KW
                w_pl = be.split(sv["w4"].permute(0, 2, 3, 1).reshape(N, K).t().contiguous())   # [P, K (h,w,c), N]
                be.gemm(g_pl, w_pl, L["dgrad"], out_pl=out_pl, out_ld=K, mask=prev_y[0])
            g_pl = out_pl
        return grads


class _TCEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x_pl, B, *params):
        ctx.enc = enc
        return enc.forward(x_pl, B)

    @staticmethod
    def backward(ctx, dz):
        return (None, None, None, *ctx.enc.backward(dz.contiguous()))


def tc_encode(enc, x_pl, B):
    """Differentiable call of a ``TensorCoreNatureCNN``: gradients reach the conv / linear parameters."""
    return _TCEncoderFn.apply(enc, x_pl, B, *enc.parameters())

"""CPU: index arithmetic and layer geometry of the EXPERIMENTAL tensor-core conv path (K12, xuance_b200/csrc/conv_tc.cu).

The kernel cannot run here, so its data movement is emulated on the host WITH THE KERNEL'S OWN index functions
(xuance_b200/csrc/conv_index.h is compiled into tests/csrc/conv_emul.cpp by g++): units are staged into an emulated
shared-memory image, read back the way the tcgen05 descriptors address it, and the result is compared with torch's fp64
convolution / autograd on the un-split operands - which also checks the split-bf16 numerics claim (three products of
hi/lo bf16 pairs reproduce the fp32 result to ~1e-5)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from xuance_b200.torch.utils import tc_conv as tc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    src, so = os.path.join(HERE, "csrc", "conv_emul.cpp"), os.path.join(HERE, "csrc", "_conv_emul.so")
    hdr = os.path.join(HERE, "..", "xuance_b200", "csrc", "conv_index.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    P, i, l = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.emul_gemm_gather.argtypes = [P, P, P, P] + [i] * 9 + [P, P] + [i] * 7 + [l, i, i, P]
    lib.emul_gemm_gather.restype = i
    lib.emul_pack_weight.argtypes = [P, i, i, i, i, P]
    return lib


def _split(x):
    hi = x.float().bfloat16().float()
    lo = (x.float() - hi).bfloat16().float()
    return hi.contiguous(), lo.contiguous()


def _run(lib, geom, x_nhwc, w_mat, out, out_ld, out_c0=0, stages=3):
    """x_nhwc: float64 tensor viewed as the geometry's [B, IH, IW, C]; w_mat float64 [N, K]; out float64 [rows, out_ld]."""
    xh, xl = _split(x_nhwc)
    wh, wl = _split(w_mat)
    N, K = w_mat.shape
    assert K == geom.K and xh.numel() == geom.B * geom.IH * geom.IW * geom.C
    dy, dx = np.asarray(geom.dy, np.int8), np.asarray(geom.dx, np.int8)
    rc = lib.emul_gemm_gather(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), geom.B, geom.IH, geom.IW, geom.C,
                              geom.OY, geom.OX, geom.sy, geom.sx, geom.T, dy.ctypes.data, dx.ctypes.data, N, geom.out_H,
                              geom.out_W, geom.oys, geom.oxs, geom.oy0, geom.ox0, out_ld, out_c0, stages, out.data_ptr())
    assert rc == 0


def _pack(lib, w):
    N, C, KH, KW = w.shape
    wf = w.float().contiguous()
    packed = torch.empty((N, KH * KW * C), dtype=torch.float32)
    lib.emul_pack_weight(wf.data_ptr(), N, C, KH, KW, packed.data_ptr())
    ref = w.permute(0, 2, 3, 1).reshape(N, -1).float()         # (kh, kw, c) column order
    assert torch.equal(packed, ref)
    return packed.double()


TOL = dict(rtol=0, atol=4e-5)     # |x - hi - lo| <= 2^-17 |x| per operand; sums of <= 6400 products of O(1) values / sqrt(K)


@pytest.mark.parametrize("name,B,H,W,C,N,k,s", [("conv1", 2, 84, 84, 4, 32, 8, 4), ("conv2", 2, 21, 21, 32, 64, 4, 2),
                                                ("conv3", 3, 10, 10, 64, 64, 3, 1), ("odd", 1, 13, 9, 8, 32, 4, 2)])
def test_forward_conv_layers(emul, name, B, H, W, C, N, k, s):
    """The three NatureCNN convolutions with the reference's padding rule (k - s)//2 (layers.py:46), NHWC input."""
    torch.manual_seed(len(name))
    pad = (k - s) // 2
    x = torch.rand(B, C, H, W, dtype=torch.float64)                       # observations / 255 live in [0, 1]
    w = torch.randn(N, C, k, k, dtype=torch.float64) / np.sqrt(C * k * k)
    g = tc.conv_forward_geometry(B, H, W, C, k, k, s, pad)
    assert g.K == C * k * k and (g.fold == 2) == (C == 4)
    out = torch.full((g.M, N), float("nan"), dtype=torch.float64)
    _run(emul, g, x.permute(0, 2, 3, 1).contiguous(), _pack(emul, w), out, N)
    want = F.conv2d(x, w, stride=s, padding=pad).permute(0, 2, 3, 1).reshape(g.M, N)
    np.testing.assert_allclose(out.numpy(), want.numpy(), **TOL)


def test_linear_layer_with_column_split_and_row_tail(emul):
    """Linear(6400 -> 512) over the NHWC-flattened conv3 output: N = 512 runs as two 256-column calls; 130 rows leave a
    2-row tail tile.  The reference flattens NCHW (cnn.py:92), so the packed weight permutes its columns to (h, w, c)."""
    torch.manual_seed(0)
    Bn, Cc, Hh, Ww, N = 130, 64, 10, 10, 512
    feat = torch.rand(Bn, Cc, Hh, Ww, dtype=torch.float64)
    w = torch.randn(N, Cc * Hh * Ww, dtype=torch.float64) / 80.0
    packed = _pack(emul, w.reshape(N, Cc, Hh, Ww))                         # columns (h, w, c)
    g = tc.linear_geometry(Bn, Cc * Hh * Ww)
    out = torch.full((Bn, N), float("nan"), dtype=torch.float64)
    x_nhwc = feat.permute(0, 2, 3, 1).reshape(Bn, 1, 1, -1).contiguous()
    for c0 in (0, 256):
        _run(emul, g, x_nhwc, packed[c0:c0 + 256].contiguous(), out, N, out_c0=c0, stages=2)
    want = feat.reshape(Bn, -1) @ w.t()
    np.testing.assert_allclose(out.numpy(), want.numpy(), **TOL)


@pytest.mark.parametrize("name,B,H,W,C,N,k,s", [("conv3", 2, 10, 10, 64, 64, 3, 1), ("conv2", 2, 21, 21, 32, 64, 4, 2),
                                                ("s4", 1, 20, 20, 8, 32, 8, 4)])
def test_data_gradient_phases(emul, name, B, H, W, C, N, k, s):
    """grad_input of a convolution as one gathered GEMM per stride phase over the output gradient, vs autograd."""
    torch.manual_seed(1)
    pad = (k - s) // 2
    x = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(N, C, k, k, dtype=torch.float64) / np.sqrt(N * k * k / (s * s))
    y = F.conv2d(x, w, stride=s, padding=pad)
    gy = torch.randn_like(y)
    (want,) = torch.autograd.grad(y, x, gy)
    gy_nhwc = gy.permute(0, 2, 3, 1).contiguous()
    out = torch.full((B * H * W, C), float("nan"), dtype=torch.float64)
    phases = tc.conv_dgrad_geometries(B, H, W, C, k, k, s, pad, N)
    assert len(phases) == s * s and sum(len(t) for _, t in phases) == k * k
    for g, taps in phases:
        _run(emul, g, gy_nhwc, tc.dgrad_weight_matrix(w, taps), out, C)
    np.testing.assert_allclose(out.reshape(B, H, W, C).permute(0, 3, 1, 2).numpy(), want.numpy(), **TOL)

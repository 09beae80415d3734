"""GPU bring-up tests of the EXPERIMENTAL K12 tensor-core layers (xuance_b200/csrc/conv_tc.cu).  They run only with
XB_EXPERIMENTAL_TC=1: the kernel is compiled and host-verified (tests/test_conv_index.py) but has not been on hardware
yet, so it stays out of the default ``-m gpu`` run until it has."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("XB_EXPERIMENTAL_TC") != "1",
                                                  reason="K12 bring-up: set XB_EXPERIMENTAL_TC=1")]
DEV = "cuda:0"


def test_split_and_pack():
    from xuance_b200.torch.utils import tc_conv as tc
    x = torch.randn(1000, 37, device=DEV)
    hi, lo = tc.split_bf16(x)
    assert torch.equal(hi, x.bfloat16()) and torch.equal(lo, (x - hi.float()).bfloat16())
    w = torch.randn(32, 4, 8, 8, device=DEV)
    wh, wl = tc.pack_conv_weight(w)
    ref = w.permute(0, 2, 3, 1).reshape(32, -1)
    assert torch.equal(wh, ref.bfloat16()) and torch.equal(wl, (ref - wh.float()).bfloat16())


@pytest.mark.parametrize("B,H,W,C,N,k,s", [(2, 84, 84, 4, 32, 8, 4), (3, 21, 21, 32, 64, 4, 2), (5, 10, 10, 64, 64, 3, 1),
                                            (256, 21, 21, 32, 64, 4, 2)])
def test_forward_conv(B, H, W, C, N, k, s):
    from xuance_b200.torch.utils import tc_conv as tc
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    pad = (k - s) // 2
    x = torch.rand(B, H, W, C, device=DEV)
    w = torch.randn(N, C, k, k, device=DEV) / np.sqrt(C * k * k)
    b = torch.randn(N, device=DEV) * 0.1
    g = tc.conv_forward_geometry(B, H, W, C, k, k, s, pad)
    out = torch.full((g.M, N), float("nan"), device=DEV)
    oh, ol = torch.zeros((g.M, N), dtype=torch.bfloat16, device=DEV), torch.zeros((g.M, N), dtype=torch.bfloat16, device=DEV)
    tc.gemm_gather(*tc.split_bf16(x), *tc.pack_conv_weight(w), g, bias=b, relu=True, out_f32=out, out_hi=oh, out_lo=ol)
    want = F.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=s, padding=pad))
    want = want.permute(0, 2, 3, 1).reshape(g.M, N)
    np.testing.assert_allclose(out.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=5e-5)
    np.testing.assert_allclose((oh.float() + ol.float()).cpu().numpy(), out.cpu().numpy(), rtol=2e-5, atol=1e-6)


def test_encoder_matches_cudnn_fp32():
    from helpers import build_product_ppo_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    m_ref = build_product_ppo_model(4, DEV).to(DEV)
    m_tc = build_product_ppo_model(4, DEV).to(DEV)
    m_tc.load_state_dict(m_ref.state_dict())
    m_tc.representation.set_compute("tc")
    obs = torch.randint(0, 256, (64, 84, 84, 4), dtype=torch.uint8, device=DEV)
    R = torch.randn(64, 512, device=DEV)
    z_ref = m_ref.representation(obs).embeddings
    z_tc = m_tc.representation(obs).embeddings
    np.testing.assert_allclose(z_tc.detach().cpu().numpy(), z_ref.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    (z_ref * R).sum().backward()
    (z_tc * R).sum().backward()
    for (k, p), (_, q) in zip(m_ref.representation.named_parameters(), m_tc.representation.named_parameters()):
        scale = float(p.grad.abs().max())
        np.testing.assert_allclose(q.grad.cpu().numpy(), p.grad.cpu().numpy(), rtol=0, atol=3e-4 * scale, err_msg=k)

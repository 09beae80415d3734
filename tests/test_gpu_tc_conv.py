"""GPU tests of the K12 tensor-core layers (xuance_b200/csrc/conv_tc.cu).

Forward (operand preparation + the gathered GEMM of the three NatureCNN convolutions) passed on B200 in round 1
(profiles/r01_k12_bringup_forward.log) and runs by default.  The backward tests (data-gradient phases, MN-major weight
gradient, whole encoder) run only with XB_EXPERIMENTAL_TC=1: the backward executed on hardware and agreed with cuDNN fp32
to ~4e-3 of max|grad| on the first layer (profiles/r01_k12_bringup_encoder.log) - the size of a single ReLU-boundary
mask flip, but not yet separated from a defect - so per-layer parity below is next round's first item."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bringup = pytest.mark.skipif(os.environ.get("XB_EXPERIMENTAL_TC") != "1", reason="K12 backward bring-up: set XB_EXPERIMENTAL_TC=1")
DEV = "cuda:0"


def _planes_ref(x, planes):
    r, out = x.clone(), []
    for _ in range(planes):
        h = r.bfloat16()
        out.append(h)
        r = r - h.float()
    return torch.stack(out)


@pytest.mark.parametrize("planes", [2, 3])
def test_split_and_pack(planes):
    from xuance_b200.torch.utils import tc_conv as tc
    for shape in ((1000, 37), (4099,), (8, 8)):                      # 37000 = 8*4625, a ragged length, one vector
        x = torch.randn(*shape, device=DEV)
        assert torch.equal(tc.split_bf16(x, planes), _planes_ref(x, planes))
    w = torch.randn(32, 4, 8, 8, device=DEV)
    assert torch.equal(tc.pack_conv_weight(w, planes), _planes_ref(w.permute(0, 2, 3, 1).reshape(32, -1), planes))


def _forward_conv(B, H, W, C, N, k, s, planes, atol):
    from xuance_b200.torch.utils import tc_conv as tc
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    pad = (k - s) // 2
    x = torch.rand(B, H, W, C, device=DEV)
    w = torch.randn(N, C, k, k, device=DEV) / np.sqrt(C * k * k)
    b = torch.randn(N, device=DEV) * 0.1
    g = tc.conv_forward_geometry(B, H, W, C, k, k, s, pad)
    out = torch.full((g.M, N), float("nan"), device=DEV)
    opl = torch.zeros((planes, g.M, N), dtype=torch.bfloat16, device=DEV)
    tc.gemm_gather(tc.split_bf16(x, planes), tc.pack_conv_weight(w, planes), g, bias=b, relu=True, out_f32=out, out_pl=opl)
    want = F.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=s, padding=pad))
    want = want.permute(0, 2, 3, 1).reshape(g.M, N)
    np.testing.assert_allclose(out.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=atol)
    np.testing.assert_allclose(opl.float().sum(0).cpu().numpy(), out.cpu().numpy(), rtol=2e-5 if planes == 2 else 3e-7, atol=1e-6)


@pytest.mark.parametrize("B,H,W,C,N,k,s", [(2, 84, 84, 4, 32, 8, 4), (3, 21, 21, 32, 64, 4, 2), (5, 10, 10, 64, 64, 3, 1),
                                            (256, 21, 21, 32, 64, 4, 2)])
def test_forward_conv(B, H, W, C, N, k, s):
    """Two planes per operand (hi, lo): the configuration that passed on B200 in round 1."""
    _forward_conv(B, H, W, C, N, k, s, 2, 5e-5)


@bringup
@pytest.mark.parametrize("B,H,W,C,N,k,s", [(2, 84, 84, 4, 32, 8, 4), (3, 21, 21, 32, 64, 4, 2), (64, 10, 10, 64, 64, 3, 1)])
def test_forward_conv_three_planes(B, H, W, C, N, k, s):
    """Three planes (hi, mid, lo), six products: float32-grade (fp32 accumulation in TMEM bounds it, not the operands)."""
    _forward_conv(B, H, W, C, N, k, s, 3, 2e-6)


@bringup
@pytest.mark.parametrize("B,H,W,C,N,k,s", [(3, 21, 21, 32, 64, 4, 2), (5, 10, 10, 64, 64, 3, 1), (64, 21, 21, 32, 64, 4, 2)])
def test_data_gradient_with_mask(B, H, W, C, N, k, s):
    """grad_input of one convolution (one GEMM per stride phase) times a GIVEN ReLU-derivative mask vs autograd."""
    from xuance_b200.torch.utils import tc_conv as tc
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    pad = (k - s) // 2
    x = torch.randn(B, C, H, W, device=DEV, dtype=torch.float64, requires_grad=True)
    w = torch.randn(N, C, k, k, device=DEV) / np.sqrt(N * k * k / (s * s))
    y = F.conv2d(x, w.double(), stride=s, padding=pad)
    gy = torch.randn(y.shape, device=DEV)
    (want,) = torch.autograd.grad(y, x, gy.double())
    act = torch.randn(B, H, W, C, device=DEV)                      # the "saved activation": mask = act > 0
    act_hi = act.bfloat16()
    want = want.permute(0, 2, 3, 1) * (act_hi.float() > 0)
    g_pl = tc.split_bf16(gy.permute(0, 2, 3, 1).reshape(-1, N).contiguous())
    opl = torch.full((2, B * H * W, C), float("nan"), dtype=torch.bfloat16, device=DEV)
    for geom, taps in tc.conv_dgrad_geometries(B, H, W, C, k, k, s, pad, N):
        tc.gemm_gather(g_pl, tc.split_bf16(tc.dgrad_weight_matrix(w, taps)), geom, out_pl=opl, out_ld=C, relu_mask=act_hi)
    got = opl.float().sum(0).reshape(B, H, W, C)
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=5e-5)


@bringup
@pytest.mark.parametrize("B,H,W,C,N,k,s", [(2, 84, 84, 4, 32, 8, 4), (4, 21, 21, 32, 64, 4, 2), (5, 10, 10, 64, 64, 3, 1),
                                            (64, 84, 84, 4, 32, 8, 4), (256, 10, 10, 64, 64, 3, 1)])
def test_weight_gradient(B, H, W, C, N, k, s):
    """grad_weight of one convolution (MN-major operands, site splits, ordered reduce) vs autograd in float64."""
    from xuance_b200.torch.utils import tc_conv as tc
    torch.manual_seed(2)
    pad = (k - s) // 2
    x = torch.rand(B, H, W, C, device=DEV)
    w = (torch.randn(N, C, k, k, device=DEV, dtype=torch.float64) / np.sqrt(C * k * k)).requires_grad_(True)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w, stride=s, padding=pad)
    gy = torch.randn(y.shape, device=DEV) / np.sqrt(y[0, 0].numel() * B)
    (want,) = torch.autograd.grad(y, w, gy.double())
    g = tc.conv_forward_geometry(B, H, W, C, k, k, s, pad)
    for planes, atol in ((2, 5e-5), (3, 2e-6)):
        x_pl = tc.split_bf16(x, planes)
        g_pl = tc.split_bf16(gy.permute(0, 2, 3, 1).reshape(g.M, N).contiguous(), planes)
        for splits in sorted({1, tc.wgrad_splits(g.M, g.K)}):
            dw = tc.wgrad_reduce(tc.wgrad_gather(x_pl, g_pl, g, splits), N, C, k, k)
            np.testing.assert_allclose(dw.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=atol,
                                       err_msg="planes=%d splits=%d" % (planes, splits))


@bringup
@pytest.mark.parametrize("planes,fwd_tol,grad_tol", [(2, 1e-4, 1e-2), (3, 5e-6, 2e-5)])
def test_encoder_matches_cudnn_fp32(planes, fwd_tol, grad_tol):
    """Whole encoder forward + backward vs the cuDNN fp32 path.  With two planes the forward is within ~1e-5, which lets an
    activation that close to zero take the other side of its ReLU than in the float32 network: one such flip shifts the
    upstream gradients by O(1/B) (host emulation at B = 64: 3e-3 relative on the first layer), so
    that comparison is made in norm with a loose bound.  With three planes the operands are exact to 2^-24 and the
    gradients agree to float32 rounding."""
    from helpers import build_product_ppo_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    m_ref = build_product_ppo_model(4, DEV).to(DEV)
    m_tc = build_product_ppo_model(4, DEV).to(DEV)
    m_tc.load_state_dict(m_ref.state_dict())
    m_tc.representation.tc_planes = planes
    m_tc.representation.set_compute("tc")
    obs = torch.randint(0, 256, (64, 84, 84, 4), dtype=torch.uint8, device=DEV)
    R = torch.randn(64, 512, device=DEV)
    z_ref = m_ref.representation(obs).embeddings
    z_tc = m_tc.representation(obs).embeddings
    np.testing.assert_allclose(z_tc.detach().cpu().numpy(), z_ref.detach().cpu().numpy(), rtol=fwd_tol, atol=fwd_tol)
    (z_ref * R).sum().backward()
    (z_tc * R).sum().backward()
    for (k, p), (_, q) in zip(m_ref.representation.named_parameters(), m_tc.representation.named_parameters()):
        rel = float((q.grad - p.grad).norm() / p.grad.norm())
        assert rel < grad_tol, (k, rel)


@bringup
@pytest.mark.parametrize("planes", [2, 3])
def test_gather_obs_planes(planes):
    """K3-P: uint8 rows gathered straight into the bf16 planes of x/255 (bit-exact vs gather + true division + split)."""
    from xuance_b200 import _lib
    buf = torch.randint(0, 256, (6, 5, 84, 84, 4), dtype=torch.uint8, device=DEV)
    idx = torch.tensor([29, 0, 7, 7, 13, 1, 28], dtype=torch.int64, device=DEV)
    out = torch.empty((planes, idx.numel(), 84, 84, 4), dtype=torch.bfloat16, device=DEV)
    _lib.call("xb_gather_obs_planes", _lib.ptr(buf), _lib.ptr(idx), idx.numel(), 84 * 84 * 4, planes, _lib.ptr(out))
    ref = _planes_ref(buf.reshape(-1, 84, 84, 4)[idx].float() / 255.0, planes)
    assert torch.equal(out, ref)
    out2 = torch.empty((planes, 30, 84, 84, 4), dtype=torch.bfloat16, device=DEV)
    _lib.call("xb_gather_obs_planes", _lib.ptr(buf), None, 30, 84 * 84 * 4, planes, _lib.ptr(out2))
    assert torch.equal(out2, _planes_ref(buf.reshape(-1, 84, 84, 4).float() / 255.0, planes))


@bringup
def test_ppo_update_tc_three_planes_matches_fp32():
    """PPO_Learner.update with compute='tc' (three planes) next to the cuDNN fp32 learner: same losses, same parameters."""
    from helpers import build_product_ppo_model, ppo_config
    from xuance_b200.common import BaseCallback
    from xuance_b200.torch.learners import PPO_Learner
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    A, B = 6, 256
    m_ref = build_product_ppo_model(A, DEV)
    m_tc = build_product_ppo_model(A, DEV)
    m_tc.load_state_dict(m_ref.state_dict())
    m_tc.representation.tc_planes = 3
    m_tc.representation.set_compute("tc")
    cfg = ppo_config(DEV, running_steps=256 * 128 * 10)
    l_ref, l_tc = PPO_Learner(cfg, m_ref, BaseCallback()), PPO_Learner(cfg, m_tc, BaseCallback())
    rng = np.random.default_rng(5)
    for it in range(3):
        s = {"obs": torch.from_numpy(rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8)).cuda(),
             "actions": rng.integers(0, A, size=B).astype(np.float32), "returns": rng.normal(size=B).astype(np.float32),
             "advantages": rng.normal(size=B).astype(np.float32),
             "aux_batch": {"old_logp": (rng.normal(size=B) * 0.05 - np.log(A)).astype(np.float32)}}
        i_ref, i_tc = l_ref.update(**s), l_tc.update(**s)
        for k in ("actor_loss", "critic_loss", "entropy", "predict_value"):
            np.testing.assert_allclose(i_tc[k], i_ref[k], rtol=2e-4, atol=1e-5, err_msg=k)
    for (k, p), (_, q) in zip(m_ref.state_dict().items(), m_tc.state_dict().items()):
        np.testing.assert_allclose(q.cpu().numpy(), p.cpu().numpy(), rtol=1e-3, atol=1e-4, err_msg=k)

"""GPU parity: K1 store, K2 GAE scan, K3 gathers vs the NumPy oracle (through the C-ABI via the product classes)."""
import numpy as np
import pytest
import torch

from helpers import synth_rollout, fill_buffers
from oracle.onpolicy import OnPolicyBufferOracle

pytestmark = pytest.mark.gpu

GAE_TOL = dict(rtol=1e-5, atol=2e-6)   # warp scan reorders fp32 ops relative to the sequential loop


def _mk(N, T, obs_shape, atari, use_gae=True, use_advnorm=True):
    from xuance_b200.common import DummyOnPolicyBuffer, DummyOnPolicyBuffer_Atari, Box, Discrete
    cls = DummyOnPolicyBuffer_Atari if atari else DummyOnPolicyBuffer
    space = Box(0, 255, obs_shape, np.uint8) if atari else Box(-10, 10, obs_shape, np.float32)
    prod = cls(space, Discrete(4), {"old_logp": ()}, N, T, use_gae=use_gae, use_advnorm=use_advnorm, device="cuda:0")
    orc = OnPolicyBufferOracle(obs_shape, (), {"old_logp": ()}, N, T, use_gae=use_gae, use_advnorm=use_advnorm,
                               obs_dtype=np.uint8 if atari else np.float32)
    return prod, orc


@pytest.mark.parametrize("N,T,obs_shape,atari", [(8, 16, (84, 84, 4), True), (5, 37, (4,), False),
                                                  (3, 128, (12, 12, 4), True), (2, 300, (17,), False)])
@pytest.mark.parametrize("use_gae", [True, False])
def test_store_gae_sample_match_oracle(N, T, obs_shape, atari, use_gae):
    rng = np.random.default_rng(N * 1000 + T)
    ro = synth_rollout(rng, N, T, obs_shape, obs_dtype=np.uint8 if atari else np.float32, p_term=0.05)
    prod, orc = _mk(N, T, obs_shape, atari, use_gae=use_gae)
    mids = [(T // 3, 1 % N, np.float32(0.37)), (T // 2, 0, 0.0), (T // 2 + 1, 0, np.float32(-1.5))]
    fill_buffers([prod, orc], ro, mids)
    # stored bytes: bit-exact
    assert np.array_equal(prod.observations.cpu().numpy(), orc.observations)
    for name in ("actions", "rewards", "values", "terminals"):
        assert np.array_equal(getattr(prod, name).cpu().numpy(), getattr(orc, name)), name
    assert np.array_equal(prod.auxiliary_infos["old_logp"].cpu().numpy(), orc.aux["old_logp"])
    assert prod.ptr == orc.ptr and prod.size == orc.size and np.array_equal(prod.start_ids, orc.start_ids)
    # GAE / returns: fp32 tolerance
    np.testing.assert_allclose(prod.advantages.cpu().numpy(), orc.advantages, **GAE_TOL)
    np.testing.assert_allclose(prod.returns.cpu().numpy(), orc.returns, **GAE_TOL)
    # sample: gathers bit-exact w.r.t. the product's own arrays, values within tolerance of the oracle
    idx = rng.permutation(N * T)[: max(4, (N * T) // 2)]
    sp, so = prod.sample(idx), orc.sample(idx)
    assert sp["batch_size"] == so["batch_size"]
    assert np.array_equal(sp["obs"].cpu().numpy(), so["obs"])
    assert np.array_equal(sp["actions"].cpu().numpy(), so["actions"])
    assert np.array_equal(sp["values"].cpu().numpy(), so["values"])
    assert np.array_equal(sp["aux_batch"]["old_logp"].cpu().numpy(), so["aux_batch"]["old_logp"])
    np.testing.assert_allclose(sp["returns"].cpu().numpy(), so["returns"], **GAE_TOL)
    np.testing.assert_allclose(sp["advantages"].cpu().numpy(), so["advantages"], rtol=2e-5, atol=5e-6)
    for k in ("obs", "actions", "returns", "values", "advantages"):
        assert tuple(sp[k].shape) == so[k].shape and str(sp[k].dtype).split(".")[-1] == str(so[k].dtype), k


def test_unfinished_steps_are_zero_and_clear_resets():
    rng = np.random.default_rng(3)
    N, T = 4, 20
    ro = synth_rollout(rng, N, T, (4,), obs_dtype=np.float32)
    prod, orc = _mk(N, T, (4,), False)
    for t in range(T):
        for b in (prod, orc):
            b.store(ro["obs"][t], ro["acts"][t], ro["rews"][t], ro["vals"][t], ro["terms"][t], {"old_logp": ro["logp"][t]})
        if t == 7:
            for b in (prod, orc):
                b.finish_path(0.25, 2)
    for b in (prod, orc):   # only env 0 and env 2 finish at the end
        b.finish_path(1.0, 0)
        b.finish_path(0.0, 2)
    np.testing.assert_allclose(prod.advantages.cpu().numpy(), orc.advantages, **GAE_TOL)
    np.testing.assert_allclose(prod.returns.cpu().numpy(), orc.returns, **GAE_TOL)
    assert float(prod.advantages[1].abs().sum()) == 0.0 and float(prod.returns[3].abs().sum()) == 0.0
    prod.clear(), orc.clear()
    assert prod.ptr == 0 and prod.size == 0 and not prod.full
    assert float(prod.returns.abs().sum()) == 0.0


@pytest.mark.parametrize("fmt_name", ["F32_NHWC", "F32_NCHW", "BF16_NHWC", "F16_NHWC"])
def test_gather_obs_formats(fmt_name):
    from xuance_b200 import _lib
    rng = np.random.default_rng(11)
    S, B, H, W, C = 300, 257, 84, 84, 4
    src = rng.integers(0, 256, size=(S, H, W, C), dtype=np.uint8)
    src[0].flat[:256] = np.arange(256, dtype=np.uint8)   # every byte value appears
    idx = rng.integers(0, S, size=B)
    idx[0] = 0
    fmt = getattr(_lib, "OBS_" + fmt_name)
    dt = {"F32_NHWC": torch.float32, "F32_NCHW": torch.float32, "BF16_NHWC": torch.bfloat16, "F16_NHWC": torch.float16}[fmt_name]
    shape = (B, C, H, W) if fmt_name == "F32_NCHW" else (B, H, W, C)
    out = torch.empty(shape, dtype=dt, device="cuda:0")
    d_src, d_idx = torch.from_numpy(src).cuda(), torch.from_numpy(idx).cuda()
    _lib.call("xb_gather_obs", _lib.ptr(d_src), _lib.ptr(d_idx), B, H, W, C, _lib.ptr(out), fmt)
    ref = (torch.from_numpy(src[idx]) / 255.0).to(torch.float32)      # the reference's x / 255.0 (cnn.py:99)
    if fmt_name == "F32_NCHW":
        ref = ref.permute(0, 3, 1, 2).contiguous()
    ref = ref.to(dt)
    assert torch.equal(out.cpu(), ref), "u8/255 conversion must be bit-exact (correctly rounded division)"
    # identity index (idx == NULL) = the encoder's preprocessing path
    out2 = torch.empty((S,) + shape[1:], dtype=dt, device="cuda:0")
    _lib.call("xb_gather_obs", _lib.ptr(d_src), None, S, H, W, C, _lib.ptr(out2), fmt)
    ref2 = (torch.from_numpy(src) / 255.0).to(torch.float32)
    if fmt_name == "F32_NCHW":
        ref2 = ref2.permute(0, 3, 1, 2).contiguous()
    assert torch.equal(out2.cpu(), ref2.to(dt))


@pytest.mark.parametrize("row_bytes,B", [(28224, 1000), (16, 33), (68, 100), (4096, 5), (70000 * 16, 7), (2048, 1)])
def test_gather_rows_bit_exact(row_bytes, B):
    from xuance_b200 import _lib
    rng = np.random.default_rng(row_bytes)
    S = 50
    src = rng.integers(0, 256, size=(S, row_bytes), dtype=np.uint8)
    idx = rng.integers(0, S, size=B)
    d_src, d_idx = torch.from_numpy(src).cuda(), torch.from_numpy(idx).cuda()
    out = torch.zeros((B, row_bytes), dtype=torch.uint8, device="cuda:0")
    _lib.call("xb_gather_rows", _lib.ptr(d_src), _lib.ptr(d_idx), B, row_bytes, _lib.ptr(out))
    assert np.array_equal(out.cpu().numpy(), src[idx])


def test_full_size_rollout_properties():
    """BASELINE config 2 (256 envs x 128 steps, 84x84x4 u8): size-independent checks - the gather of a permutation is
    a permutation of the rows (checksum), normalised advantages have mean 0 / std 1, GAE matches the oracle."""
    from xuance_b200.common import DummyOnPolicyBuffer_Atari, Box, Discrete
    from oracle.onpolicy import gae_segment
    N, T = 256, 128
    g = torch.Generator(device="cuda:0").manual_seed(0)
    buf = DummyOnPolicyBuffer_Atari(Box(0, 255, (84, 84, 4), np.uint8), Discrete(4), {"old_logp": ()}, N, T, device="cuda:0")
    rng = np.random.default_rng(0)
    ro = synth_rollout(rng, N, T, (1, 1, 4))   # scalars from numpy; frames generated on device
    row_sums = torch.zeros((N, T), dtype=torch.int64, device="cuda:0")
    for t in range(T):
        obs = torch.randint(0, 256, (N, 84, 84, 4), dtype=torch.uint8, device="cuda:0", generator=g)
        row_sums[:, t] = obs.reshape(N, -1).sum(1, dtype=torch.int64)
        buf.store(obs, torch.from_numpy(ro["acts"][t]).cuda(), torch.from_numpy(ro["rews"][t]).cuda(),
                  torch.from_numpy(ro["vals"][t]).cuda(), torch.from_numpy(ro["terms"][t]).cuda(),
                  {"old_logp": torch.from_numpy(ro["logp"][t]).cuda()})
    for i in range(N):
        buf.finish_path(0.0 if ro["terms"][T - 1, i] else ro["boot"][i], i)
    adv = buf.advantages.cpu().numpy()
    for i in (0, 17, 255):
        a, r = gae_segment(ro["rews"][:, i], ro["vals"][:, i], ro["terms"][:, i].astype(np.float32),
                           0.0 if ro["terms"][T - 1, i] else ro["boot"][i], 0.99, 0.95)
        np.testing.assert_allclose(adv[i], a, **GAE_TOL)
    perm = np.random.default_rng(1).permutation(N * T)
    total = 0
    for k in range(4):
        s = buf.sample(perm[k * 8192:(k + 1) * 8192])
        total += int(s["obs"].reshape(8192, -1).sum(1, dtype=torch.int64).sum())
        got = s["obs"].reshape(8192, -1).sum(1, dtype=torch.int64)
        want = row_sums.reshape(-1)[torch.from_numpy(perm[k * 8192:(k + 1) * 8192]).cuda()]
        assert torch.equal(got, want)
        a = s["advantages"].double()
        assert abs(float(a.mean())) < 1e-5 and abs(float(a.std(unbiased=False)) - 1.0) < 1e-4
    assert total == int(row_sums.sum())

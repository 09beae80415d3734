"""Host logic of K12's padded-row / TMA-box / halo modes (CPU): the geometry tables BoxNatureCNN builds are executed by a NumPy
restatement of what include/xb200.h says xb_gemm_box_tc / xb_gemm_halo_tc compute, and compared with torch convolutions and
their autograd.  The CUDA kernels are pinned against the same float64 references in tests/test_gpu_tc_conv.py; this file pins
the tables themselves (chunk order, pixel-pair view, row phases, shifts, valid extents, output placement)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from xuance_b200.torch.utils import tc_conv as tc


def _plan(B):
    torch.manual_seed(3)
    convs = [nn.Conv2d(4, 32, 8, 4, padding=2), nn.Conv2d(32, 64, 4, 2, padding=1), nn.Conv2d(64, 64, 3, 1, padding=1)]
    enc = tc.BoxNatureCNN(convs, None, (84, 84, 4), backend=tc.CudaBackend(3))
    assert enc._box_ok()
    return convs, enc._plan(B)


def box_gemm(x, w, bg, out, N, relu=False, mask=None):
    """xb_gemm_box_tc in NumPy.  x [B*hp_in, W, C] float64, w [N, n_chunks*64]; writes into out [B*out_H, out_W, N]."""
    rows, W, C = x.shape
    for R in range(bg.B * bg.hp_out):
        b, y = divmod(R, bg.hp_out)
        if not (bg.y0 <= y <= bg.y1):
            continue
        for xs in range(bg.box_px):
            a = np.zeros(64 * len(bg.chunks))
            for i, (c0, w0, r0) in enumerate(bg.chunks):
                row, px = R * bg.row_step + r0, w0 + xs
                if 0 <= row < rows and 0 <= px < W:
                    a[64 * i:64 * i + 64] = x[row, px, c0:c0 + 64]
            v = w @ a
            if relu:
                v = np.maximum(v, 0.0)
            orow, ocol = b * bg.out_H + (y - bg.y0) * bg.oys + bg.oy0, xs * bg.oxs + bg.ox0
            if mask is not None:
                v = v * (mask[orow, ocol + bg.mask_x0] > 0)
            out[orow, ocol] = v


def halo_gemm(x, w, hg, out, mask=None):
    """xb_gemm_halo_tc in NumPy.  x [B*hp, W, 64]; w [n_sub*N, n_chunks*64]; out [B*out_H, out_W, N (same_cols) or n_sub*N]."""
    rows, W, _ = x.shape
    for R in range(hg.B * hg.hp):
        b, y = divmod(R, hg.hp)
        for hc in range(hg.halo_w):
            px = hc + hg.halo_w0
            for j, sb in enumerate(hg.subs):
                if not (hg.y0 <= y <= sb["y1"] and 0 <= px <= sb["x1"]):
                    continue
                a = np.zeros(64 * hg.n_chunks)
                for i, (dr, dc) in enumerate(sb["shifts"]):
                    assert hg.halo_w0 <= px + dc <= hg.halo_w0 + hg.halo_w - 1, "a tap of a valid site leaves its raster row"
                    if 0 <= R + dr < rows and 0 <= px + dc < W:
                        a[64 * i:64 * i + 64] = x[R + dr, px + dc]
                v = w[j * hg.N:(j + 1) * hg.N] @ a
                orow, ocol = b * hg.out_H + (y - hg.y0) * hg.oys + sb["oy0"], px * hg.oxs + sb["ox0"]
                if mask is not None:
                    v = v * (mask[orow, ocol + hg.mask_x0] > 0)
                c0 = 0 if hg.same_cols else j * hg.N
                out[orow, ocol, c0:c0 + hg.N] = v


def packed(wt):
    """torch conv weight [N, C, KH, KW] -> [N, (kh, kw, c)] (xb_pack_conv_weight order)."""
    return wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1).double().numpy()


def test_box_and_halo_tables_compute_the_convolutions():
    B = 2
    convs, P = _plan(B)
    hp1, hp2, off1, W1p, xo1 = P["hp1"], P["hp2"], P["off1"], P["W1p"], P["xo1"]
    w2, b2, w3 = convs[1].weight.detach(), convs[1].bias.detach(), convs[2].weight.detach()
    rng = np.random.default_rng(0)
    a1 = rng.standard_normal((B, 21, 21, 32))
    act1 = np.zeros((B, hp1, W1p, 32))
    act1[:, off1:off1 + 21, xo1:xo1 + 21] = a1
    # conv2 forward: TMA boxes over the pixel-pair view, row phases
    pairs = act1.reshape(B * hp1, W1p // 2, 64)
    out2 = np.zeros((B * hp2, 10, 64))
    box_gemm(pairs, packed(w2), P["fwd2"], out2, 64)
    t = lambda a: torch.from_numpy(a).permute(0, 3, 1, 2)
    want2 = F.conv2d(t(a1), w2.double(), None, stride=2, padding=1).permute(0, 2, 3, 1).numpy()
    got2 = out2.reshape(B, hp2, 10, 64)
    np.testing.assert_allclose(got2[:, 1:11], want2, atol=1e-10)
    assert np.abs(got2[:, 0]).max() == 0 and np.abs(got2[:, 11]).max() == 0
    # conv3 forward: box and halo tables
    a2 = rng.standard_normal((B, 10, 10, 64))
    act2 = np.zeros((B, hp2, 10, 64))
    act2[:, 1:11] = a2
    act2 = act2.reshape(B * hp2, 10, 64)
    want3 = F.conv2d(t(a2), w3.double(), None, stride=1, padding=1).permute(0, 2, 3, 1).numpy()
    for fn, geo in ((box_gemm, P["fwd3"]), (halo_gemm, P["h_fwd3"])):
        out3 = np.zeros((B * 10, 10, 64))
        fn(act2, packed(w3), geo, out3) if fn is halo_gemm else fn(act2, packed(w3), geo, out3, 64)
        np.testing.assert_allclose(out3.reshape(B, 10, 10, 64), want3, atol=1e-10)
    # conv3 data gradient (flipped taps), box and halo, with the ReLU mask of act2
    g3 = rng.standard_normal((B, 10, 10, 64))
    g3p = np.zeros((B, hp2, 10, 64))
    g3p[:, 1:11] = g3
    g3p = g3p.reshape(B * hp2, 10, 64)
    x2 = t(a2).requires_grad_(True)
    (dx2,) = torch.autograd.grad(F.conv2d(x2, w3.double(), stride=1, padding=1), x2, t(g3))
    mask2 = act2.reshape(B * hp2, 10, 64)
    want_d2 = dx2.permute(0, 2, 3, 1).numpy() * (a2 > 0)
    wd3 = tc.dgrad_weight_matrix(w3, P["taps3"]).double().numpy()
    for fn, geo in ((box_gemm, P["dg3"]), (halo_gemm, P["h_dg3"])):
        d2 = np.zeros((B * hp2, 10, 64))
        fn(g3p, wd3, geo, d2, mask=mask2) if fn is halo_gemm else fn(g3p, wd3, geo, d2, 64, mask=mask2)
        np.testing.assert_allclose(d2.reshape(B, hp2, 10, 64)[:, 1:11], want_d2, atol=1e-10)
    # conv2 data gradient: four stride phases (four box launches / one halo launch), mask = act1 in its wider layout
    g2 = rng.standard_normal((B, 10, 10, 64))
    g2p = np.zeros((B, hp2, 10, 64))
    g2p[:, 1:11] = g2
    g2p = g2p.reshape(B * hp2, 10, 64)
    x1 = t(a1).requires_grad_(True)
    (dx1,) = torch.autograd.grad(F.conv2d(x1, w2.double(), stride=2, padding=1), x1, t(g2))
    want_d1 = dx1.permute(0, 2, 3, 1).numpy() * (a1 > 0)
    mask1 = act1.reshape(B * hp1, W1p, 32)
    d1 = np.zeros((B * hp1, 21, 32))
    for bg, taps in P["dg2"]:
        box_gemm(g2p, tc.dgrad_weight_matrix(w2, taps).double().numpy(), bg, d1, 32, mask=mask1)
    np.testing.assert_allclose(d1.reshape(B, hp1, 21, 32)[:, off1:off1 + 21], want_d1, atol=1e-10)
    d1h = np.zeros((B * hp1, 21, 32))
    wd2_all = np.concatenate([tc.dgrad_weight_matrix(w2, taps).double().numpy() for _, taps in P["dg2"]], 0)
    halo_gemm(g2p, wd2_all, P["h_dg2"], d1h, mask=mask1)
    np.testing.assert_allclose(d1h.reshape(B, hp1, 21, 32)[:, off1:off1 + 21], want_d1, atol=1e-10)
    assert np.abs(d1h.reshape(B, hp1, 21, 32)[:, :off1]).max() == 0


def test_halo_tile_holds_every_tap_of_its_128_positions():
    """The resident tile of M tile t starts at row floor((128 t + lo) / halo_w) and has the rows xb_gemm_halo_tc allocates:
    every (position + shift) of the tile must fall inside it (the kernel reads it through a descriptor offset)."""
    _, P = _plan(5)
    for key in ("h_fwd3", "h_dg3", "h_dg2"):
        hg = P[key]
        shifts = [dr * hg.halo_w + dc for sb in hg.subs for dr, dc in sb["shifts"]]
        lo, hi = min(0, min(shifts)), max(0, max(shifts))
        halo_rows = (hg.halo_w - 1 + 128 + hi - lo + hg.halo_w - 1) // hg.halo_w
        assert halo_rows * hg.halo_w * 128 * 3 * 2 + 2 * 3 * hg.N * 128 <= 212 * 1024        # two tiles + two ring stages fit
        for tile in range(-(-hg.M // 128)):
            r_lo = (tile * 128 + lo) // hg.halo_w
            s0 = tile * 128 - r_lo * hg.halo_w
            assert 0 <= s0 + lo and s0 + hi + 127 < halo_rows * hg.halo_w, (key, tile)

// Index arithmetic of the gathered-operand tensor-core GEMM (conv_tc.cu), kept free of CUDA so that the SAME functions
// are compiled into the kernel and into the host-side layout test (tests/csrc/conv_index_test.cpp, run by
// tests/test_conv_index.py): the test stages operands with these functions into an emulated shared-memory image, reads
// them back the way the tcgen05 shared-memory descriptors do (validated on hardware by K9-TC) and compares with a direct
// convolution.
//
// The GEMM:  D[m, n] = sum_k A[m, k] * W[n, k]
//   m  <-> one site (b, y, x) of an output grid [B, OY, OX]
//   k  <-> (tap t, channel c), k = t*C + c, C a multiple of 8
//   A[m, (t, c)] = in[b, y*sy + dy[t], x*sx + dx[t], c]     (zero outside [0,IH) x [0,IW))    NHWC input
// which covers
//   * forward convolution:      dy[t] = kh - pad, dx[t] = kw - pad, (sy, sx) = stride              (cnn.py:45-50)
//   * a fully connected layer:  one tap, IH = IW = OY = OX = 1, C = in_features
//   * data gradient, stride 1:  the flipped taps dy[t] = pad - kh over the output-gradient tensor
//   * data gradient, stride s:  one such GEMM per phase (y mod s, x mod s) with that phase's taps, sy = sx = 1
// Operands travel as TWO bf16 tensors each (x = hi + lo, see tc_common.cuh::split_bf16).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define XB_HD __host__ __device__ __forceinline__
#else
#define XB_HD inline
#endif

#define XB_CONV_MAX_TAPS 64
#define XB_CONV_KC 64           // K elements per pipeline stage (8 core matrices, 128 B per operand row)
#define XB_CONV_TILE_M 128      // rows per CTA tile = TMEM lanes

struct XbConvGeom {
    int B, IH, IW, C;           // input tensor [B, IH, IW, C] (bf16 hi / lo planes), C % 8 == 0
    int OY, OX;                 // output grid per image
    int sy, sx;                 // input step per output step
    int T;                      // taps; K = T * C, K % XB_CONV_KC == 0
    int N;                      // output channels (GEMM N), N % 16 == 0, N <= 256
    int8_t dy[XB_CONV_MAX_TAPS], dx[XB_CONV_MAX_TAPS];
};

// byte offset of element (r, k) of a [rows x KP] bf16 operand in the K-major no-swizzle canonical layout
// (8-row x 16-byte core matrices; K-adjacent cores 128 B apart; 8-row groups KP/8*128 B apart)
XB_HD uint32_t xb_canon_off(int r, int k, int KP) {
    return (uint32_t)((r >> 3) * (KP >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// byte offset of element (mn, k) of a [MN x KP] bf16 operand in the MN-major no-swizzle canonical layout: a core matrix
// is 8 k-rows of 16 bytes (8 consecutive mn elements); K-adjacent cores 128 B apart (descriptor LBO), 8-mn groups
// KP/8*128 B apart (descriptor SBO) - the transpose of xb_canon_off inside each core matrix, same core placement.
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>, LayoutType::INTERLEAVE: ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)))
XB_HD uint32_t xb_canon_off_mn(int mn, int k, int KP) {
    return (uint32_t)((mn >> 3) * (KP >> 3) * 128 + (k >> 3) * 128 + (k & 7) * 16 + (mn & 7) * 2);
}

// site m -> (b, y, x)
XB_HD void xb_conv_site(const XbConvGeom &g, int64_t m, int &b, int &y, int &x) {
    const int per_img = g.OY * g.OX;
    b = (int)(m / per_img);
    const int rem = (int)(m - (int64_t)b * per_img);
    y = rem / g.OX;
    x = rem - y * g.OX;
}

// element offset (in bf16 elements, into the NHWC input) of the 8-channel unit starting at GEMM column k0 (k0 % 8 == 0)
// for site (b, y, x); returns -1 when the tap falls outside the image (the unit is zero-filled)
XB_HD int64_t xb_conv_unit_src(const XbConvGeom &g, int b, int y, int x, int k0) {
    const int t = k0 / g.C;
    const int c = k0 - t * g.C;
    const int iy = y * g.sy + g.dy[t];
    const int ix = x * g.sx + g.dx[t];
    if (iy < 0 || iy >= g.IH || ix < 0 || ix >= g.IW) return -1;
    return (((int64_t)b * g.IH + iy) * g.IW + ix) * g.C + c;
}

// packed weight [N, (kh, kw, c)] element i  <-  torch weight [N, C, KH, KW] element xb_pack_weight_src(i, ...)
XB_HD int64_t xb_pack_weight_src(int64_t i, int C, int KH, int KW) {
    const int64_t K = (int64_t)C * KH * KW;
    const int64_t n = i / K;
    int64_t r = i - n * K;
    const int c = (int)(r % C);
    r /= C;
    const int kw = (int)(r % KW), kh = (int)(r / KW);
    return ((n * C + c) * KH + kh) * KW + kw;
}

// sites per split of the weight-gradient reduction: a multiple of the chunk length; 0 if `splits` would leave one empty
XB_HD int64_t xb_wgrad_sites_per_split(int64_t M, int splits) {
    int64_t per = (M + splits - 1) / splits;
    per = (per + XB_CONV_KC - 1) / XB_CONV_KC * XB_CONV_KC;
    return ((int64_t)(splits - 1) * per >= M) ? 0 : per;
}

// ---------------------------------------------------------------------------------------------------------------------
// What ONE producer thread (`row` in [0,128)) stages for one K chunk.  emit_a / emit_w receive (byte offset of the 16-byte
// unit inside the stage's hi plane of that operand, element offset of its 8 source values or -1 for a zero unit).
// The kernel turns each call into two cp.async (hi and lo plane); the host test writes an emulated shared memory.
// ---------------------------------------------------------------------------------------------------------------------
// forward / data-gradient GEMM: thread = site (b, y, x) (live = inside the problem); K-major operands
template <class EmitA, class EmitW>
XB_HD void xb_stage_fwd(const XbConvGeom &g, int row, bool live, int b, int y, int x, int kc, EmitA &&emit_a,
                        EmitW &&emit_w) {
    const int K = g.T * g.C;
    for (int u = 0; u < XB_CONV_KC / 8; ++u) {
        const int64_t off = live ? xb_conv_unit_src(g, b, y, x, kc * XB_CONV_KC + u * 8) : -1;
        emit_a(xb_canon_off(row, u * 8, XB_CONV_KC), off);
    }
    for (int idx = row; idx < g.N * (XB_CONV_KC / 8); idx += XB_CONV_TILE_M) {
        const int n = idx >> 3, u = idx & 7;
        emit_w(xb_canon_off(n, u * 8, XB_CONV_KC), (int64_t)n * K + kc * XB_CONV_KC + u * 8);
    }
}

// weight-gradient GEMM: the chunk holds 64 consecutive sites starting at chunk_site0 (sites >= site_end are zero);
// thread = (site pp = row & 63, half uh = row >> 6 of the tile's 16 column units); MN-major operands;
// the second operand is the output gradient G[site, g_c0 + n] of a matrix with g_ld elements per site (g.N = tile width)
template <class EmitA, class EmitG>
XB_HD void xb_stage_wgrad(const XbConvGeom &g, int row, int64_t mt, int64_t chunk_site0, int64_t site_end, int64_t g_ld,
                          int g_c0, EmitA &&emit_a, EmitG &&emit_g) {
    const int K = g.T * g.C;
    const int pp = row & (XB_CONV_KC - 1), uh = row >> 6;
    const int64_t site = chunk_site0 + pp;
    const bool in_run = site < site_end;
    int b = 0, y = 0, x = 0;
    if (in_run) xb_conv_site(g, site, b, y, x);
    for (int j = 0; j < 8; ++j) {
        const int u = uh * 8 + j;
        const int64_t kcol = mt * XB_CONV_TILE_M + u * 8;
        const int64_t off = (in_run && kcol < K) ? xb_conv_unit_src(g, b, y, x, (int)kcol) : -1;
        emit_a(xb_canon_off_mn(u * 8, pp, XB_CONV_KC), off);
    }
    for (int j = uh; j < g.N / 8; j += 2)
        emit_g(xb_canon_off_mn(j * 8, pp, XB_CONV_KC), in_run ? site * g_ld + g_c0 + j * 8 : (int64_t)-1);
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-coalesced producer mapping (MAP = 1; same shared-memory image, different division of the units among the 128
// producer threads).  In the canonical no-swizzle layouts the shared-memory bank of a 16-byte unit depends only on
// (row & 7) [K-major] / (site & 7) [MN-major], so a warp instruction is conflict-free as soon as its 32 lanes hold each of
// the 8 residues four times - which leaves the other lane bits free to walk along MEMORY-contiguous units: lane ->
// (residue = lane & 7, unit offset = lane >> 3), i.e. 8 rows x 4 consecutive units = 8 x 64 contiguous bytes per
// instruction instead of 32 scattered 16-byte pieces (half of every 32-byte sector wasted).
// ---------------------------------------------------------------------------------------------------------------------
// forward: producer thread t = (warp w, lane l) owns rows 32w + 8*gi + (l & 7), gi < 4, and units (l >> 3) and 4 + (l >> 3)
XB_HD int xb_v2_row(int t, int gi) { return 32 * (t >> 5) + 8 * gi + (t & 7); }

// sites: the thread's four sites as {b, y, x} (b < 0: row outside the problem)
template <class EmitA, class EmitW>
XB_HD void xb_stage_fwd_v2(const XbConvGeom &g, int t, const int (*sites)[3], int kc, EmitA &&emit_a, EmitW &&emit_w) {
    const int K = g.T * g.C;
    const int w = t >> 5, l = t & 31, rs = l & 7, uo = l >> 3;
    for (int half = 0; half < 2; ++half) {
        const int u = half * 4 + uo, k0 = kc * XB_CONV_KC + u * 8;
        for (int gi = 0; gi < 4; ++gi) {
            const int *s = sites[gi];
            const int64_t off = s[0] >= 0 ? xb_conv_unit_src(g, s[0], s[1], s[2], k0) : -1;
            emit_a(xb_canon_off(32 * w + 8 * gi + rs, u * 8, XB_CONV_KC), off);
        }
    }
    for (int q = w; q < (g.N >> 3) * 2; q += 4) {            // (8-row group, half) pairs of the weight tile, dealt to the warps
        const int n = (q >> 1) * 8 + rs, u = (q & 1) * 4 + uo;
        emit_w(xb_canon_off(n, u * 8, XB_CONV_KC), (int64_t)n * K + kc * XB_CONV_KC + u * 8);
    }
}

// weight gradient: thread t owns sites 16w + 8*si + (l & 7), si < 2, of the chunk and the column units 4*quad + (l >> 3)
template <class EmitA, class EmitG>
XB_HD void xb_stage_wgrad_v2(const XbConvGeom &g, int t, int64_t mt, int64_t chunk_site0, int64_t site_end, int64_t g_ld,
                             int g_c0, EmitA &&emit_a, EmitG &&emit_g) {
    const int K = g.T * g.C;
    const int w = t >> 5, l = t & 31, rs = l & 7, uo = l >> 3;
    for (int si = 0; si < 2; ++si) {
        const int pp = 16 * w + 8 * si + rs;
        const int64_t site = chunk_site0 + pp;
        const bool in_run = site < site_end;
        int b = 0, y = 0, x = 0;
        if (in_run) xb_conv_site(g, site, b, y, x);
        for (int quad = 0; quad < 4; ++quad) {
            const int u = quad * 4 + uo;
            const int64_t kcol = mt * XB_CONV_TILE_M + u * 8;
            const int64_t off = (in_run && kcol < K) ? xb_conv_unit_src(g, b, y, x, (int)kcol) : -1;
            emit_a(xb_canon_off_mn(u * 8, pp, XB_CONV_KC), off);
        }
        for (int j = uo; j < (g.N >> 3); j += 4)
            emit_g(xb_canon_off_mn(j * 8, pp, XB_CONV_KC), in_run ? site * g_ld + g_c0 + j * 8 : (int64_t)-1);
    }
}

// shared-memory bytes of one pipeline stage: the P planes of A (hi | [mid |] lo) followed by the P planes of W.
// P = 2: x = hi + lo (|err| <= 2^-16 |x|, 3 products);  P = 3: x = hi + mid + lo (<= 2^-24 |x|, 6 products)
XB_HD uint32_t xb_conv_stage_bytes(int N, int P = 2) {
    return (uint32_t)(P * XB_CONV_TILE_M * XB_CONV_KC * 2 + P * N * XB_CONV_KC * 2);
}
XB_HD uint32_t xb_conv_a_plane_bytes() { return (uint32_t)(XB_CONV_TILE_M * XB_CONV_KC * 2); }
XB_HD uint32_t xb_conv_w_plane_bytes(int N) { return (uint32_t)(N * XB_CONV_KC * 2); }

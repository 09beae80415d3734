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

// shared-memory bytes of one pipeline stage: A hi | A lo | W hi | W lo
XB_HD uint32_t xb_conv_stage_bytes(int N) { return (uint32_t)(2 * XB_CONV_TILE_M * XB_CONV_KC * 2 + 2 * N * XB_CONV_KC * 2); }
XB_HD uint32_t xb_conv_a_plane_bytes() { return (uint32_t)(XB_CONV_TILE_M * XB_CONV_KC * 2); }
XB_HD uint32_t xb_conv_w_plane_bytes(int N) { return (uint32_t)(N * XB_CONV_KC * 2); }

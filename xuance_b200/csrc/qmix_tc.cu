// K9-TC: the QMIX mixing network forward as ONE tensor-core kernel (tcgen05.mma, accumulators in TMEM).
//
//   Z      = X[rows x S] . W1cat^T[S x 128]              (first layers of hyper_w_1 | hyper_b_1 | hyper_w_2 | hyper_b_2)
//   W1raw  = relu(Z[:,  0: 32] + c) . Wb1^T [32 x n*32]  (second layer of hyper_w_1)
//   W2raw  = relu(Z[:, 64: 96] + c) . Wb2^T [32 x 32]    (second layer of hyper_w_2)
//   b2     = relu(Z[:, 96:128] + c) . wb2^T [32 x 1]     (second layer of hyper_b_2)
//   q_tot  = elu(q . |W1raw + c| + (Z[:,32:64] + c)) . |W2raw + c| + b2 + c         per row
// (xuance/torch/rl_models/heads/q_mix_head.py:52-95).  One CTA (128 threads) owns a tile of 128 rows: M = 128,
// thread i <-> TMEM lane i <-> row i.  All GEMM operands sit in shared memory in the K-major no-swizzle canonical
// layout (8-row x 16-byte core matrices), written by the threads themselves; the weights are staged once per CTA,
// the CTA then loops over its tiles (persistent grid = SM count).
//
// Numerics: the reference computes these layers in fp32.  The tensor pipe takes bf16 here, so every fp32 operand is
// split x = hi + lo (two bf16 values, |x - hi - lo| <= 2^-16 |x|) and each product is the sum of the three MMAs
// hi.hi + hi.lo + lo.hi (lo.lo <= 2^-16 relative, the size of the split residual itself) accumulated in fp32 in TMEM: relative error ~1e-5 per layer (a plain bf16 or TF32
// pass would be ~4e-3 / ~5e-4).  tests/test_gpu_qmix.py pins it against the fp32 oracle.
#include "tc_common.cuh"

namespace {
using namespace xbtc;

constexpr int TC_ROWS = 128;   // M
constexpr int TC_HH = 32;      // hypernet hidden = mixing hidden = 32 (the shipped QMIX configs)
constexpr int TC_N1 = 128;     // 4 x 32 first-layer outputs
constexpr int TC_TMEM_COLS = 512;
constexpr int COL_Z = 0, COL_W1 = 128, COL_W2 = 384, COL_B2 = 416;   // W1raw may take up to 256 columns (n <= 8)

// stage a [rows x K] fp32 row-major matrix (leading dimension ld; global OR shared memory) as hi/lo bf16 canonical
// operands, zero padded; one thread-iteration = 8 consecutive k of one row (two 16-byte shared-memory stores);
// U units are loaded before the first is used; (r, k0) advance incrementally (no division in the loop)
__device__ __forceinline__ void stage_operand(const float *__restrict__ g, int rows_valid, int rows_total, int K, int ld,
                                              int KP, uint8_t *hi, uint8_t *lo) {
    const int upr = KP >> 3;  // units per row
    const int total = rows_total * upr;
    constexpr int U = 4;
    const int dr = blockDim.x / upr, dk = blockDim.x - dr * upr;   // unit index += blockDim.x  <=>  (r += dr, ku += dk)
    int r = threadIdx.x / upr, ku = threadIdx.x - r * upr;
    for (int e0 = threadIdx.x; e0 < total; e0 += blockDim.x * U) {
        float x[U][8];
        int rr[U], kk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            rr[u] = r, kk[u] = ku << 3;
            const float *src = g + (int64_t)r * ld + kk[u];
            const bool in = (e0 + u * (int)blockDim.x) < total && r < rows_valid;
#pragma unroll
            for (int i = 0; i < 8; ++i) x[u][i] = (in && kk[u] + i < K) ? src[i] : 0.f;
            r += dr, ku += dk;
            if (ku >= upr) ku -= upr, ++r;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (e0 + u * (int)blockDim.x < total) store_unit(x[u], rr[u], kk[u], KP, hi, lo);
    }
}

// issue the three split products of one GEMM: D[128 x N] (+)= A[128 x KP] . B[N x KP]^T
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, const uint8_t *a_hi, const uint8_t *a_lo,
                                           const uint8_t *b_hi, const uint8_t *b_lo, int KP, int N) {
    const uint32_t idesc = make_idesc(TC_ROWS, N);
    const uint8_t *as[2] = {a_hi, a_lo};
    const uint8_t *bs[2] = {b_hi, b_lo};
    uint32_t acc = 0;
    for (int pa = 0; pa < 2; ++pa)
        for (int pb = 0; pb < 2 - pa; ++pb)   // (hi,hi) (hi,lo) (lo,hi); lo.lo <= 2^-16 relative is dropped
            for (int ks = 0; ks < KP / 16; ++ks) {  // one MMA = K 16 = two core matrices = 256 B along K
                const uint64_t da = make_desc(smem_u32(as[pa]) + ks * 256, KP);
                const uint64_t db = make_desc(smem_u32(bs[pb]) + ks * 256, KP);
                mma_bf16(d_tmem, da, db, idesc, acc);
                acc = 1;
            }
}

struct TcParams {
    const float *states, *q;                 // [R,S], [R,n]
    const float *w_l1[4], *b_l1[4];          // first layers of hyper_w_1 | hyper_b_1 | hyper_w_2 | hyper_b_2: [32,S], [32]
    const float *wb1, *bias_wb1;             // [n*32,32], [n*32]
    const float *wb2, *bias_wb2;             // [32,32], [32]
    const float *wb2c, *bias_wb2c;           // [1,32], [1]
    float *q_tot;                            // [R]
    int64_t R;
    int S, n;
};

__global__ void __launch_bounds__(128, 1) qmix_mix_tc_kernel(TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mma_bar, x_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_bias1[TC_N1], s_bias_wb1[256], s_bias_wb2[TC_HH], s_bias_b2;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int S = p.S, n = p.n, KP1 = (S + 15) & ~15, N2A = n * TC_HH;
    // ---- shared memory carve-up (every operand base 128-B aligned)
    const int szA1 = TC_ROWS * KP1 * 2, szB1 = TC_N1 * KP1 * 2, szA2 = TC_ROWS * TC_HH * 2;
    const int szB2a = N2A * TC_HH * 2, szB2b = TC_HH * TC_HH * 2, szB2c = 16 * TC_HH * 2;
    uint8_t *A1h = smem, *A1l = A1h + szA1, *B1h = A1l + szA1, *B1l = B1h + szB1;
    // the three GEMM2 A operands (relu of the hidden blocks) reuse the X operand region: it is dead once GEMM1 completed
    uint8_t *A2h[3], *A2l[3];
    for (int i = 0; i < 3; ++i) {
        A2h[i] = A1h + (size_t)i * 2 * szA2;
        A2l[i] = A2h[i] + szA2;
    }
    uint8_t *B2ah = B1l + szB1, *B2al = B2ah + szB2a, *B2bh = B2al + szB2a, *B2bl = B2bh + szB2b;
    uint8_t *B2ch = B2bl + szB2b, *B2cl = B2ch + szB2c;
    float *Xraw = reinterpret_cast<float *>(B2cl + szB2c);   // [128 x S] fp32: the NEXT tile, fetched by the TMA unit
    // ---- one-time setup: mbarrier, TMEM, weights
    const int64_t n_tiles = (p.R + TC_ROWS - 1) / TC_ROWS;
    // a tile is one contiguous run of rows_valid*S*4 bytes; the bulk copy needs a multiple of 16 bytes
    auto tile_copy_bytes = [&](int64_t t) -> uint32_t {
        const int64_t rows = (p.R - t * TC_ROWS) < TC_ROWS ? (p.R - t * TC_ROWS) : TC_ROWS;
        return (uint32_t)(rows * S * 4);
    };
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        mbar_init(&x_bar, 1);
        mbar_fence_init();
        if ((int64_t)blockIdx.x < n_tiles && (tile_copy_bytes(blockIdx.x) & 15) == 0) {   // prefetch the first tile
            mbar_expect_tx(&x_bar, tile_copy_bytes(blockIdx.x));
            bulk_g2s(Xraw, p.states + (int64_t)blockIdx.x * TC_ROWS * S, tile_copy_bytes(blockIdx.x), &x_bar);
        }
    }
    __syncwarp();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int blk = 0; blk < 4; ++blk)   // rows 32*blk .. 32*blk+31 of the concatenated first-layer weight
        stage_operand(p.w_l1[blk], TC_HH, TC_HH, S, S, KP1, B1h + (size_t)blk * (TC_HH / 8) * (KP1 / 8) * 128,
                      B1l + (size_t)blk * (TC_HH / 8) * (KP1 / 8) * 128);
    stage_operand(p.wb1, N2A, N2A, TC_HH, TC_HH, TC_HH, B2ah, B2al);
    stage_operand(p.wb2, TC_HH, TC_HH, TC_HH, TC_HH, TC_HH, B2bh, B2bl);
    stage_operand(p.wb2c, 1, 16, TC_HH, TC_HH, TC_HH, B2ch, B2cl);
    for (int i = tid; i < TC_N1; i += blockDim.x) s_bias1[i] = p.b_l1[i >> 5][i & 31];
    for (int i = tid; i < N2A; i += blockDim.x) s_bias_wb1[i] = p.bias_wb1[i];
    if (tid < TC_HH) s_bias_wb2[tid] = p.bias_wb2[tid];
    if (tid == 0) s_bias_b2 = p.bias_wb2c[0];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    uint32_t phase = 0, xphase = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TC_ROWS;
        const int rows_valid = (int)((p.R - row0) < TC_ROWS ? (p.R - row0) : TC_ROWS);
        // ---- this row's agent utilities (latency hidden behind the staging + first GEMM)
        float qv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) qv[i] = (i < n && tid < rows_valid) ? p.q[(row0 + tid) * n + i] : 0.f;
        // ---- stage X tile (split to bf16 hi/lo) from the prefetched raw copy (or from global for an unaligned tail)
        if ((tile_copy_bytes(tile) & 15) == 0) {
            mbar_wait(&x_bar, xphase);
            xphase ^= 1;
            stage_operand(Xraw, rows_valid, TC_ROWS, S, S, KP1, A1h, A1l);
        } else {
            stage_operand(p.states + row0 * S, rows_valid, TC_ROWS, S, S, KP1, A1h, A1l);
        }
        fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        __syncthreads();       // also orders the previous tile's TMEM reads before this tile's MMAs
        tc_fence_after();
        if (tid == 0) {
            issue_gemm(tmem + COL_Z, A1h, A1l, B1h, B1l, KP1, TC_N1);
            mma_commit(&mma_bar);
            // Xraw was consumed by every thread (barrier above): fetch the next tile while this one computes
            const int64_t nt = tile + gridDim.x;
            if (nt < n_tiles && (tile_copy_bytes(nt) & 15) == 0) {
                mbar_expect_tx(&x_bar, tile_copy_bytes(nt));
                bulk_g2s(Xraw, p.states + nt * TC_ROWS * S, tile_copy_bytes(nt), &x_bar);
            }
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- layer-1 epilogue: thread = row.  relu(Z + bias) of the three hidden blocks -> GEMM2 A operands
        float b1v[TC_HH];
        {
            float z[32];
            const int blocks[3] = {0, 64, 96};
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                tmem_ld32(lane_addr + COL_Z + blocks[g], z);
#pragma unroll
                for (int k0 = 0; k0 < 32; k0 += 8) {
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = fmaxf(z[k0 + i] + s_bias1[blocks[g] + k0 + i], 0.f);
                    store_unit(x, tid, k0, TC_HH, A2h[g], A2l[g]);
                }
            }
            tmem_ld32(lane_addr + COL_Z + 32, z);
#pragma unroll
            for (int k = 0; k < 32; ++k) b1v[k] = z[k] + s_bias1[32 + k];
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            issue_gemm(tmem + COL_W1, A2h[0], A2l[0], B2ah, B2al, TC_HH, N2A);
            issue_gemm(tmem + COL_W2, A2h[1], A2l[1], B2bh, B2bl, TC_HH, TC_HH);
            issue_gemm(tmem + COL_B2, A2h[2], A2l[2], B2ch, B2cl, TC_HH, 16);
            mma_commit(&mma_bar);
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- mixing epilogue (q_mix_head.py:81-94), thread = row
        const bool live = tid < rows_valid;
        float pre[TC_HH];
#pragma unroll
        for (int j = 0; j < TC_HH; ++j) pre[j] = b1v[j];
        float w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < n) {
                tmem_ld32(lane_addr + COL_W1 + i * TC_HH, w);
                const float qi = qv[i];
#pragma unroll
                for (int j = 0; j < TC_HH; ++j) pre[j] += qi * fabsf(w[j] + s_bias_wb1[i * TC_HH + j]);
            }
        }
        tmem_ld32(lane_addr + COL_W2, w);
        float y = 0.f;
#pragma unroll
        for (int j = 0; j < TC_HH; ++j) {
            const float h = pre[j] > 0.f ? pre[j] : (expf(pre[j]) - 1.f);
            y += h * fabsf(w[j] + s_bias_wb2[j]);
        }
        tmem_ld32(lane_addr + COL_B2, w);
        if (live) p.q_tot[row0 + tid] = y + w[0] + s_bias_b2;
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
    }
}

}  // namespace

extern "C" int xb_qmix_mix_fused_fwd(const float *states, const float *q, const float *const *w_l1,
                                     const float *const *b_l1, const float *wb1, const float *bias_wb1,
                                     const float *wb2, const float *bias_wb2, const float *wb2c,
                                     const float *bias_wb2c, int64_t R, int S, int n, int H, int HH, float *q_tot,
                                     void *stream) {
    if (!states || !q || !w_l1 || !b_l1 || !wb1 || !bias_wb1 || !wb2 || !bias_wb2 || !wb2c || !bias_wb2c || !q_tot)
        return XB_EINVAL;
    for (int i = 0; i < 4; ++i)
        if (!w_l1[i] || !b_l1[i]) return XB_EINVAL;
    if (R <= 0 || S <= 0 || n <= 0) return XB_EINVAL;
    if (H != TC_HH || HH != TC_HH || n > 8 || S > 160) return XB_ERANGE;   // shapes of the shipped QMIX configs
    const int KP1 = (S + 15) & ~15;
    const size_t smem = (size_t)2 * TC_ROWS * KP1 * 2 + (size_t)2 * TC_N1 * KP1 * 2 + 2 * (size_t)n * TC_HH * TC_HH * 2 +
                        2 * TC_HH * TC_HH * 2 + 2 * 16 * TC_HH * 2 + (size_t)TC_ROWS * S * 4;
    if ((size_t)6 * TC_ROWS * TC_HH * 2 > (size_t)2 * TC_ROWS * KP1 * 2) return XB_ERANGE;   // A2 operands alias A1
    if (smem > 220 * 1024) return XB_ERANGE;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(qmix_mix_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr = true;
    }
    TcParams p{states, q, {w_l1[0], w_l1[1], w_l1[2], w_l1[3]}, {b_l1[0], b_l1[1], b_l1[2], b_l1[3]},
               wb1, bias_wb1, wb2, bias_wb2, wb2c, bias_wb2c, q_tot, R, S, n};
    int64_t tiles = (R + TC_ROWS - 1) / TC_ROWS;
    int grid = (int)(tiles < xb_sm_count() ? tiles : xb_sm_count());
    qmix_mix_tc_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(p);
    return xb_launch_status();
}

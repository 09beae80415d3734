// K12 (EXPERIMENTAL; nothing on the default path calls it.  Status after round 1: the forward GEMM passed on B200 -
// tests/test_gpu_tc_conv.py, profiles/r01_k12_bringup_forward.log; the data-gradient and weight-gradient modes ran on
// hardware inside the whole-encoder backward and agreed with cuDNN fp32 to ~4e-3 of max|grad|, per-layer parity pending):
// gathered-operand GEMM on the 5th-generation tensor cores for the NatureCNN layers (cnn.py:45-50, 84-101) - the 85 %
// of the fp32 PPO step that cuDNN's CUDA-core fp32 convolutions take.
//
//   D[m, n] = sum_{t, c} in[b, y*sy + dy[t], x*sx + dx[t], c] * W[n, t*C + c]  (+ bias[n], ReLU)      conv_index.h
//
// Numerics as K9-TC: every fp32 operand travels as x = hi + lo (two bf16 tensors) and each product is the three MMAs
// hi.hi + hi.lo + lo.hi accumulated in fp32 in TMEM (~1e-5 relative, i.e. inside the fp32 parity tolerance).
//
// Structure (one CTA per SM, persistent over 128-row tiles, 9 warps):
//   warps 5-8  producers : thread = tile row.  Per K chunk of 64 the row's 8 units (16 B = 8 channels of one tap) and
//                          this thread's share of the weight rows are fetched with 16-byte cp.async straight into the
//                          K-major no-swizzle canonical layout (zero-fill for padding taps / tail rows); a chunk is
//                          published (wait_group -> fence.proxy.async -> mbarrier arrive) one chunk behind the issue
//                          point so that a group is always in flight.
//   warp 4     MMA       : lane 0 waits full[stage], issues 3 x 4 tcgen05.mma (M 128, N, K 16) and commits onto
//                          empty[stage]; after the last chunk commits onto acc_full[a].  Two accumulators in TMEM.
//   warps 0-3  epilogue  : thread = row = TMEM lane.  tcgen05.ld 32 columns at a time, bias + ReLU, then the row is
//                          written as fp32 and / or as the hi / lo bf16 pair the next layer consumes.
#include <cstdlib>

#include "tc_common.cuh"
#include "conv_index.h"

namespace {
using namespace xbtc;

constexpr int KC = XB_CONV_KC, TILE_M = XB_CONV_TILE_M;
constexpr int EPI_WARPS = 4, PROD_WARPS = 4;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;
constexpr int MAX_STAGES = 6;

struct ConvParams {
    XbConvGeom g;                          // g.N = columns per work item (the tile width N)
    const __nv_bfloat16 *in[3];            // A planes: [B, IH, IW, C]
    const __nv_bfloat16 *w[3];             // B planes.  forward: weight [N_total, K].  weight gradient: output gradient [P, w_ld]
    const float *bias;                     // [N_total] or null (forward)
    const __nv_bfloat16 *mask;             // forward, nullable: result elements are zeroed where mask <= 0; same
                                           // addressing as the output (the ReLU derivative of a saved activation's hi plane)
    __nv_bfloat16 *out[3];                 // forward: p_out result planes (nullable)
    float *out_f32;                        // forward: nullable.  weight gradient: partials [splits, K, N_total]
    int64_t M;                             // sites B * OY * OX (GEMM rows forward, reduction length for the weight gradient)
    int relu, stages, p_out;
    // forward: placement of site (b, y, x): row ((b*out_H + y*oys + oy0)*out_W + x*oxs + ox0) of an output matrix whose
    // rows are out_ld elements apart; work item (m tile, n tile nt) fills columns [out_c0 + nt*N, out_c0 + (nt+1)*N)
    int out_H, out_W, oys, oxs, oy0, ox0;
    int64_t out_ld;
    int out_c0;
    int n_tiles;                           // column tiles per row tile (N_total = n_tiles * N)
    int64_t w_ld;                          // weight gradient: elements between consecutive sites of the output gradient
    // weight gradient: sites are cut into `splits` runs of sites_per_split (a multiple of KC)
    int splits;
    int64_t sites_per_split;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// x = hi + mid + lo, three bf16 values: 24 mantissa bits, |x - hi - mid - lo| <= 2^-24 |x|
__device__ __forceinline__ void split3_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &mid, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r1);
    lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}

// kind::f16 instruction descriptor with both operands MN-major (bits 15 / 16; cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N) { return make_idesc(M, N) | (1u << 15) | (1u << 16); }

// WGRAD = false:  D[site, n]  = sum_k A[site, k] W[n, k]            work item = (128-site tile, n tile),        K-major operands
// WGRAD = true :  D[kcol, n]  = sum_site A[site, kcol] G[site, n]   work item = (128-kcol tile, n tile, split), MN-major operands
// PA / PB: bf16 planes of the A / B operand (x = sum of its planes, plane q = bf16 of the residual left by planes < q),
// PA <= PB.  The products kept are those whose plane indices sum to < PB; they are accumulated in PB SEPARATE float32
// accumulators D_g, g = pa + pb, which the epilogue adds smallest first: the tensor core truncates (rounds toward zero)
// when it adds a product block into the accumulator, so keeping the 2^-8 / 2^-16 sized correction terms out of the
// hi.hi accumulator both shortens its chain of truncating additions by the number of products and keeps the corrections'
// own truncation error at 2^-8 / 2^-16 of an ulp of the result.  The B planes of a stage are adjacent in shared memory, so
// in the canonical layouts they ARE one operand of PB*N rows: ONE tcgen05.mma per A plane multiplies it with the first
// (PB - pa) B planes and lands in the accumulator columns of groups pa .. PB-1 - every A plane is read from shared memory
// once per K step instead of once per product (shared-memory operand reads, not the tensor pipe, bound N <= 64 tiles).
// MAP: how the 16-byte units of a stage are dealt to the producer threads (conv_index.h): 0 = thread per row,
// 1 = row-coalesced (8 rows x 4 memory-contiguous units per warp instruction).
template <bool WGRAD, int PA, int PB, int MAP>
__global__ void __launch_bounds__(THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
    static_assert(PA >= 1 && PA <= PB && PB <= 3, "plane counts");
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias[256];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const XbConvGeom &g = p.g;
    const int N = g.N, K = g.T * g.C, S = p.stages;
    const uint32_t a_plane = xb_conv_a_plane_bytes(), w_plane = xb_conv_w_plane_bytes(N);
    const uint32_t stage_bytes = PA * a_plane + PB * w_plane;
    const int64_t m_tiles = WGRAD ? (K + TILE_M - 1) / TILE_M : (p.M + TILE_M - 1) / TILE_M;
    const int64_t mn_tiles = m_tiles * p.n_tiles;
    const int64_t n_work = WGRAD ? mn_tiles * p.splits : mn_tiles;
    // chunks of one work item
    auto chunks_of = [&](int64_t w) -> int {
        if (!WGRAD) return K / KC;
        const int64_t sp = w / mn_tiles, s0 = sp * p.sites_per_split;
        const int64_t cnt = (p.M - s0) < p.sites_per_split ? (p.M - s0) : p.sites_per_split;
        return (int)((cnt + KC - 1) / KC);
    };
    const uint32_t acc_cols = (uint32_t)(PB * N);           // one accumulator = PB groups of N columns
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2 * acc_cols) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], PROD_WARPS * 32);       // one asynchronous arrival per producer thread
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t smem_base = smem_u32(smem);

    if (warp > MMA_WARP) {
        // ------------------------------------------------------------------ producers
        const int row = tid - (MMA_WARP + 1) * 32;    // 0..127
        uint32_t it = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int n_chunks = chunks_of(w);
            const int64_t sp = WGRAD ? w / mn_tiles : 0;
            const int64_t rem = WGRAD ? w - sp * mn_tiles : w;
            const int64_t mt = rem / p.n_tiles;
            const int nt = (int)(rem - mt * p.n_tiles);
            // forward: this thread's site, fixed for the tile
            const int64_t m = mt * TILE_M + row;
            const bool live = !WGRAD && m < p.M;
            int b = 0, y = 0, x = 0;
            if (live) xb_conv_site(g, m, b, y, x);
            int sites4[4][3];                                  // MAP = 1, forward: the four rows this thread feeds
            if (MAP == 1 && !WGRAD) {
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    const int64_t mm = mt * TILE_M + xb_v2_row(row, gi);
                    sites4[gi][0] = -1, sites4[gi][1] = 0, sites4[gi][2] = 0;
                    if (mm < p.M) xb_conv_site(g, mm, sites4[gi][0], sites4[gi][1], sites4[gi][2]);
                }
            }
            const int64_t w_off = WGRAD ? 0 : (int64_t)nt * N * K;        // forward: first weight row of this n tile
            const int64_t site_end = WGRAD ? ((sp + 1) * p.sites_per_split < p.M ? (sp + 1) * p.sites_per_split : p.M) : 0;
            for (int kc = 0; kc < n_chunks; ++kc) {
                const int stage = (int)(it % (uint32_t)S);
                mbar_wait(&empty_bar[stage], ((it / (uint32_t)S) & 1u) ^ 1u);
                const uint32_t base = smem_base + (uint32_t)stage * stage_bytes;
                const uint32_t wbase = base + PA * a_plane;
                auto emit_a = [&](uint32_t dst_off, int64_t src) {       // src < 0: the 16 bytes are zero-filled
                    const uint32_t nbytes = src >= 0 ? 16u : 0u;
                    const int64_t o = src >= 0 ? src : 0;
#pragma unroll
                    for (int q = 0; q < PA; ++q) cp_async16(base + q * a_plane + dst_off, p.in[q] + o, nbytes);
                };
                auto emit_w = [&](uint32_t dst_off, int64_t src) {
                    const uint32_t nbytes = src >= 0 ? 16u : 0u;
                    const int64_t o = src >= 0 ? src + w_off : 0;
#pragma unroll
                    for (int q = 0; q < PB; ++q) cp_async16(wbase + q * w_plane + dst_off, p.w[q] + o, nbytes);
                };
                if (MAP == 0) {
                    if (!WGRAD) xb_stage_fwd(g, row, live, b, y, x, kc, emit_a, emit_w);
                    else xb_stage_wgrad(g, row, mt, sp * p.sites_per_split + (int64_t)kc * KC, site_end, p.w_ld, nt * N, emit_a, emit_w);
                } else {
                    if (!WGRAD) xb_stage_fwd_v2(g, row, sites4, kc, emit_a, emit_w);
                    else xb_stage_wgrad_v2(g, row, mt, sp * p.sites_per_split + (int64_t)kc * KC, site_end, p.w_ld, nt * N, emit_a, emit_w);
                }
                // asynchronous publication: the barrier receives this thread's arrival when the copies issued above have
                // landed - the producer never waits for its own loads, so up to `stages` chunks of loads are in flight
                // (the round-1 wait_group -> fence -> arrive sequence exposed the full load latency once per chunk:
                // ~100 cycles per KB staged, measured on B200, whatever the mapping)
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
                ++it;
            }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ MMA issue (one thread)
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int n_chunks = chunks_of(w);
                const uint32_t a = tcount & 1u;
                mbar_wait(&acc_empty[a], ((tcount >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem + a * acc_cols;
                uint32_t acc = 0;
                for (int kc = 0; kc < n_chunks; ++kc) {
                    const int stage = (int)(it % (uint32_t)S);
                    mbar_wait(&full_bar[stage], (it / (uint32_t)S) & 1u);
                    fence_proxy_async();       // the stage was written by cp.async (generic proxy); the MMA reads it through the async proxy
                    tc_fence_after();
                    const uint32_t base = smem_base + (uint32_t)stage * stage_bytes;
                    const uint32_t w_addr = base + PA * a_plane;
#pragma unroll
                    for (int ks = 0; ks < KC / 16; ++ks) {
#pragma unroll
                        for (int pa = 0; pa < PA; ++pa) {
                            // A plane pa x B planes 0 .. PB-1-pa (one operand of (PB-pa)*N rows) -> accumulator groups pa .. PB-1.
                            // pa = 0 covers every group, so its first instruction of a work item (acc = 0) initialises them all.
                            const int nb = PB - pa;
                            const uint32_t idesc = WGRAD ? make_idesc_mn(TILE_M, nb * N) : make_idesc(TILE_M, nb * N);
                            mma_bf16(d_tmem + (uint32_t)(pa * N), make_desc(base + pa * a_plane + ks * 256, KC),
                                     make_desc(w_addr + ks * 256, KC), idesc, pa == 0 ? acc : 1u);
                        }
                        acc = 1;
                    }
                    mma_commit(&empty_bar[stage]);     // the stage is free once these MMAs have read it
                    ++it;
                }
                mma_commit(&acc_full[a]);
                ++tcount;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: thread = row = TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t tcount = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
            const uint32_t a = tcount & 1u;
            const int64_t sp = WGRAD ? w / mn_tiles : 0;
            const int64_t rem = WGRAD ? w - sp * mn_tiles : w;
            const int64_t mt = rem / p.n_tiles;
            const int nt = (int)(rem - mt * p.n_tiles);
            bool live;
            int64_t orow = 0;
            if (!WGRAD) {
                const int64_t m = mt * TILE_M + tid;
                live = m < p.M;
                if (live) {
                    int b, y, x;
                    xb_conv_site(g, m, b, y, x);
                    orow = (((int64_t)b * p.out_H + (y * p.oys + p.oy0)) * p.out_W + (x * p.oxs + p.ox0)) * p.out_ld + p.out_c0 +
                           (int64_t)nt * N;
                }
                // the bias slice of this n tile (only the epilogue warps read / write s_bias; named barrier 1, 128 threads)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = tid; i < N; i += EPI_WARPS * 32) s_bias[i] = p.bias ? p.bias[nt * N + i] : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            } else {
                const int64_t kcol = mt * TILE_M + tid;
                live = kcol < K;
                orow = ((sp * K + kcol) * p.n_tiles + nt) * (int64_t)N;       // partials [splits, K, N_total]
            }
            mbar_wait(&acc_full[a], (tcount >> 1) & 1u);
            tc_fence_after();
            for (int c0 = 0; c0 < N; c0 += 32) {
                float v[32];
                tmem_ld32(lane_addr + a * acc_cols + (uint32_t)((PB - 1) * N + c0), v);      // smallest group first
#pragma unroll
                for (int gq = PB - 2; gq >= 0; --gq) {
                    float u[32];
                    tmem_ld32(lane_addr + a * acc_cols + (uint32_t)(gq * N + c0), u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += u[j];
                }
                if (!WGRAD) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] += s_bias[c0 + j];
                        if (p.relu) v[j] = fmaxf(v[j], 0.f);
                    }
                }
                if (!WGRAD && p.mask && live) {
                    const uint4 *mk = reinterpret_cast<const uint4 *>(p.mask + orow + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 q = mk[j];
                        const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // bf16 > 0  <=>  sign bit clear and not (+)zero
                            const uint32_t lo16 = ws[i] & 0xffffu, hi16 = ws[i] >> 16;
                            if (!(lo16 != 0u && lo16 < 0x8000u)) v[8 * j + 2 * i] = 0.f;
                            if (!(hi16 != 0u && hi16 < 0x8000u)) v[8 * j + 2 * i + 1] = 0.f;
                        }
                    }
                }
                if (live) {
                    if (p.out_f32) {
                        float4 *o = reinterpret_cast<float4 *>(p.out_f32 + orow + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    if (!WGRAD && p.p_out > 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __nv_bfloat16 h[3][8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) split3_bf16(v[8 * j + i], h[0][i], h[1][i], h[2][i]);
#pragma unroll
                            for (int q = 0; q < 3; ++q)
                                if (q < p.p_out)
                                    reinterpret_cast<uint4 *>(p.out[q] + orow + c0)[j] =
                                        make_uint4(pack2(h[q][0], h[q][1]), pack2(h[q][2], h[q][3]), pack2(h[q][4], h[q][5]),
                                                   pack2(h[q][6], h[q][7]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            ++tcount;
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
    }
}

// weight-gradient finish: sum the split partials [splits, K, N] in split order and scatter to torch's [N, C, KH, KW]
// (column k = (kh, kw, c) of the packed layout -> xb_pack_weight_src); accumulate != 0 adds to dw (autograd .grad)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ ws, int splits, int N, int C, int KH,
                                                           int KW, float scale, float *__restrict__ dw, int accumulate) {
    const int64_t K = (int64_t)C * KH * KW, total = (int64_t)N * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / K, k = i - n * K;              // i indexes the PACKED layout [N, (kh, kw, c)]
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += ws[((int64_t)sp * K + k) * N + n];
        s = __fmul_rn(s, scale);
        const int64_t dst = xb_pack_weight_src(i, C, KH, KW);
        dw[dst] = accumulate ? dw[dst] + s : s;
    }
}

// ---------------------------------------------------------------- operand preparation (HBM-bound, elementwise)
// planes: out is [P, n] bf16; plane 0 = bf16(x), plane q = bf16 of the residual left by planes < q
template <int P>
__global__ void __launch_bounds__(256) split_bf16_kernel(const float *__restrict__ x, int64_t n8, int64_t n,
                                                         __nv_bfloat16 *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        float v[8];
        const int64_t e = i * 8;
        const bool full = e + 8 <= n;
        if (full) {
            const float4 a = reinterpret_cast<const float4 *>(x + e)[0], b = reinterpret_cast<const float4 *>(x + e)[1];
            v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = e + j < n ? x[e + j] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < P; ++q) {
            __nv_bfloat16 h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                h[j] = __float2bfloat16_rn(v[j]);
                v[j] -= __bfloat162float(h[j]);
            }
            __nv_bfloat16 *dst = out + (int64_t)q * n + e;
            if (full && ((n & 7) == 0)) {
                *reinterpret_cast<uint4 *>(dst) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
            } else {
                for (int j = 0; j < 8 && e + j < n; ++j) dst[j] = h[j];
            }
        }
    }
}

// torch weight [N, C, KH, KW] (also a Linear over a flattened [C, H, W] feature map) -> [P, N, (kh, kw, c)] bf16 planes
template <int P>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float *__restrict__ w, int N, int C, int KH, int KW,
                                                          float scale, __nv_bfloat16 *__restrict__ out) {
    const int64_t total = (int64_t)N * C * KH * KW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float v = __fmul_rn(w[xb_pack_weight_src(i, C, KH, KW)], scale);   // scale 1/255: the layer reads raw uint8 pixels
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            out[(int64_t)q * total + i] = h;
            v -= __bfloat162float(h);
        }
    }
}

int fill_params(ConvParams &p, int pa, int pb, const void *in, int64_t in_plane, const void *w, int64_t w_plane, int B,
                int IH, int IW, int C, int OY, int OX, int sy, int sx, int T, const int8_t *dy, const int8_t *dx, int N,
                int n_tile) {
    if (pa < 1 || pb > 3 || pa > pb) return XB_EINVAL;
    if (!in || !w || !dy || !dx) return XB_EINVAL;
    if (B <= 0 || IH <= 0 || IW <= 0 || OY <= 0 || OX <= 0 || sy <= 0 || sx <= 0 || T <= 0 || N <= 0 || n_tile <= 0)
        return XB_EINVAL;
    // one tcgen05.mma spans pb * n_tile columns (<= 256, a multiple of 16); two accumulators of pb * n_tile columns in TMEM
    if (T > XB_CONV_MAX_TAPS || n_tile % 32 != 0 || N % n_tile != 0 || pb * n_tile > 256 || C % 8 != 0) return XB_ERANGE;
    if (!xb_aligned(in, 16) || !xb_aligned(w, 16) || in_plane % 8 != 0 || w_plane % 8 != 0) return XB_EALIGN;
    p.g.B = B, p.g.IH = IH, p.g.IW = IW, p.g.C = C, p.g.OY = OY, p.g.OX = OX, p.g.sy = sy, p.g.sx = sx, p.g.T = T, p.g.N = n_tile;
    for (int t = 0; t < XB_CONV_MAX_TAPS; ++t) p.g.dy[t] = t < T ? dy[t] : 0, p.g.dx[t] = t < T ? dx[t] : 0;
    const __nv_bfloat16 *ib = (const __nv_bfloat16 *)in, *wb = (const __nv_bfloat16 *)w;
    for (int q = 0; q < 3; ++q) p.in[q] = q < pa ? ib + q * in_plane : nullptr, p.w[q] = q < pb ? wb + q * w_plane : nullptr;
    p.M = (int64_t)B * OY * OX;
    p.n_tiles = N / n_tile;
    const uint32_t stage_bytes = pa * xb_conv_a_plane_bytes() + pb * xb_conv_w_plane_bytes(n_tile);
    int stages = (int)((200u * 1024u) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return XB_ERANGE;
    p.stages = stages;
    return XB_OK;
}

template <bool WGRAD, int PA, int PB, int MAP>
int launch_map(const ConvParams &p, int64_t work, void *stream) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(conv_tc_kernel<WGRAD, PA, PB, MAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    const int grid = (int)(work < xb_sm_count() ? work : xb_sm_count());
    const size_t smem = (size_t)p.stages * (PA * xb_conv_a_plane_bytes() + PB * xb_conv_w_plane_bytes(p.g.N));
    conv_tc_kernel<WGRAD, PA, PB, MAP><<<grid, THREADS, smem, (cudaStream_t)stream>>>(p);
    return xb_launch_status();
}

// XB_K12_MAP=0 selects the thread-per-row producer mapping; default 1 = row-coalesced (8 rows x 64 contiguous bytes per warp
// instruction: a quarter of the L1 wavefronts of mapping 0, measured on B200 - DESIGN.md section 3)
template <bool WGRAD, int PA, int PB>
int launch_pp(const ConvParams &p, int64_t work, void *stream) {
    static int map = -1;
    if (map < 0) {
        const char *e = getenv("XB_K12_MAP");
        map = (e && e[0] == '0') ? 0 : 1;
    }
    return map == 1 ? launch_map<WGRAD, PA, PB, 1>(p, work, stream) : launch_map<WGRAD, PA, PB, 0>(p, work, stream);
}

template <bool WGRAD>
int launch(int pa, int pb, const ConvParams &p, int64_t work, void *stream) {
    if (pa == 1 && pb == 1) return launch_pp<WGRAD, 1, 1>(p, work, stream);
    if (pa == 1 && pb == 2) return launch_pp<WGRAD, 1, 2>(p, work, stream);
    if (pa == 1 && pb == 3) return launch_pp<WGRAD, 1, 3>(p, work, stream);
    if (pa == 2 && pb == 2) return launch_pp<WGRAD, 2, 2>(p, work, stream);
    if (pa == 2 && pb == 3) return launch_pp<WGRAD, 2, 3>(p, work, stream);
    if (pa == 3 && pb == 3) return launch_pp<WGRAD, 3, 3>(p, work, stream);
    return XB_EINVAL;
}

}  // namespace

extern "C" int xb_split_bf16(const float *x, int64_t n, int planes, void *out, void *stream) {
    if (!x || !out || n <= 0 || planes < 1 || planes > 3) return XB_EINVAL;
    if (!xb_aligned(x, 16) || !xb_aligned(out, 16)) return XB_EALIGN;
    const int64_t n8 = (n + 7) / 8;
    int64_t want = (n8 + 255) / 256;
    const int grid = (int)(want < (int64_t)xb_sm_count() * 8 ? want : (int64_t)xb_sm_count() * 8);
    if (planes == 1) split_bf16_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n8, n, (__nv_bfloat16 *)out);
    else if (planes == 2) split_bf16_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n8, n, (__nv_bfloat16 *)out);
    else split_bf16_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n8, n, (__nv_bfloat16 *)out);
    return xb_launch_status();
}

extern "C" int xb_pack_conv_weight(const float *w, int N, int C, int KH, int KW, int planes, float scale, void *out,
                                   void *stream) {
    if (!w || !out || N <= 0 || C <= 0 || KH <= 0 || KW <= 0 || planes < 1 || planes > 3) return XB_EINVAL;
    const int64_t total = (int64_t)N * C * KH * KW;
    int64_t want = (total + 255) / 256;
    const int grid = (int)(want < (int64_t)xb_sm_count() * 8 ? want : (int64_t)xb_sm_count() * 8);
    cudaStream_t s = (cudaStream_t)stream;
    if (planes == 1) pack_weight_kernel<1><<<grid, 256, 0, s>>>(w, N, C, KH, KW, scale, (__nv_bfloat16 *)out);
    else if (planes == 2) pack_weight_kernel<2><<<grid, 256, 0, s>>>(w, N, C, KH, KW, scale, (__nv_bfloat16 *)out);
    else pack_weight_kernel<3><<<grid, 256, 0, s>>>(w, N, C, KH, KW, scale, (__nv_bfloat16 *)out);
    return xb_launch_status();
}

extern "C" int xb_gemm_gather_tc(int planes_a, int planes_b, const void *in, int64_t in_plane, const void *w, int64_t w_plane,
                                 const float *bias, const void *relu_mask, int B, int IH, int IW, int C, int OY, int OX,
                                 int sy, int sx, int T, const int8_t *dy, const int8_t *dx, int N, int n_tile, int relu,
                                 void *out_planes, int64_t out_plane, int planes_out, float *out_f32, int out_H, int out_W,
                                 int oys, int oxs, int oy0, int ox0, int64_t out_ld, int out_c0, void *stream) {
    ConvParams p;
    const int rc = fill_params(p, planes_a, planes_b, in, in_plane, w, w_plane, B, IH, IW, C, OY, OX, sy, sx, T, dy, dx, N, n_tile);
    if (rc != XB_OK) return rc;
    if (!out_planes && !out_f32) return XB_EINVAL;
    if (out_planes && (planes_out < 1 || planes_out > 3)) return XB_EINVAL;
    if (((int64_t)T * C) % XB_CONV_KC != 0) return XB_ERANGE;
    if (out_H <= 0 || out_W <= 0 || oys <= 0 || oxs <= 0 || oy0 < 0 || ox0 < 0 || (OY - 1) * oys + oy0 >= out_H ||
        (OX - 1) * oxs + ox0 >= out_W || out_c0 < 0 || out_ld < (int64_t)out_c0 + N)
        return XB_EINVAL;
    if (out_ld % 8 != 0 || out_c0 % 8 != 0 || out_plane % 8 != 0) return XB_EALIGN;   // 16-byte row segments
    if ((out_planes && !xb_aligned(out_planes, 16)) || (out_f32 && !xb_aligned(out_f32, 16)) ||
        (relu_mask && !xb_aligned(relu_mask, 16)))
        return XB_EALIGN;
    p.bias = bias;
    p.mask = (const __nv_bfloat16 *)relu_mask;
    __nv_bfloat16 *ob = (__nv_bfloat16 *)out_planes;
    p.p_out = ob ? planes_out : 0;
    for (int q = 0; q < 3; ++q) p.out[q] = (ob && q < planes_out) ? ob + q * out_plane : nullptr;
    p.out_f32 = out_f32;
    p.relu = relu;
    p.out_H = out_H, p.out_W = out_W, p.oys = oys, p.oxs = oxs, p.oy0 = oy0, p.ox0 = ox0;
    p.out_ld = out_ld, p.out_c0 = out_c0;
    p.splits = 1, p.sites_per_split = 0, p.w_ld = 0;
    const int64_t tiles = (p.M + TILE_M - 1) / TILE_M * p.n_tiles;
    return launch<false>(planes_a, planes_b, p, tiles, stream);
}

extern "C" int xb_wgrad_gather_tc(int planes_a, int planes_b, const void *in, int64_t in_plane, const void *g, int64_t g_plane,
                                  int64_t g_ld, int B, int IH, int IW, int C, int OY, int OX, int sy, int sx, int T,
                                  const int8_t *dy, const int8_t *dx, int N, int n_tile, int splits, float *partials,
                                  void *stream) {
    ConvParams p;
    const int rc = fill_params(p, planes_a, planes_b, in, in_plane, g, g_plane, B, IH, IW, C, OY, OX, sy, sx, T, dy, dx, N, n_tile);
    if (rc != XB_OK) return rc;
    if (!partials || splits <= 0 || g_ld < N) return XB_EINVAL;
    if (!xb_aligned(partials, 16) || g_ld % 8 != 0) return XB_EALIGN;
    p.bias = nullptr, p.mask = nullptr, p.out_f32 = partials, p.p_out = 0;
    for (int q = 0; q < 3; ++q) p.out[q] = nullptr;
    p.relu = 0;
    p.out_H = p.out_W = p.oys = p.oxs = 1, p.oy0 = p.ox0 = 0, p.out_ld = N, p.out_c0 = 0;
    p.w_ld = g_ld;
    const int64_t per = xb_wgrad_sites_per_split(p.M, splits);
    if (per == 0) return XB_EINVAL;     // too many splits for this many sites
    p.splits = splits, p.sites_per_split = per;
    const int64_t K = (int64_t)T * C, work = (K + TILE_M - 1) / TILE_M * p.n_tiles * splits;
    return launch<true>(planes_a, planes_b, p, work, stream);
}

extern "C" int xb_wgrad_reduce(const float *partials, int splits, int N, int C, int KH, int KW, float scale, float *dw,
                               int accumulate, void *stream) {
    if (!partials || !dw || splits <= 0 || N <= 0 || C <= 0 || KH <= 0 || KW <= 0) return XB_EINVAL;
    const int64_t total = (int64_t)N * C * KH * KW;
    int64_t want = (total + 255) / 256;
    const int grid = (int)(want < (int64_t)xb_sm_count() * 8 ? want : (int64_t)xb_sm_count() * 8);
    wgrad_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(partials, splits, N, C, KH, KW, scale, dw, accumulate);
    return xb_launch_status();
}

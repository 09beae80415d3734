// K3-P (EXPERIMENTAL, feeds K12): minibatch gather of uint8 observation rows straight into the bf16 plane tensors the
// tensor-core layers consume - planes[q, b, :] with x/255 = sum_q planes[q] (xb_split_bf16's definition applied to the
// 256 possible values once per CTA).  Same structure as the validated K3 gather_obs_kernel (rollout.cu): one thread drives a
// 3-stage ring of 1-D bulk-async (TMA) row copies, all threads convert through a shared-memory LUT and write 16-byte
// segments; replaces sample_batch + `observations / 255.0` (memory_tools.py:64-84, cnn.py:98) + the split pass.
#include "xb_common.cuh"

namespace {
constexpr int OP_STAGES = 3, OP_THREADS = 256;

template <int P>
__global__ void __launch_bounds__(OP_THREADS) gather_obs_planes_kernel(const uint8_t *__restrict__ src,
                                                                       const int64_t *__restrict__ idx, int64_t B,
                                                                       int row_bytes, uint16_t *__restrict__ dst) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[OP_STAGES];
    __shared__ uint16_t lut[P][256];
    const int tid = threadIdx.x;
    {
        // P >= 2: the float32 value the reference network sees, x / 255.  P == 1: the pixel value itself, which a bf16 holds
        // exactly (8 significant bits) - the first layer's packed weights then carry the 1/255 (xb_pack_conv_weight scale)
        float r = P == 1 ? (float)(tid & 255) : __fdiv_rn((float)(tid & 255), 255.0f);
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const __nv_bfloat16 h = __float2bfloat16_rn(r);
            lut[q][tid & 255] = __bfloat16_as_ushort(h);
            r -= __bfloat162float(h);
        }
    }
    if (tid == 0) {
        for (int s = 0; s < OP_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t n_mine = B > (int64_t)blockIdx.x ? (B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int64_t k) {
        const int s = (int)(k % OP_STAGES);
        const int64_t b = blockIdx.x + k * gridDim.x;
        const int64_t srow = idx ? idx[b] : b;
        mbar_expect_tx(&full[s], (uint32_t)row_bytes);
        bulk_g2s(smem + (size_t)s * row_bytes, src + srow * (int64_t)row_bytes, (uint32_t)row_bytes, &full[s]);
    };
    if (tid == 0)
        for (int k = 0; k < OP_STAGES - 1 && k < n_mine; ++k) issue(k);
    const int units = row_bytes / 8;                                // 8 input bytes -> one 16-byte store per plane
    const int64_t plane = B * (int64_t)row_bytes;
    for (int64_t k = 0; k < n_mine; ++k) {
        const int s = (int)(k % OP_STAGES);
        const uint32_t parity = (uint32_t)((k / OP_STAGES) & 1);
        if (tid == 0 && k + OP_STAGES - 1 < n_mine) issue(k + OP_STAGES - 1);   // stage (k-1)%STAGES is free (barrier below)
        mbar_wait(&full[s], parity);
        const uint8_t *row = smem + (size_t)s * row_bytes;
        const int64_t b = blockIdx.x + k * gridDim.x;
#pragma unroll 2
        for (int c = tid; c < units; c += OP_THREADS) {
            const uint2 p = *reinterpret_cast<const uint2 *>(row + c * 8);
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = (p.x >> (8 * j)) & 0xffu;
                v[4 + j] = (p.y >> (8 * j)) & 0xffu;
            }
#pragma unroll
            for (int q = 0; q < P; ++q) {
                const uint4 o = {(uint32_t)lut[q][v[0]] | ((uint32_t)lut[q][v[1]] << 16),
                                 (uint32_t)lut[q][v[2]] | ((uint32_t)lut[q][v[3]] << 16),
                                 (uint32_t)lut[q][v[4]] | ((uint32_t)lut[q][v[5]] << 16),
                                 (uint32_t)lut[q][v[6]] | ((uint32_t)lut[q][v[7]] << 16)};
                stg_stream16(dst + q * plane + b * (int64_t)row_bytes + (int64_t)c * 8, o);
            }
        }
        __syncthreads();
    }
}
}  // namespace

extern "C" int xb_gather_obs_planes(const uint8_t *src, const int64_t *idx, int64_t B, int64_t row_bytes, int planes,
                                    void *dst, void *stream) {
    if (!src || !dst || B < 0 || row_bytes <= 0 || planes < 1 || planes > 3) return XB_EINVAL;
    if (B == 0) return XB_OK;
    if (row_bytes % 16 != 0 || !xb_aligned(src, 16) || !xb_aligned(dst, 16)) return XB_EALIGN;
    if (row_bytes * OP_STAGES > 200 * 1024) return XB_ERANGE;
    const size_t smem = (size_t)OP_STAGES * row_bytes;
    int64_t ctas = (int64_t)xb_sm_count() * 2;
    if (ctas > B) ctas = B;
    cudaStream_t s = (cudaStream_t)stream;
    if (planes == 1) {
        static bool set1 = false;
        if (!set1) {
            cudaFuncSetAttribute(gather_obs_planes_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            set1 = true;
        }
        gather_obs_planes_kernel<1><<<(int)ctas, OP_THREADS, smem, s>>>(src, idx, B, (int)row_bytes, (uint16_t *)dst);
    } else if (planes == 2) {
        static bool set2 = false;
        if (!set2) {
            cudaFuncSetAttribute(gather_obs_planes_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            set2 = true;
        }
        gather_obs_planes_kernel<2><<<(int)ctas, OP_THREADS, smem, s>>>(src, idx, B, (int)row_bytes, (uint16_t *)dst);
    } else {
        static bool set3 = false;
        if (!set3) {
            cudaFuncSetAttribute(gather_obs_planes_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            set3 = true;
        }
        gather_obs_planes_kernel<3><<<(int)ctas, OP_THREADS, smem, s>>>(src, idx, B, (int)row_bytes, (uint16_t *)dst);
    }
    return xb_launch_status();
}

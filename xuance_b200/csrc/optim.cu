// K7 flat-bucket optimiser step: global grad norm (clip_grad_norm_), Adam, Polyak soft update.
// Pure streaming kernels over one contiguous float32 bucket: float4 accesses, grid = k * SM count.
#include "xb_common.cuh"

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float *__restrict__ g, int64_t n, float grad_scale,
                                                         float *__restrict__ norm_out, double *__restrict__ scratch) {
    __shared__ double red[32];
    double acc[1] = {0.0};
    const int64_t n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = g4[i];
        acc[0] += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        float v = g[(n4 << 2) + threadIdx.x];
        acc[0] += (double)v * v;
    }
    grid_sum_finalize<1>(acc, scratch, red, [&](double(&t)[1]) { norm_out[0] = (float)(sqrt(t[0]) * (double)grad_scale); });
}

extern "C" int xb_grad_sumsq(const float *g, int64_t n, float grad_scale, float *norm_out, double *scratch,
                             void *stream) {
    if (!g || !norm_out || !scratch || n <= 0) return XB_EINVAL;
    if (!xb_aligned(g, 16)) return XB_EALIGN;
    int64_t want = ((n >> 2) + 255) / 256;
    int64_t cap = (int64_t)xb_sm_count() * 8;
    if (cap > XB_MAX_PARTIALS) cap = XB_MAX_PARTIALS;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    grad_sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, n, grad_scale, norm_out, scratch);
    return xb_launch_status();
}

// torch.optim.Adam (no amsgrad / weight decay / maximize), same operation order as torch's _single_tensor_adam:
//   exp_avg.lerp_(grad, 1-b1); exp_avg_sq.mul_(b2).addcmul_(grad, grad, 1-b2)
//   denom = exp_avg_sq.sqrt() / bias_correction2_sqrt + eps; param.addcdiv_(exp_avg, denom, value=-step_size)
__device__ __forceinline__ void adam_one(float &p, float &g, float &m, float &v, float coef, float b1, float b2,
                                         float eps, float step_size, float bc2_sqrt) {
    g = g * coef;
    m = m + (g - m) * (1.f - b1);  // lerp
    v = v * b2;
    v = v + (1.f - b2) * g * g;    // addcmul: value * t1 * t2
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
    p = p - step_size * __fdiv_rn(m, denom);
}

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, int64_t n, const float *__restrict__ hyper,
                                                   float b1, float b2, float eps, float max_norm,
                                                   const float *__restrict__ norm, float grad_scale, int write_back) {
    const float step_size = hyper[0], bc2_sqrt = hyper[1];
    float coef = grad_scale;
    if (max_norm > 0.f) {
        float c = max_norm / (norm[0] + 1e-6f);
        coef *= (c < 1.f ? c : 1.f);
    }
    const int64_t n4 = n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(p), *g4 = reinterpret_cast<float4 *>(g);
    float4 *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 P = p4[i], G = g4[i], M = m4[i], V = v4[i];
        adam_one(P.x, G.x, M.x, V.x, coef, b1, b2, eps, step_size, bc2_sqrt);
        adam_one(P.y, G.y, M.y, V.y, coef, b1, b2, eps, step_size, bc2_sqrt);
        adam_one(P.z, G.z, M.z, V.z, coef, b1, b2, eps, step_size, bc2_sqrt);
        adam_one(P.w, G.w, M.w, V.w, coef, b1, b2, eps, step_size, bc2_sqrt);
        p4[i] = P;
        m4[i] = M;
        v4[i] = V;
        if (write_back) g4[i] = G;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        int64_t i = (n4 << 2) + threadIdx.x;
        float P = p[i], G = g[i], M = m[i], V = v[i];
        adam_one(P, G, M, V, coef, b1, b2, eps, step_size, bc2_sqrt);
        p[i] = P;
        m[i] = M;
        v[i] = V;
        if (write_back) g[i] = G;
    }
}

extern "C" int xb_adam_step(float *p, float *g, float *m, float *v, int64_t n, const float *hyper, float beta1,
                            float beta2, float eps, float max_norm, const float *norm, float grad_scale,
                            int write_back_grad, void *stream) {
    if (!p || !g || !m || !v || !hyper || n <= 0) return XB_EINVAL;
    if (max_norm > 0.f && !norm) return XB_EINVAL;
    if (!xb_aligned(p, 16) || !xb_aligned(g, 16) || !xb_aligned(m, 16) || !xb_aligned(v, 16)) return XB_EALIGN;
    int64_t want = ((n >> 2) + 255) / 256;
    int64_t cap = (int64_t)xb_sm_count() * 8;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hyper, beta1, beta2, eps, max_norm, norm,
                                                        grad_scale, write_back_grad);
    return xb_launch_status();
}

__global__ void __launch_bounds__(256) soft_update_kernel(float *__restrict__ tgt, const float *__restrict__ src,
                                                          int64_t n, float tau) {
    const float keep = 1.f - tau;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        // tp.mul_(1 - tau); tp.add_(tau * ep)  -> two roundings, no FMA
        tgt[i] = __fadd_rn(__fmul_rn(tgt[i], keep), __fmul_rn(tau, src[i]));
    }
}

extern "C" int xb_soft_update(float *target, const float *source, int64_t n, float tau, void *stream) {
    if (!target || !source || n <= 0) return XB_EINVAL;
    int64_t want = (n + 255) / 256, cap = (int64_t)xb_sm_count() * 8;
    soft_update_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(target, source, n, tau);
    return xb_launch_status();
}

// SAC elementwise stages fused with their reductions and backward seeds (xuance/torch/learners/policy_gradient/
// sac_learner.py:52-88): actor loss, twin-critic TD backup + loss.  alpha is read from device memory (it is
// exp(log_alpha), a learnable scalar) so no host synchronisation is needed between the three optimiser steps.
#include "xb_common.cuh"

// p_loss = mean(alpha*log_pi - min(q1,q2));  torch.min backward: ties split 1/2 - 1/2.
__global__ void __launch_bounds__(256) sac_actor_kernel(const float *__restrict__ log_pi, const float *__restrict__ q1,
                                                        const float *__restrict__ q2, const float *__restrict__ alpha,
                                                        int64_t B, float inv_bt, float *__restrict__ dlog_pi,
                                                        float *__restrict__ dq1, float *__restrict__ dq2,
                                                        float *__restrict__ stats, double *__restrict__ scratch) {
    __shared__ double red[3 * 32];
    const float a = alpha[0];
    double acc[3] = {0, 0, 0};
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const float lp = log_pi[b], x = q1[b], y = q2[b];
        const float mq = fminf(x, y);
        dlog_pi[b] = a * inv_bt;
        dq1[b] = -inv_bt * (x < y ? 1.f : (x == y ? 0.5f : 0.f));
        dq2[b] = -inv_bt * (y < x ? 1.f : (x == y ? 0.5f : 0.f));
        acc[0] += (double)(a * lp - mq);
        acc[1] += (double)mq;
        acc[2] += (double)lp;
    }
    grid_sum_finalize<3>(acc, scratch, red, [&](double(&t)[3]) {
        stats[0] = (float)(t[0] * (double)inv_bt);  // Ploss
        stats[1] = (float)(t[1] * (double)inv_bt);  // Qvalue = mean(min(q1,q2))
        stats[2] = (float)(t[2] * (double)inv_bt);  // mean(log_pi) (alpha loss needs it)
        stats[3] = 0.f;
    });
}

extern "C" int xb_sac_actor_loss(const float *log_pi, const float *q1, const float *q2, const float *alpha, int64_t B,
                                 int64_t B_total, float *dlog_pi, float *dq1, float *dq2, float *stats, double *scratch,
                                 void *stream) {
    if (!log_pi || !q1 || !q2 || !alpha || !dlog_pi || !dq1 || !dq2 || !stats || !scratch) return XB_EINVAL;
    if (B <= 0 || B_total < B) return XB_EINVAL;
    int64_t want = (B + 255) / 256;
    int grid = (int)(want < XB_MAX_PARTIALS ? want : XB_MAX_PARTIALS);
    sac_actor_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(log_pi, q1, q2, alpha, B, 1.0f / (float)B_total, dlog_pi,
                                                             dq1, dq2, stats, scratch);
    return xb_launch_status();
}

// backup = r + (1-d)*gamma*(target_q - alpha*log_pi_next);  q_loss = mse(q1,backup) + mse(q2,backup)
__global__ void __launch_bounds__(256) sac_critic_kernel(const float *__restrict__ q1, const float *__restrict__ q2,
                                                         const float *__restrict__ target_q,
                                                         const float *__restrict__ log_pi_next,
                                                         const float *__restrict__ rew, const float *__restrict__ term,
                                                         const float *__restrict__ alpha, float gamma, int64_t B,
                                                         float inv_bt, float *__restrict__ dq1, float *__restrict__ dq2,
                                                         float *__restrict__ backup, float *__restrict__ stats,
                                                         double *__restrict__ scratch) {
    __shared__ double red[2 * 32];
    const float a = alpha[0];
    double acc[2] = {0, 0};
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const float tv = __fsub_rn(target_q[b], __fmul_rn(a, log_pi_next[b]));
        const float y = __fadd_rn(rew[b], __fmul_rn(__fmul_rn(__fsub_rn(1.f, term[b]), gamma), tv));
        backup[b] = y;
        const float e1 = q1[b] - y, e2 = q2[b] - y;
        dq1[b] = 2.f * e1 * inv_bt;
        dq2[b] = 2.f * e2 * inv_bt;
        acc[0] += (double)e1 * e1;
        acc[1] += (double)e2 * e2;
    }
    grid_sum_finalize<2>(acc, scratch, red, [&](double(&t)[2]) {
        stats[0] = (float)(t[0] * (double)inv_bt) + (float)(t[1] * (double)inv_bt);  // Qloss
        stats[1] = 0.f;
    });
}

extern "C" int xb_sac_critic_loss(const float *q1, const float *q2, const float *target_q, const float *log_pi_next,
                                  const float *rew, const float *term, const float *alpha, float gamma, int64_t B,
                                  int64_t B_total, float *dq1, float *dq2, float *backup, float *stats, double *scratch,
                                  void *stream) {
    if (!q1 || !q2 || !target_q || !log_pi_next || !rew || !term || !alpha || !dq1 || !dq2 || !backup || !stats ||
        !scratch)
        return XB_EINVAL;
    if (B <= 0 || B_total < B) return XB_EINVAL;
    int64_t want = (B + 255) / 256;
    int grid = (int)(want < XB_MAX_PARTIALS ? want : XB_MAX_PARTIALS);
    sac_critic_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q1, q2, target_q, log_pi_next, rew, term, alpha, gamma, B,
                                                              1.0f / (float)B_total, dq1, dq2, backup, stats, scratch);
    return xb_launch_status();
}

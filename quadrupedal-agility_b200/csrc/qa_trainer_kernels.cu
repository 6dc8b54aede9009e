// Trainer-side streaming kernels (sm_100a), all HBM-bound:
//   K6 qa_gather_minibatch  -- the 9 advanced-index gathers of RolloutStorage.mini_batch_generator
//                              (bbc/rsl_rl/storage/rollout_storage.py:147-155) in one launch
//   K8 qa_clip_adam         -- nn.utils.clip_grad_norm_ + torch.optim.Adam.step on ONE flat fp32 buffer
//                              (bbc/rsl_rl/algorithms/gail.py:409-412 and :361-365), device-side LR
#include "qa_b200.h"
#include "qa_common.cuh"

// ------------------------------------------------------------------------------------------
// K6: one warp per minibatch row; every tensor's row is copied with lane-strided (coalesced) 4 B
// accesses.  Rows are 2684 B (671 floats), i.e. not 16 B aligned, so wider vectors are not available.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_minibatch(const __grid_constant__ QaGatherArgs g) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (long long j = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); j < g.num_rows;
         j += (long long)gridDim.x * warps_per_block) {
        const long long src_row = g.indices[j];
#pragma unroll 1
        for (int t = 0; t < g.num_tensors; ++t) {
            const int w = g.width[t];
            const float* s = g.src[t] + src_row * w;
            float* d = g.dst[t] + j * w;
            int c = lane;
            // 4 independent loads in flight per lane
            for (; c + 96 < w; c += 128) {
                const float v0 = __ldcs(s + c), v1 = __ldcs(s + c + 32), v2 = __ldcs(s + c + 64), v3 = __ldcs(s + c + 96);
                d[c] = v0;
                d[c + 32] = v1;
                d[c + 64] = v2;
                d[c + 96] = v3;
            }
            for (; c < w; c += 32) d[c] = __ldcs(s + c);
        }
    }
}

extern "C" int qa_gather_minibatch(const QaGatherArgs* g, void* stream) {
    QA_CHECK_PTR(g);
    if (g->num_rows == 0) return 0;
    QA_CHECK_PTR(g->indices);
    if (g->num_rows < 0 || g->num_tensors <= 0) return QA_EINVAL;
    if (g->num_tensors > QA_GATHER_MAX_TENSORS) return QA_ERANGE;
    for (int t = 0; t < g->num_tensors; ++t) {
        QA_CHECK_PTR(g->src[t]);
        QA_CHECK_PTR(g->dst[t]);
        if (g->width[t] <= 0) return QA_EINVAL;
    }
    long long blocks = (g->num_rows + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_gather_minibatch<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*g);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K8
// ------------------------------------------------------------------------------------------
struct AdamWorkspace {
    double sumsq;
};

__global__ void __launch_bounds__(256) k_grad_sumsq(QaClipAdamArgs a) {
    __shared__ double s_red[8];
    double acc = 0.0;
    const long long n4 = a.numel >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(a.grads);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        const float x = v.x * a.grad_scale, y = v.y * a.grad_scale, z = v.z * a.grad_scale, w = v.w * a.grad_scale;
        acc += (double)(x * x) + (double)(y * y) + (double)(z * z) + (double)(w * w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.numel & 3)) {
        const float x = a.grads[(n4 << 2) + threadIdx.x] * a.grad_scale;
        acc += (double)(x * x);
    }
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += s_red[k];
        atomicAdd(&reinterpret_cast<AdamWorkspace*>(a.workspace)->sumsq, s);
        if (blockIdx.x == 0) *a.step += 1;          // the update kernel reads the incremented step
    }
}

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float b1, float b2, float eps,
                                          float step_size, float bc2_sqrt) {
    m = m + (g - m) * (1.f - b1);                     // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + (g * g) * (1.f - b2);                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    return p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) k_clip_adam(QaClipAdamArgs a) {
    const double sumsq = reinterpret_cast<const AdamWorkspace*>(a.workspace)->sumsq;
    const float total_norm = (float)sqrt(sumsq);
    float coef = 1.f;
    if (a.max_grad_norm > 0.f) coef = fminf(a.max_grad_norm / (total_norm + 1e-6f), 1.0f);   // clip_grad_norm_
    const float gs = a.grad_scale * coef;
    const int step = *a.step;
    const float lr = *a.lr;
    const float bc1 = 1.f - powf(a.beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(a.beta2, (float)step));
    const float step_size = lr / bc1;
    const long long n4 = a.numel >> 2;
    float4* p4 = reinterpret_cast<float4*>(a.params);
    const float4* g4 = reinterpret_cast<const float4*>(a.grads);
    float4* m4 = reinterpret_cast<float4*>(a.exp_avg);
    float4* v4 = reinterpret_cast<float4*>(a.exp_avg_sq);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 p = p4[i], m = m4[i], v = v4[i];
        const float4 g = g4[i];
        p.x = adam_one(p.x, g.x * gs, m.x, v.x, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        p.y = adam_one(p.y, g.y * gs, m.y, v.y, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        p.z = adam_one(p.z, g.z * gs, m.z, v.z, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        p.w = adam_one(p.w, g.w * gs, m.w, v.w, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        p4[i] = p;
        m4[i] = m;
        v4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.numel & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        float m = a.exp_avg[i], v = a.exp_avg_sq[i];
        a.params[i] = adam_one(a.params[i], a.grads[i] * gs, m, v, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        a.exp_avg[i] = m;
        a.exp_avg_sq[i] = v;
    }
    if (a.grad_norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.grad_norm_out = total_norm;
}

extern "C" int qa_clip_adam(const QaClipAdamArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->numel == 0) return 0;
    QA_CHECK_PTR(a->params);
    QA_CHECK_PTR(a->grads);
    QA_CHECK_PTR(a->exp_avg);
    QA_CHECK_PTR(a->exp_avg_sq);
    QA_CHECK_PTR(a->lr);
    QA_CHECK_PTR(a->step);
    QA_CHECK_PTR(a->workspace);
    if (a->numel < 0) return QA_EINVAL;
    if ((((uintptr_t)a->params | (uintptr_t)a->grads | (uintptr_t)a->exp_avg | (uintptr_t)a->exp_avg_sq) & 15u) != 0)
        return QA_EINVAL;                               // flat buffers are 16 B aligned (float4 path)
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t err = cudaMemsetAsync(a->workspace, 0, sizeof(AdamWorkspace), s);
    if (err != cudaSuccess) return (int)err;
    long long blocks = ((a->numel >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_grad_sumsq<<<(unsigned)blocks, 256, 0, s>>>(*a);
    err = cudaGetLastError();
    if (err != cudaSuccess) return (int)err;
    k_clip_adam<<<(unsigned)blocks, 256, 0, s>>>(*a);
    QA_LAUNCH_RET();
}

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
            const float* s = g.src[t] + src_row * (g.src_pitch[t] > 0 ? g.src_pitch[t] : w) + g.src_col0[t];
            float* d = g.dst[t] + j * (g.dst_pitch[t] > 0 ? g.dst_pitch[t] : w) + g.dst_col0[t];
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
        if (g->src_col0[t] < 0 || g->dst_col0[t] < 0) return QA_EINVAL;
        if (g->dst_pitch[t] != 0 && g->dst_pitch[t] < g->dst_col0[t] + g->width[t]) return QA_EINVAL;
        if (g->src_pitch[t] != 0 && g->src_pitch[t] < g->src_col0[t] + g->width[t]) return QA_EINVAL;
        if ((g->dst_pitch[t] == 0 && g->dst_col0[t] != 0) || (g->src_pitch[t] == 0 && g->src_col0[t] != 0)) return QA_EINVAL;
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
    const float wd = a.weight_decay;
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
        p.x = adam_one(p.x, g.x * gs + wd * p.x, m.x, v.x, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        p.y = adam_one(p.y, g.y * gs + wd * p.y, m.y, v.y, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        p.z = adam_one(p.z, g.z * gs + wd * p.z, m.z, v.z, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        p.w = adam_one(p.w, g.w * gs + wd * p.w, m.w, v.w, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
        p4[i] = p;
        m4[i] = m;
        v4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.numel & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        float m = a.exp_avg[i], v = a.exp_avg_sq[i];
        a.params[i] = adam_one(a.params[i], a.grads[i] * gs + wd * a.params[i], m, v, a.beta1, a.beta2, a.eps, step_size, bc2_sqrt);
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

// K8c: a chain of Adam steps over one flat buffer (see qa_b200.h).  Thread = one float4 of the union range; the ops are applied
// to it in order, in registers.
__global__ void __launch_bounds__(256) k_adam_chain(const __grid_constant__ QaAdamChainArgs a, long long first4, long long last4) {
    __shared__ float s_step_size[QA_ADAM_CHAIN_MAX], s_bc2_sqrt[QA_ADAM_CHAIN_MAX];
    __shared__ unsigned s_last;
    if (threadIdx.x < a.num_ops) {
        const QaAdamChainOp& o = a.ops[threadIdx.x];
        const int step = *o.step + 1;
        const float bc1 = 1.f - powf(a.beta1, (float)step);
        s_bc2_sqrt[threadIdx.x] = sqrtf(1.f - powf(a.beta2, (float)step));
        s_step_size[threadIdx.x] = *o.lr / bc1;
    }
    __syncthreads();
    for (long long i = first4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < last4; i += (long long)gridDim.x * blockDim.x) {
        float4 p = reinterpret_cast<float4*>(a.params)[i];
        const float4 g = reinterpret_cast<const float4*>(a.grads)[i];
        bool touched = false;
        for (int k = 0; k < a.num_ops; ++k) {
            const QaAdamChainOp& o = a.ops[k];
            if (i * 4 < o.lo || i * 4 >= o.hi) continue;
            const long long j = i - (o.lo >> 2);
            float4 m = reinterpret_cast<float4*>(o.exp_avg)[j], v = reinterpret_cast<float4*>(o.exp_avg_sq)[j];
            const float wd = o.weight_decay, ss = s_step_size[k], bq = s_bc2_sqrt[k];
            p.x = adam_one(p.x, g.x * a.grad_scale + wd * p.x, m.x, v.x, a.beta1, a.beta2, a.eps, ss, bq);
            p.y = adam_one(p.y, g.y * a.grad_scale + wd * p.y, m.y, v.y, a.beta1, a.beta2, a.eps, ss, bq);
            p.z = adam_one(p.z, g.z * a.grad_scale + wd * p.z, m.z, v.z, a.beta1, a.beta2, a.eps, ss, bq);
            p.w = adam_one(p.w, g.w * a.grad_scale + wd * p.w, m.w, v.w, a.beta1, a.beta2, a.eps, ss, bq);
            reinterpret_cast<float4*>(o.exp_avg)[j] = m;
            reinterpret_cast<float4*>(o.exp_avg_sq)[j] = v;
            touched = true;
        }
        if (touched) reinterpret_cast<float4*>(a.params)[i] = p;
    }
    // every block has read the step counters before it takes its ticket: the last one increments them
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < a.num_ops) *a.ops[threadIdx.x].step += 1;
        if (threadIdx.x == 0) *a.ticket = 0u;
    }
}

extern "C" int qa_adam_chain(const QaAdamChainArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    QA_CHECK_PTR(a->params);
    QA_CHECK_PTR(a->grads);
    QA_CHECK_PTR(a->ticket);
    if (a->num_ops <= 0 || a->num_ops > QA_ADAM_CHAIN_MAX) return QA_EINVAL;
    if ((((uintptr_t)a->params | (uintptr_t)a->grads) & 15u) != 0) return QA_EINVAL;
    long long lo = -1, hi = -1;
    for (int k = 0; k < a->num_ops; ++k) {
        const QaAdamChainOp& o = a->ops[k];
        const void* need[] = {o.exp_avg, o.exp_avg_sq, o.lr, o.step};
        for (const void* p : need) QA_CHECK_PTR(p);
        if (o.lo < 0 || o.hi <= o.lo || (o.lo & 3) || (o.hi & 3) || (((uintptr_t)o.exp_avg | (uintptr_t)o.exp_avg_sq) & 15u)) return QA_EINVAL;
        lo = lo < 0 || o.lo < lo ? o.lo : lo;
        hi = o.hi > hi ? o.hi : hi;
    }
    const long long n4 = (hi - lo) >> 2;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    k_adam_chain<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a, lo >> 2, hi >> 2);
    QA_LAUNCH_RET();
}

// K8's update pass alone: the norm (and the step increment) were produced by K31 while it reduced the gradients
extern "C" int qa_adam_apply(const QaClipAdamArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->numel == 0) return 0;
    const void* need[] = {a->params, a->grads, a->exp_avg, a->exp_avg_sq, a->lr, a->step, a->workspace};
    for (const void* p : need) QA_CHECK_PTR(p);
    if (a->numel < 0) return QA_EINVAL;
    if ((((uintptr_t)a->params | (uintptr_t)a->grads | (uintptr_t)a->exp_avg | (uintptr_t)a->exp_avg_sq) & 15u) != 0) return QA_EINVAL;
    long long blocks = ((a->numel >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_clip_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K9: activation backward fused with the bias gradient:  gz = gy * act'(y),  db[c] = sum_r gz[r,c]
//     (the element-wise half of the backward of every Linear+ELU/ReLU, actor_critic.py:113-129).  The ELU
//     derivative is recovered from the saved OUTPUT: elu'(z) = 1 for y > 0, y + 1 otherwise.
//     Block = 32 columns x 8 warps; each warp strides over the rows of a 256-row strip with lane = column
//     (coalesced 128 B accesses), partial column sums meet in shared memory and leave with one atomicAdd per
//     column per block.  3 passes over M x N (read gy, y; write gz).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_act_bwd(QaActBwdArgs a, int rows_per_block) {
    __shared__ float s_part[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min((long long)a.M, r0 + rows_per_block);
    float acc = 0.f;
    const float addend_scale = (a.addend != nullptr && a.addend_scale != nullptr) ? __ldg(a.addend_scale) : 1.f;
    if (col < a.N) {
        for (long long r = r0 + w; r < r1; r += 8) {
            float g = a.gy[r * a.gy_pitch + col];
            if (a.addend != nullptr) g = g + addend_scale * a.addend[r * a.addend_pitch + col];
            if (a.act != 0) {
                const float y = a.y[r * a.y_pitch + col];
                if (a.act == 1) g = y > 0.f ? g : g * (y + 1.0f);
                else g = y > 0.f ? g : 0.f;
            }
            if (a.gz != nullptr) a.gz[r * a.gz_pitch + col] = g;
            acc += g;
        }
    }
    if (a.db != nullptr) {
        s_part[w][lane] = acc;
        __syncthreads();
        if (w == 0 && col < a.N) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += s_part[k][lane];
            atomicAdd(a.db + col, s);
        }
    }
}

extern "C" int qa_act_bwd(const QaActBwdArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->M == 0) return 0;
    QA_CHECK_PTR(a->gy);
    if (a->M < 0 || a->N <= 0 || a->act < 0 || a->act > 2) return QA_EINVAL;
    if (a->act != 0) QA_CHECK_PTR(a->y);
    if (a->gz == nullptr && a->db == nullptr) return QA_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (a->db != nullptr && a->zero_db) {
        cudaError_t e = cudaMemsetAsync(a->db, 0, sizeof(float) * a->N, s);
        if (e != cudaSuccess) return (int)e;
    }
    // strip height: 256 rows for wide tensors; narrow ones (a few 32-column blocks) get shorter strips so that the grid still
    // covers the chip several times over (a warp walks its strip row by row: long strips are latency bound)
    const int col_blocks = (a->N + 31) / 32;
    int rows = 256;
    while (rows > 32 && (long long)col_blocks * ((a->M + rows - 1) / rows) < 148 * 6) rows >>= 1;
    dim3 grid(col_blocks, (unsigned)((a->M + rows - 1) / rows));
    k_act_bwd<<<grid, 256, 0, s>>>(*a, rows);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K10: PPO loss, forward + backward in one pass (gail.py:367-408): per sample the Normal log-prob, ratio,
//      clipped surrogate, clipped value loss, bound loss and KL; the gradients w.r.t. the action mean, the value and
//      the (broadcast) std parameter are produced in the same kernel, so autograd resumes at the network outputs.
//      Thread per sample, 12 action dims in registers; block-level reductions, one atomic per block per quantity.
//      torch.max(a, b) sends half of the gradient to each side on ties, which matters inside the clip range where
//      surrogate == surrogate_clipped exactly: both halves carry the same derivative there.
// ------------------------------------------------------------------------------------------
#define PL_A QA_NUM_DOF
__global__ void __launch_bounds__(256) k_ppo_loss(QaPpoLossArgs p) {
    __shared__ float s_red[8][QA_PPO_STATS + PL_A];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float invM = 1.0f / (float)p.M;
    float st[QA_PPO_STATS + PL_A];
#pragma unroll
    for (int k = 0; k < QA_PPO_STATS + PL_A; ++k) st[k] = 0.f;
    if (i < p.M) {
        float mu[PL_A], sg[PL_A], diff[PL_A];
        float logp = 0.f, kl = 0.f, bl = 0.f;
#pragma unroll
        for (int j = 0; j < PL_A; ++j) {
            mu[j] = p.mu[i * p.mu_pitch + j];
            sg[j] = p.std[j];
            const float a = p.actions[i * PL_A + j];
            diff[j] = a - mu[j];
            logp += -(diff[j] * diff[j]) / (2.f * sg[j] * sg[j]) - logf(sg[j]) - 0.9189385332046727f;
            const float os = p.old_sigma[i * PL_A + j], om = p.old_mu[i * PL_A + j];
            kl += logf(sg[j] / os + 1.e-5f) + (os * os + (om - mu[j]) * (om - mu[j])) / (2.0f * sg[j] * sg[j]) - 0.5f;
            const float hi = fmaxf(mu[j] - 1.0f, 0.f), lo = fminf(mu[j] + 1.0f, 0.f);
            bl += lo * lo + hi * hi;
        }
        const float adv = p.advantages[i];
        const float ratio = expf(logp - p.old_logp[i]);
        const float rc = fminf(fmaxf(ratio, 1.0f - p.clip), 1.0f + p.clip);
        const float s1 = -adv * ratio, s2 = -adv * rc;
        const float surr = fmaxf(s1, s2);
        // d surr / d ratio
        const float in_range = (ratio >= 1.0f - p.clip && ratio <= 1.0f + p.clip) ? 1.f : 0.f;
        float w1, w2;
        if (s1 > s2) { w1 = 1.f; w2 = 0.f; } else if (s2 > s1) { w1 = 0.f; w2 = 1.f; } else { w1 = 0.5f; w2 = 0.5f; }
        const float dsurr_dratio = -adv * (w1 + w2 * in_range);
        const float dlogp = p.c_surr * invM * dsurr_dratio * ratio;
        // value loss
        const float v = p.value[i * p.value_pitch], R = p.returns[i];
        float vl, dv;
        if (p.use_clipped_value_loss) {
            const float tv = p.target_values[i];
            const float dvt = v - tv;
            const float vc = tv + fminf(fmaxf(dvt, -p.clip), p.clip);
            const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
            vl = fmaxf(l1, l2);
            const float dvc = (dvt >= -p.clip && dvt <= p.clip) ? 1.f : 0.f;
            float u1, u2;
            if (l1 > l2) { u1 = 1.f; u2 = 0.f; } else if (l2 > l1) { u1 = 0.f; u2 = 1.f; } else { u1 = 0.5f; u2 = 0.5f; }
            dv = u1 * 2.f * (v - R) + u2 * 2.f * (vc - R) * dvc;
        } else {
            vl = (R - v) * (R - v);
            dv = 2.f * (v - R);
        }
        p.dvalue[i] = p.c_value * invM * dv;
#pragma unroll
        for (int j = 0; j < PL_A; ++j) {
            const float s2j = sg[j] * sg[j];
            const float hi = fmaxf(mu[j] - 1.0f, 0.f), lo = fminf(mu[j] + 1.0f, 0.f);
            p.dmu[i * PL_A + j] = dlogp * diff[j] / s2j + p.c_bound * invM * 2.f * (lo + hi);
            // d loss / d sigma_j of this sample: through log-prob, plus the entropy term (-c_ent * mean(sum log sigma))
            st[QA_PPO_STATS + j] = dlogp * ((diff[j] * diff[j]) / (s2j * sg[j]) - 1.f / sg[j]) - p.c_entropy * invM / sg[j];
        }
        st[0] = surr * invM;
        st[1] = vl * invM;
        st[2] = bl * invM;
        st[3] = kl * invM;
    }
    // block reduction: warp shuffles, then 8 partials in smem
#pragma unroll
    for (int k = 0; k < QA_PPO_STATS + PL_A; ++k) st[k] = warp_sum(st[k]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < QA_PPO_STATS + PL_A; ++k) s_red[w][k] = st[k];
    }
    __syncthreads();
    if (threadIdx.x < QA_PPO_STATS + PL_A) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_red[k][threadIdx.x];
        if (threadIdx.x < QA_PPO_STATS) atomicAdd(p.stats + threadIdx.x, s);
        else atomicAdd(p.dstd + (threadIdx.x - QA_PPO_STATS), s);
    }
}

extern "C" int qa_ppo_loss(const QaPpoLossArgs* p, void* stream) {
    QA_CHECK_PTR(p);
    if (p->M <= 0) return QA_EINVAL;
    QA_CHECK_PTR(p->mu);
    QA_CHECK_PTR(p->std);
    QA_CHECK_PTR(p->value);
    QA_CHECK_PTR(p->actions);
    QA_CHECK_PTR(p->old_logp);
    QA_CHECK_PTR(p->advantages);
    QA_CHECK_PTR(p->returns);
    QA_CHECK_PTR(p->target_values);
    QA_CHECK_PTR(p->old_mu);
    QA_CHECK_PTR(p->old_sigma);
    QA_CHECK_PTR(p->dmu);
    QA_CHECK_PTR(p->dvalue);
    QA_CHECK_PTR(p->dstd);
    QA_CHECK_PTR(p->stats);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(p->stats, 0, sizeof(float) * QA_PPO_STATS, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(p->dstd, 0, sizeof(float) * PL_A, s);
    if (e != cudaSuccess) return (int)e;
    k_ppo_loss<<<(unsigned)((p->M + 255) / 256), 256, 0, s>>>(*p);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K12: row losses (gail.py:352-365), forward + gradient in one pass.  Warp per row, lanes = columns (coalesced),
//      one shuffle butterfly per row, one atomic per block.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_loss(QaRowLossArgs p) {
    __shared__ float s_red[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float invM = 1.0f / (float)p.M;
    const float inv_all = invM / (float)p.W;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * 8 + w; i < p.M; i += (long long)gridDim.x * 8) {
        float d = 0.f;
        if (lane < p.W) d = p.a[i * p.a_pitch + lane] - p.b[i * p.b_pitch + lane];
        const float ss = warp_sum(d * d);
        if (p.mode == 0) {
            acc += ss * inv_all;
            if (lane < p.W) p.da[i * p.da_pitch + lane] = 2.f * d * inv_all;
        } else {
            const float n = sqrtf(ss);
            acc += n * invM;
            if (lane < p.W) p.da[i * p.da_pitch + lane] = n > 0.f ? d / n * invM : 0.f;   // torch: subgradient 0 at 0
        }
    }
    if (lane == 0) s_red[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_red[k];
        atomicAdd(p.loss, s);
    }
}

extern "C" int qa_row_loss(const QaRowLossArgs* p, void* stream) {
    QA_CHECK_PTR(p);
    if (p->M <= 0 || p->W <= 0 || p->W > 32 || p->mode < 0 || p->mode > 1) return QA_EINVAL;
    QA_CHECK_PTR(p->a);
    QA_CHECK_PTR(p->b);
    QA_CHECK_PTR(p->da);
    QA_CHECK_PTR(p->loss);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(p->loss, 0, sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
    long long blocks = (p->M + 63) / 64;
    if (blocks > 148 * 4) blocks = 148 * 4;
    k_row_loss<<<(unsigned)blocks, 256, 0, s>>>(*p);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K13: adaptive-KL learning rate (gail.py:368-379) + running sums of the logged statistics, one thread.
// ------------------------------------------------------------------------------------------
__global__ void k_ppo_scalars(QaPpoScalarsArgs p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float kl = *p.kl;
    if (p.desired_kl > 0.f) {
        float lr = *p.lr;
        if (kl > p.desired_kl * 2.0f) lr = fmaxf(lr / 1.5f, p.lr_min);
        else if (kl < p.desired_kl / 2.0f && kl > 0.0f) lr = fminf(lr * 1.5f, p.lr_max);
        *p.lr = lr;
    }
    float ent = 0.f;                                            // Normal.entropy().sum(-1).mean() with a broadcast std
    for (int j = 0; j < p.num_actions; ++j) ent += 1.4189385332046727f + logf(p.std[j]);
    p.stats_accum[0] += p.ppo_stats[0];
    p.stats_accum[1] += p.ppo_stats[1];
    p.stats_accum[2] += p.ppo_stats[2];
    p.stats_accum[3] += ent;
    p.stats_accum[4] += *p.priv_reg_loss;
    p.stats_accum[5] += *p.estimator_loss;
    p.stats_accum[6] += kl;
}

extern "C" int qa_ppo_scalars(const QaPpoScalarsArgs* p, void* stream) {
    QA_CHECK_PTR(p);
    QA_CHECK_PTR(p->ppo_stats);
    QA_CHECK_PTR(p->std);
    QA_CHECK_PTR(p->priv_reg_loss);
    QA_CHECK_PTR(p->estimator_loss);
    QA_CHECK_PTR(p->kl);
    QA_CHECK_PTR(p->lr);
    QA_CHECK_PTR(p->stats_accum);
    if (p->num_actions <= 0) return QA_EINVAL;
    k_ppo_scalars<<<1, 32, 0, (cudaStream_t)stream>>>(*p);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K15: TSC PPO loss, forward + backward in one pass (tsc/rsl_rl/algorithms/ppo.py:176-262).  Thread per sample: the
//      3 mode logits and the 18 continuous dims live in registers.  Categorical follows torch.distributions:
//      probs = softmax(z) (re-normalised), logits = log(clamp(probs, eps, 1 - eps)), entropy = -sum(probs * logits);
//      the clamp passes gradient only inside [eps, 1 - eps].
// ------------------------------------------------------------------------------------------
#define TL_D QA_TSC_NUM_MODES
#define TL_A QA_TSC_NUM_CONT
#define TL_STATS 4
__global__ void __launch_bounds__(256) k_ppo_loss_tsc(QaPpoLossTscArgs p) {
    __shared__ float s_red[8][TL_STATS + TL_A];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float invM = 1.0f / (float)p.M;
    const float EPSF = 1.1920928955078125e-07f;
    float st[TL_STATS + TL_A];
#pragma unroll
    for (int k = 0; k < TL_STATS + TL_A; ++k) st[k] = 0.f;
    if (i < p.M) {
        const float adv = p.advantages[i];
        // ---- mode head ----
        float z[TL_D], pr[TL_D], L[TL_D], msk[TL_D];
        float zmax = -INFINITY;
#pragma unroll
        for (int k = 0; k < TL_D; ++k) {
            z[k] = p.logits[i * p.logits_pitch + k];
            zmax = fmaxf(zmax, z[k]);
        }
        float S = 0.f;
#pragma unroll
        for (int k = 0; k < TL_D; ++k) {
            pr[k] = expf(z[k] - zmax);
            S += pr[k];
        }
        float S2 = 0.f;
#pragma unroll
        for (int k = 0; k < TL_D; ++k) {
            pr[k] = pr[k] / S;
            S2 += pr[k];
        }
        float Hd = 0.f;
#pragma unroll
        for (int k = 0; k < TL_D; ++k) {
            pr[k] = pr[k] / S2;                                     // Categorical(probs=...) normalises again
            msk[k] = (pr[k] >= EPSF && pr[k] <= 1.f - EPSF) ? 1.f : 0.f;
            L[k] = logf(fminf(fmaxf(pr[k], EPSF), 1.f - EPSF));
            Hd -= pr[k] * L[k];
        }
        int a_d = (int)p.actions[i * p.actions_pitch];
        a_d = min(max(a_d, 0), TL_D - 1);
        float logp_d = 0.f, pa = 1.f, ma = 0.f;
#pragma unroll
        for (int k = 0; k < TL_D; ++k)
            if (k == a_d) logp_d = L[k], pa = pr[k], ma = msk[k];
        const float lo_c = 1.0f - p.clip, hi_c = 1.0f + p.clip;
        float surr_total = 0.f;
        float dlogp_d;
        {
            const float ratio = expf(logp_d - p.old_logp_d[i]);
            const float rc = fminf(fmaxf(ratio, lo_c), hi_c);
            const float s1 = -adv * ratio, s2 = -adv * rc;
            surr_total += fmaxf(s1, s2);
            const float in_range = (ratio >= lo_c && ratio <= hi_c) ? 1.f : 0.f;
            float w1, w2;
            if (s1 > s2) { w1 = 1.f; w2 = 0.f; } else if (s2 > s1) { w1 = 0.f; w2 = 1.f; } else { w1 = 0.5f; w2 = 0.5f; }
            dlogp_d = invM * (-adv * (w1 + w2 * in_range)) * ratio;
        }
        {
            // g_k = d loss / d probs_k ; dz_j = probs_j (g_j - sum_k g_k probs_k)
            float g[TL_D], G = 0.f;
#pragma unroll
            for (int k = 0; k < TL_D; ++k) {
                g[k] = p.c_entropy * invM * (L[k] + msk[k]);
                if (k == a_d) g[k] += dlogp_d * ma / pa;
                G += g[k] * pr[k];
            }
#pragma unroll
            for (int k = 0; k < TL_D; ++k) p.dlogits[i * p.dlogits_pitch + k] = pr[k] * (g[k] - G);
        }
        // ---- continuous head ----
        float mu[TL_A], sg[TL_A], diff[TL_A];
        float logp = 0.f, kl = 0.f, ent_c = 0.f;
#pragma unroll
        for (int j = 0; j < TL_A; ++j) {
            mu[j] = p.mu[i * p.mu_pitch + j];
            sg[j] = p.std[j];
            const float a = p.actions[i * p.actions_pitch + 1 + j];
            diff[j] = a - mu[j];
            const float lsg = logf(sg[j]);
            logp += -(diff[j] * diff[j]) / (2.f * sg[j] * sg[j]) - lsg - 0.9189385332046727f;
            ent_c += 1.4189385332046727f + lsg;
            const float os = p.old_sigma[i * TL_A + j], om = p.old_mu[i * TL_A + j];
            kl += logf(sg[j] / os + 1.e-5f) + (os * os + (om - mu[j]) * (om - mu[j])) / (2.0f * sg[j] * sg[j]) - 0.5f;
        }
        ent_c = ent_c / (float)TL_A;
        float dlogp_c;
        {
            const float ratio = expf(logp - p.old_logp_c[i]);
            const float rc = fminf(fmaxf(ratio, lo_c), hi_c);
            const float s1 = -adv * ratio, s2 = -adv * rc;
            surr_total += fmaxf(s1, s2);
            const float in_range = (ratio >= lo_c && ratio <= hi_c) ? 1.f : 0.f;
            float w1, w2;
            if (s1 > s2) { w1 = 1.f; w2 = 0.f; } else if (s2 > s1) { w1 = 0.f; w2 = 1.f; } else { w1 = 0.5f; w2 = 0.5f; }
            dlogp_c = invM * (-adv * (w1 + w2 * in_range)) * ratio;
        }
        // ---- value loss ----
        const float v = p.value[i * p.value_pitch], R = p.returns[i];
        float vl, dv;
        if (p.use_clipped_value_loss) {
            const float tv = p.target_values[i];
            const float dvt = v - tv;
            const float vc = tv + fminf(fmaxf(dvt, -p.clip), p.clip);
            const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
            vl = fmaxf(l1, l2);
            const float dvc = (dvt >= -p.clip && dvt <= p.clip) ? 1.f : 0.f;
            float u1, u2;
            if (l1 > l2) { u1 = 1.f; u2 = 0.f; } else if (l2 > l1) { u1 = 0.f; u2 = 1.f; } else { u1 = 0.5f; u2 = 0.5f; }
            dv = u1 * 2.f * (v - R) + u2 * 2.f * (vc - R) * dvc;
        } else {
            vl = (R - v) * (R - v);
            dv = 2.f * (v - R);
        }
        p.dvalue[i] = p.c_value * invM * dv;
#pragma unroll
        for (int j = 0; j < TL_A; ++j) {
            const float s2j = sg[j] * sg[j];
            p.dmu[i * p.dmu_pitch + j] = dlogp_c * diff[j] / s2j;
            st[TL_STATS + j] = dlogp_c * ((diff[j] * diff[j]) / (s2j * sg[j]) - 1.f / sg[j]) -
                               p.c_entropy * invM / ((float)TL_A * sg[j]);
        }
        st[0] = surr_total * invM;
        st[1] = vl * invM;
        st[2] = (ent_c + Hd) * invM;
        st[3] = kl * invM;                                          // same slot as K10, so K13 serves both trainers
    }
#pragma unroll
    for (int k = 0; k < TL_STATS + TL_A; ++k) st[k] = warp_sum(st[k]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < TL_STATS + TL_A; ++k) s_red[w][k] = st[k];
    }
    __syncthreads();
    if (threadIdx.x < TL_STATS + TL_A) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_red[k][threadIdx.x];
        if (threadIdx.x < TL_STATS) atomicAdd(p.stats + threadIdx.x, s);
        else atomicAdd(p.dstd + (threadIdx.x - TL_STATS), s);
    }
}

extern "C" int qa_ppo_loss_tsc(const QaPpoLossTscArgs* p, void* stream) {
    QA_CHECK_PTR(p);
    if (p->M <= 0) return QA_EINVAL;
    const void* need[] = {p->logits, p->mu, p->std, p->value, p->actions, p->old_logp_d, p->old_logp_c, p->advantages,
                          p->returns, p->target_values, p->old_mu, p->old_sigma, p->dlogits, p->dmu, p->dvalue, p->dstd,
                          p->stats};
    for (const void* q : need) QA_CHECK_PTR(q);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(p->stats, 0, sizeof(float) * TL_STATS, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(p->dstd, 0, sizeof(float) * TL_A, s);
    if (e != cudaSuccess) return (int)e;
    k_ppo_loss_tsc<<<(unsigned)((p->M + 255) / 256), 256, 0, s>>>(*p);
    QA_LAUNCH_RET();
}

// Discriminator update: the element-wise / row-wise half of one `SSInfoGAIL.update_ss_info_gail` minibatch step
// (bbc/rsl_rl/algorithms/gail.py:415-541) as five streaming kernels around the tcgen05 trunk GEMMs (K7):
//   K24 qa_disc_prepare      :419-452  replay / expert row gathers, task-obs weighting, per-step multipliers, normalisation
//                                      -> X = [policy | labelled expert | unlabelled expert] (3B, 98) + compact per-row targets
//   K25 qa_disc_heads_loss   :454-490  the three heads (d, eps, classifier) forward, every loss term that hangs off them
//                                      (CE on the soft-maxed classifier, information maximisation, LSGAN, L1), their gradients
//                                      w.r.t. the trunk output and the head parameters, the prior estimate (:462-464), the four
//                                      accuracies (:532-538) -- one pass over the trunk output
//   K27 qa_disc_gp_loss      :492-502  gradient penalty value + the gradient w.r.t. the input gradient, in place
//   K28 qa_disc_reg          :488-490, :504-507  logit regulariser + weight decay: values and gradients
//   K29 / K30 qa_norm_moments / qa_norm_merge  :527-529 + utils.py:63-83  batch moments of the three normalised batches and
//                                      their sequential Chan merge into the running normaliser (fp64), the prior's soft update,
//                                      the policy-std floor (:523-524) -- no device->host round trip (the reference syncs 3x).
// ReLU networks: the double backward of the gradient penalty only sees the activation MASKS (relu'' = 0), so it is a short
// chain of GEMMs on masked operands (schedule in qa_b200/rsl_rl/disc_plan.py).
#include "qa_b200.h"
#include "qa_common.cuh"

#define DU_C QA_DIM_C            // 5 classes
#define DU_H 256                 // trunk output width
#define DU_NH (2 + DU_C)         // heads: d, eps, classifier logits

// ------------------------------------------------------------------------------------------------------------------
// K24: warp per output row
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_disc_prepare(const __grid_constant__ QaDiscPrepareArgs a) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= 3LL * a.B) return;
    const int which = (int)(r / a.B);
    const long long i = r - (long long)which * a.B;
    const float* src;
    long long row;
    if (which == 0) row = a.idx_pi[i], src = a.replay_states;
    else if (which == 1) row = a.idx_lb[i], src = a.expert_lb;
    else row = a.idx_ulb[i], src = a.expert_ulb;
    src += row * a.width;
    const float w = (a.task_obs_weight_decay && a.task_obs_weight != nullptr) ? __ldg(a.task_obs_weight) : 1.f;
    float* x = a.x + r * a.x_pitch;
    for (int c = lane; c < a.width; c += 32) {
        const int slot = c / a.obs_dim, k = c - slot * a.obs_dim;
        float v = __ldg(src + c);
        if (a.task_obs_weight_decay && ((k >= 3 && k < 9) || k >= 33)) v = v * w;                 // :425-432
        v = v * ((float)slot * a.obs_disc_weight_step + 1.f);                                      // :438-444
        v = (v - a.norm_mean[c]) / a.norm_std[c];                                                  // utils.py:97-103
        x[c] = fminf(fmaxf(v, -a.norm_clip), a.norm_clip);
    }
    if (lane == 0) {
        if (which == 0) {
            a.tgt_eps[i] = a.replay_eps[row];
            const float* lc = a.replay_c + row * DU_C;                                             // :458 argmax of the one-hot
            int best = 0;
            float bv = lc[0];
#pragma unroll
            for (int k = 1; k < DU_C; ++k)
                if (lc[k] > bv) bv = lc[k], best = k;
            a.tgt_c[i] = best;
        } else if (which == 1) {
            a.tgt_label[i] = (int)a.expert_label[row];
        }
    }
}

extern "C" int qa_disc_prepare(const QaDiscPrepareArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->B <= 0 || a->width <= 0 || a->obs_dim <= 0 || a->x_pitch < a->width) return QA_EINVAL;
    const void* need[] = {a->replay_states, a->replay_eps, a->replay_c, a->expert_lb, a->expert_label, a->expert_ulb, a->idx_pi,
                          a->idx_lb, a->idx_ulb, a->norm_mean, a->norm_std, a->x, a->tgt_eps, a->tgt_c, a->tgt_label};
    for (const void* p : need) QA_CHECK_PTR(p);
    k_disc_prepare<<<(unsigned)((3LL * a->B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------------------------------
// K25: warp per row, lane = 8 trunk columns.  Head parameters in shared memory as [7][256] + [7].
// stats layout (QaDiscHeadsArgs::stats, accumulated): 0 ss_loss, 1 info_max_loss, 2 disc_loss, 3 us_loss, 7 acc_lb,
// 8 acc_pi, 9 acc_exp, 10 acc_ulb  (4 grad_pen, 5 logit, 6 weight decay come from K27 / K28).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_disc_heads_loss(const __grid_constant__ QaDiscHeadsArgs a) {
    __shared__ __align__(16) float s_w[DU_NH * DU_H];
    __shared__ __align__(16) float s_dw[DU_NH * DU_H];
    __shared__ float s_b[DU_NH], s_db[DU_NH], s_db2[DU_H], s_stat[8], s_prior[DU_C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < DU_NH * DU_H; i += 256) {
        const int n = i / DU_H, k = i - n * DU_H;
        const float* w = n == 0 ? a.w_d : (n == 1 ? a.w_eps : a.w_c + (size_t)(n - 2) * a.w_c_pitch);
        s_w[i] = __ldg(w + k);
        s_dw[i] = 0.f;
    }
    if (tid < DU_NH) {
        s_b[tid] = tid == 0 ? __ldg(a.b_d) : (tid == 1 ? __ldg(a.b_eps) : __ldg(a.b_c + tid - 2));
        s_db[tid] = 0.f;
    }
    if (tid < DU_H) s_db2[tid] = 0.f;
    if (tid < 8) s_stat[tid] = 0.f;
    if (tid < DU_C) s_prior[tid] = 0.f;
    __syncthreads();
    const float invB = 1.0f / (float)a.B;
    const float c_info = a.info_max_coef != nullptr ? __ldg(a.info_max_coef) : 0.f;
    float dw[DU_NH][8];
    float dbh[DU_NH], db2[8], st[8], pr[DU_C];
#pragma unroll
    for (int n = 0; n < DU_NH; ++n) {
        dbh[n] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) dw[n][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) db2[j] = 0.f, st[j] = 0.f;
#pragma unroll
    for (int k = 0; k < DU_C; ++k) pr[k] = 0.f;
    const long long rows = 3LL * a.B;
    for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
        const float* hrow = a.h2 + r * a.h2_pitch + lane * 8;
        const float4 ha = *reinterpret_cast<const float4*>(hrow), hb = *reinterpret_cast<const float4*>(hrow + 4);
        const float h[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        float z[DU_NH];
#pragma unroll
        for (int n = 0; n < DU_NH; ++n) {
            const float4 wa = *reinterpret_cast<const float4*>(s_w + n * DU_H + lane * 8);
            const float4 wb = *reinterpret_cast<const float4*>(s_w + n * DU_H + lane * 8 + 4);
            z[n] = h[0] * wa.x + h[1] * wa.y + h[2] * wa.z + h[3] * wa.w + h[4] * wb.x + h[5] * wb.y + h[6] * wb.z + h[7] * wb.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int n = 0; n < DU_NH; ++n) z[n] += __shfl_xor_sync(QA_FULL, z[n], o);
        }
        const float d = z[0] + s_b[0], eps = z[1] + s_b[1];
        float p[DU_C], c[DU_C];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < DU_C; ++k) p[k] = z[2 + k] + s_b[2 + k], mx = fmaxf(mx, p[k]);
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < DU_C; ++k) p[k] = expf(p[k] - mx), se += p[k];
        int amax = 0;
#pragma unroll
        for (int k = 0; k < DU_C; ++k) {
            p[k] = p[k] / se;                                   // softmax (discriminator.py:68)
            c[k] = fmaxf(p[k], 1e-20f);                         // clamp(min = 1e-20) (:69)
            if (c[k] > c[amax]) amax = k;
        }
        const int which = (int)(r / a.B);
        const long long i = r - (long long)which * a.B;
        float g[DU_NH];                                        // d loss / d head outputs (pre-bias == post-bias)
#pragma unroll
        for (int n = 0; n < DU_NH; ++n) g[n] = 0.f;
        float dc[DU_C];
#pragma unroll
        for (int k = 0; k < DU_C; ++k) dc[k] = 0.f;
        bool has_dc = false;
        if (which == 0) {                                       // policy batch: LSGAN target -1 (:477), L1 on eps (:485)
            const float t = a.tgt_eps[i];
            st[2] += 0.5f * (d + 1.f) * (d + 1.f) * invB;
            st[3] += fabsf(eps - t) * invB;
            g[0] = a.disc_coef * (d + 1.f) * invB;
            g[1] = a.us_coef * ((eps > t) ? 1.f : ((eps < t) ? -1.f : 0.f)) * invB;
            st[5] += (d < 0.f ? 1.f : 0.f) * invB;              // acc_pi
            st[7] += (amax == a.tgt_c[i] ? 1.f : 0.f) * invB;   // acc_ulb (:536-538)
        } else if (which == 1) {                                // labelled expert batch: CE on the soft-maxed output (:455-456)
            const int y = a.tgt_label[i];
            float m2 = c[0];
#pragma unroll
            for (int k = 1; k < DU_C; ++k) m2 = fmaxf(m2, c[k]);
            float s2 = 0.f, e2[DU_C];
#pragma unroll
            for (int k = 0; k < DU_C; ++k) e2[k] = expf(c[k] - m2), s2 += e2[k];
            float cy = 0.f;
#pragma unroll
            for (int k = 0; k < DU_C; ++k) {
                if (k == y) cy = c[k];
                dc[k] = a.ss_coef * (e2[k] / s2 - (k == y ? 1.f : 0.f)) * invB;
            }
            st[0] += (logf(s2) + m2 - cy) * invB;
            st[4] += (amax == y ? 1.f : 0.f) * invB;            // acc_lb
            has_dc = true;
        } else {                                                // unlabelled expert batch: LSGAN target +1, info-max (:466), prior
            st[2] += 0.5f * (d - 1.f) * (d - 1.f) * invB;
            g[0] = a.disc_coef * (d - 1.f) * invB;
            st[6] += (d > 0.f ? 1.f : 0.f) * invB;              // acc_exp
            float ent = 0.f;
#pragma unroll
            for (int k = 0; k < DU_C; ++k) {
                const float lg = logf(c[k] + 1e-20f);
                ent -= c[k] * lg;
                dc[k] = -c_info * (lg + c[k] / (c[k] + 1e-20f)) * invB;
                pr[k] += c[k] * invB;
            }
            st[1] += ent * invB;
            has_dc = true;
        }
        if (has_dc) {                                           // through clamp (pass where p > 1e-20) and softmax
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < DU_C; ++k) {
                if (!(p[k] > 1e-20f)) dc[k] = 0.f;
                dot += dc[k] * p[k];
            }
#pragma unroll
            for (int k = 0; k < DU_C; ++k) g[2 + k] = p[k] * (dc[k] - dot);
        }
        // gradient w.r.t. the trunk output (through relu' of layer 2), head parameter gradients, gradient-penalty operand
        float gz[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) gz[j] = 0.f;
#pragma unroll
        for (int n = 0; n < DU_NH; ++n) {
            const float4 wa = *reinterpret_cast<const float4*>(s_w + n * DU_H + lane * 8);
            const float4 wb = *reinterpret_cast<const float4*>(s_w + n * DU_H + lane * 8 + 4);
            const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                gz[j] += g[n] * w8[j];
                dw[n][j] += g[n] * h[j];
            }
            dbh[n] += g[n];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            gz[j] = h[j] > 0.f ? gz[j] : 0.f;
            db2[j] += gz[j];
        }
        float* out = a.gz2 + r * a.gz2_pitch + lane * 8;
        *reinterpret_cast<float4*>(out) = make_float4(gz[0], gz[1], gz[2], gz[3]);
        *reinterpret_cast<float4*>(out + 4) = make_float4(gz[4], gz[5], gz[6], gz[7]);
        if (which == 2 && a.v2 != nullptr) {                    // v2 = relu'(z2) * w_d : operand of the gradient penalty chain
            const float4 wa = *reinterpret_cast<const float4*>(s_w + lane * 8), wb = *reinterpret_cast<const float4*>(s_w + lane * 8 + 4);
            float* v = a.v2 + i * a.v2_pitch + lane * 8;
            *reinterpret_cast<float4*>(v) = make_float4(h[0] > 0.f ? wa.x : 0.f, h[1] > 0.f ? wa.y : 0.f, h[2] > 0.f ? wa.z : 0.f,
                                                        h[3] > 0.f ? wa.w : 0.f);
            *reinterpret_cast<float4*>(v + 4) = make_float4(h[4] > 0.f ? wb.x : 0.f, h[5] > 0.f ? wb.y : 0.f, h[6] > 0.f ? wb.z : 0.f,
                                                            h[7] > 0.f ? wb.w : 0.f);
        }
    }
    // fold the block's partial sums in shared memory, then one global atomic per element
#pragma unroll
    for (int n = 0; n < DU_NH; ++n) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(s_dw + n * DU_H + lane * 8 + j, dw[n][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(s_db2 + lane * 8 + j, db2[j]);
    if (lane == 0) {                                            // per-row scalars are identical on all lanes: lane 0 contributes
#pragma unroll
        for (int n = 0; n < DU_NH; ++n) atomicAdd(s_db + n, dbh[n]);
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(s_stat + k, st[k]);
#pragma unroll
        for (int k = 0; k < DU_C; ++k) atomicAdd(s_prior + k, pr[k]);
    }
    __syncthreads();
    for (int i = tid; i < DU_NH * DU_H; i += 256) {
        const int n = i / DU_H, k = i - n * DU_H;
        float* dst = n == 0 ? a.dw_d : (n == 1 ? a.dw_eps : a.dw_c + (size_t)(n - 2) * a.dw_c_pitch);
        atomicAdd(dst + k, s_dw[i]);
    }
    if (tid < DU_NH) atomicAdd(tid == 0 ? a.db_d : (tid == 1 ? a.db_eps : a.db_c + tid - 2), s_db[tid]);
    if (tid < DU_H) atomicAdd(a.db2 + tid, s_db2[tid]);
    if (tid < 8) {
        const int slot[8] = {0, 1, 2, 3, 7, 8, 9, 10};         // ss, info_max, disc, us, acc_lb, acc_pi, acc_exp, acc_ulb
        atomicAdd(a.stats + slot[tid], s_stat[tid]);
    }
    if (tid < DU_C) atomicAdd(a.prior_batch + tid, s_prior[tid]);
}

extern "C" int qa_disc_heads_loss(const QaDiscHeadsArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->B <= 0 || a->h2_pitch < DU_H || a->gz2_pitch < DU_H || (a->h2_pitch & 3) || (a->gz2_pitch & 3)) return QA_EINVAL;
    const void* need[] = {a->h2, a->w_d, a->b_d, a->w_eps, a->b_eps, a->w_c, a->b_c, a->tgt_eps, a->tgt_c, a->tgt_label, a->gz2,
                          a->dw_d, a->db_d, a->dw_eps, a->db_eps, a->dw_c, a->db_c, a->db2, a->stats, a->prior_batch};
    for (const void* p : need) QA_CHECK_PTR(p);
    if ((reinterpret_cast<uintptr_t>(a->h2) & 15u) || (reinterpret_cast<uintptr_t>(a->gz2) & 15u)) return QA_EINVAL;
    if (a->v2 != nullptr && ((reinterpret_cast<uintptr_t>(a->v2) & 15u) || (a->v2_pitch & 3) || a->v2_pitch < DU_H)) return QA_EINVAL;
    long long blocks = (3LL * a->B + 31) / 32;                 // >= 4 rows per warp: the fold is amortised
    if (blocks > 148 * 2) blocks = 148 * 2;
    k_disc_heads_loss<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------------------------------
// K27: gradient penalty  mean_i ||g_i||^2  (:492-502): value into stats[4]; g is overwritten by d loss / d g.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_disc_gp_loss(const __grid_constant__ QaDiscGpArgs a) {
    __shared__ float s_red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float invB = 1.0f / (float)a.B;
    float acc = 0.f;
    for (long long r = (long long)blockIdx.x * 8 + warp; r < a.B; r += (long long)gridDim.x * 8) {
        float* g = a.g + r * a.g_pitch;
        float ss = 0.f;
        for (int c = lane; c < a.width; c += 32) {
            const float v = g[c];
            ss += v * v;
            g[c] = 2.f * a.coef * invB * v;
        }
        acc += warp_sum(ss) * invB;
    }
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_red[k];
        atomicAdd(a.stats + 4, s);
    }
}

extern "C" int qa_disc_gp_loss(const QaDiscGpArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->B <= 0 || a->width <= 0 || a->g_pitch < a->width) return QA_EINVAL;
    QA_CHECK_PTR(a->g);
    QA_CHECK_PTR(a->stats);
    long long blocks = (a->B + 7) / 8;
    if (blocks > 148 * 2) blocks = 148 * 2;
    k_disc_gp_loss<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------------------------------
// K28: disc_logit_loss = sum(w_d^2) (:488-490) and disc_weight_decay = sum(W1^2) + sum(W2^2) + sum(w_d^2) (:504-507):
//      values into stats[5], stats[6]; gradients 2 c w added to the flat gradient buffer.  Padding columns hold zeros.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_disc_reg(const __grid_constant__ QaDiscRegArgs a) {
    __shared__ float s_red[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float s_logit = 0.f, s_wd = 0.f;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
    for (int seg = 0; seg < 3; ++seg) {
        const float* w = a.params + a.seg_off[seg];
        float* g = a.grads + a.seg_off[seg];
        const float cg = 2.f * (a.weight_decay_coef + (seg == 2 ? a.logit_reg_coef : 0.f));
        for (long long i = gid; i < a.seg_len[seg]; i += gsz) {
            const float v = w[i];
            s_wd += v * v;
            if (seg == 2) s_logit += v * v;
            g[i] += cg * v;
        }
    }
    s_logit = warp_sum(s_logit);
    s_wd = warp_sum(s_wd);
    if (lane == 0) s_red[0][warp] = s_logit, s_red[1][warp] = s_wd;
    __syncthreads();
    if (threadIdx.x < 2) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_red[threadIdx.x][k];
        atomicAdd(a.stats + 5 + threadIdx.x, s);
    }
}

extern "C" int qa_disc_reg(const QaDiscRegArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    QA_CHECK_PTR(a->params);
    QA_CHECK_PTR(a->grads);
    QA_CHECK_PTR(a->stats);
    for (int s = 0; s < 3; ++s)
        if (a->seg_off[s] < 0 || a->seg_len[s] < 0) return QA_EINVAL;
    k_disc_reg<<<148, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------------------------------
// K29: per batch b (3) and column c: mean and E[x^2] over the B rows, fp64 accumulation.  moments[b][0][c] = mean,
//      moments[b][1][c] = E[x^2] (pooling over ranks = averaging both; var = E[x^2] - mean^2 at merge time).
//      Block = (batch, 32-column group); 8 warps stride the rows, lane = column.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_norm_moments(const __grid_constant__ QaNormMomentsArgs a) {
    __shared__ double s_sum[8][32], s_sq[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, col = blockIdx.x * 32 + lane;
    double s = 0.0, q = 0.0;
    if (col < a.width) {
        const float* x = a.x + ((long long)b * a.B) * a.x_pitch + col;
        for (long long r = warp; r < a.B; r += 8) {
            const double v = (double)x[r * a.x_pitch];
            s += v;
            q += v * v;
        }
    }
    s_sum[warp][lane] = s;
    s_sq[warp][lane] = q;
    __syncthreads();
    if (warp == 0 && col < a.width) {
        double ts = 0.0, tq = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) ts += s_sum[k][lane], tq += s_sq[k][lane];
        a.moments[((long long)b * 2 + 0) * a.width + col] = ts / (double)a.B;
        a.moments[((long long)b * 2 + 1) * a.width + col] = tq / (double)a.B;
    }
}

extern "C" int qa_norm_moments(const QaNormMomentsArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->B <= 0 || a->width <= 0 || a->num_batches <= 0 || a->x_pitch < a->width) return QA_EINVAL;
    QA_CHECK_PTR(a->x);
    QA_CHECK_PTR(a->moments);
    dim3 grid((a->width + 31) / 32, a->num_batches);
    k_norm_moments<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// K30: sequential Chan merge of the batches into the running (mean, var, count) (utils.py:63-83), refresh of the fp32
//      (mean, std) the normalising kernels read, the prior's soft update (:462-464) and the policy-std floor (:523-524).
__global__ void __launch_bounds__(128) k_norm_merge(const __grid_constant__ QaNormMergeArgs a) {
    const int c = threadIdx.x;
    const double inv_w = 1.0 / (double)a.world_size;
    if (c < a.width) {
        double mean = a.mean[c], var = a.var[c], count = *a.count;
        const double bc = (double)a.B * (double)a.world_size;
        for (int b = 0; b < a.num_batches; ++b) {
            const double bm = a.moments[((long long)b * 2 + 0) * a.width + c] * inv_w;
            const double ex2 = a.moments[((long long)b * 2 + 1) * a.width + c] * inv_w;
            const double bv = ex2 - bm * bm;
            const double delta = bm - mean, tot = count + bc;
            const double m2 = var * count + bv * bc + delta * delta * count * bc / tot;
            mean = mean + delta * bc / tot;
            var = m2 / tot;
            count = tot;
        }
        a.mean[c] = mean;
        a.var[c] = var;
        a.mean32[c] = (float)mean;
        a.std32[c] = sqrtf((float)(var + a.epsilon));
    }
    __syncthreads();                                            // every column has read the old count
    if (c == 0) *a.count = *a.count + (double)a.num_batches * (double)a.B * (double)a.world_size;
    if (a.prior != nullptr && c < DU_C)
        a.prior[c] = a.prior[c] * (1.f - a.prior_soft_coef) + (a.prior_batch[c] * (float)inv_w) * a.prior_soft_coef;
    if (a.std != nullptr && a.min_std != nullptr && c < a.num_std) a.std[c] = fmaxf(a.std[c], a.min_std[c]);
}

extern "C" int qa_norm_merge(const QaNormMergeArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->width <= 0 || a->width > 128 || a->num_batches <= 0 || a->world_size <= 0 || a->num_std > 128) return QA_EINVAL;
    const void* need[] = {a->moments, a->mean, a->var, a->count, a->mean32, a->std32};
    for (const void* p : need) QA_CHECK_PTR(p);
    if (a->prior != nullptr) QA_CHECK_PTR(a->prior_batch);
    k_norm_merge<<<1, 128, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

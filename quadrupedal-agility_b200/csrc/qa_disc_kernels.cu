// Rollout-side element-wise work around the discriminator GEMMs, fused (bbc/rsl_rl/runners/on_policy_runner.py:163-181,
// bbc/rsl_rl/algorithms/discriminator.py:71-118, bbc/rsl_rl/algorithms/gail.py:199-206):
//   K18 qa_disc_input  -- terminal-state patch of the disc obs, 2-step history roll, task-obs weighting, normalisation
//                         -> the (N,98) discriminator input, the history for the replay buffer and the next step
//   K19 qa_disc_reward -- heads -> reward_i / reward_us / reward_ss (float64 cross-entropy on the already soft-maxed
//                         classifier output, as the reference does) -> total reward, time-out bootstrap, fp32 store
// Both are HBM-bound streaming kernels (~1.6 KB and ~0.1 KB per env); they replace ~45 torch launches per env step.
#include "qa_b200.h"
#include "qa_common.cuh"

#define DI_W QA_NUM_OBS_DISC            // 49

// one warp per env; lanes over the 2 x 49 history entries
__global__ void __launch_bounds__(256) k_disc_input(const __grid_constant__ QaDiscInputArgs a) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= a.num_envs) return;
    const bool done = a.dones[e] != 0;
    const float* prev = a.prev_disc + (size_t)e * DI_W;
    const float* next = a.next_disc + (size_t)e * DI_W;
    const float* hp = a.hist_prev + (size_t)e * 2 * DI_W;
    float* hn = a.hist_new + (size_t)e * 2 * DI_W;
    float* hx = a.hist_next + (size_t)e * 2 * DI_W;
    float* x = a.x_norm + (size_t)e * a.x_pitch;
    const float tow = a.task_obs_weight_dev != nullptr ? __ldg(a.task_obs_weight_dev) : a.task_obs_weight;
    if (lane == 0) {
        if (a.rewards_snap != nullptr) a.rewards_snap[e] = a.rewards_in[e];
        if (a.dones_snap != nullptr) a.dones_snap[e] = a.dones[e];
        if (a.time_outs_snap != nullptr) a.time_outs_snap[e] = a.time_outs_in[e];
        if (a.latent_eps_out != nullptr) a.latent_eps_out[e] = a.latent_eps_in[e];
    }
    if (a.latent_c_out != nullptr && lane < QA_DIM_C) a.latent_c_out[(size_t)e * QA_DIM_C + lane] = a.latent_c_in[(size_t)e * QA_DIM_C + lane];
    for (int i = lane; i < 2 * DI_W; i += 32) {
        const int slot = i >= DI_W ? 1 : 0, k = i - slot * DI_W;
        const float nx = next[k];
        // slot 0 = previous newest entry; slot 1 = this step's disc obs, or the terminal state of a reset env (:166-171)
        const float v = slot == 0 ? hp[DI_W + k] : (done ? prev[k] : nx);
        hn[i] = v;
        hx[i] = done ? nx : v;                                               // fresh episodes restart their history (:180-181)
        float o = v;
        if (a.task_obs_weight_decay && ((k >= 3 && k < 9) || k >= 33)) o = o * tow;   // discriminator.py:76-78
        if (a.obs_disc_weight_step != 0.f) o = o * ((float)slot * a.obs_disc_weight_step + 1.f);    // :80-84
        o = (o - a.norm_mean[i]) / a.norm_std[i];                            // Normalizer.normalize_torch, utils.py:97-103
        x[i] = fminf(fmaxf(o, -a.norm_clip), a.norm_clip);
    }
}

extern "C" int qa_disc_input(const QaDiscInputArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->num_envs < 0) return QA_EINVAL;
    if (a->num_envs == 0) return 0;
    const void* need[] = {a->dones, a->prev_disc, a->next_disc, a->hist_prev, a->hist_new, a->hist_next, a->x_norm,
                          a->norm_mean, a->norm_std};
    for (const void* p : need) QA_CHECK_PTR(p);
    if (a->x_pitch < 2 * DI_W) return QA_EINVAL;
    if (a->rewards_snap != nullptr) QA_CHECK_PTR(a->rewards_in);
    if (a->time_outs_snap != nullptr) QA_CHECK_PTR(a->time_outs_in);
    if (a->latent_eps_out != nullptr) QA_CHECK_PTR(a->latent_eps_in);
    if (a->latent_c_out != nullptr) QA_CHECK_PTR(a->latent_c_in);
    k_disc_input<<<(a->num_envs + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// thread per env
__global__ void __launch_bounds__(256) k_disc_reward(const __grid_constant__ QaDiscRewardArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.num_envs) return;
    const float* h = a.heads + (size_t)e * a.heads_pitch;                    // [d | eps | classifier logits dim_c]
    const float d = h[0], eps = h[1];
    // forward(): c = clamp(softmax(classifier(x)), 1e-20, inf)  (discriminator.py:64-69)
    float lg[QA_DIM_C], c[QA_DIM_C];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < QA_DIM_C; ++k) {
        lg[k] = h[2 + k];
        m = fmaxf(m, lg[k]);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < QA_DIM_C; ++k) {
        c[k] = expf(lg[k] - m);
        s += c[k];
    }
    // labels from the observation row the policy saw (:72-73)
    const float* lab = a.obs + (size_t)e * a.obs_pitch + a.obs_width - QA_DIM_C;
    const float label_eps = lab[-1];
    int arg = 0;
    float best = lab[0];
#pragma unroll
    for (int k = 1; k < QA_DIM_C; ++k)
        if (lab[k] > best) best = lab[k], arg = k;                           // torch.argmax: first maximum
    // reward_ss = -CrossEntropy(c, one_hot) with c ALREADY soft-maxed -> soft-maxed again, in float64 (:108)
    double cm = -1e300, cs = 0.0, c_arg = 0.0;
    double cd[QA_DIM_C];
#pragma unroll
    for (int k = 0; k < QA_DIM_C; ++k) {
        cd[k] = (double)fmaxf(c[k] / s, 1e-20f);
        cm = fmax(cm, cd[k]);
    }
#pragma unroll
    for (int k = 0; k < QA_DIM_C; ++k) {
        cs += exp(cd[k] - cm);
        if (k == arg) c_arg = cd[k];
    }
    const double log_softmax_arg = (c_arg - cm) - log(cs);
    const float r_i = fmaxf(1.f - 0.25f * ((d - 1.f) * (d - 1.f)), 0.f) * a.dt;      // MSELoss mapping (:97-98)
    const float r_us = -fabsf(eps - label_eps) * a.dt;                               // :106
    const double r_ss = log_softmax_arg * (double)a.dt;                              // -(-log p)
    const float r_t = a.reward_t[e];
    double total = (double)(a.coef_i * r_i + a.coef_us * r_us) + (double)a.coef_ss * r_ss + (double)(a.coef_t * r_t);
    if (a.time_outs != nullptr)                                                      // gail.py:203-205, in float64
        total += (double)(a.gamma * (a.values[(size_t)e * a.values_pitch] * (a.time_outs[e] ? 1.f : 0.f)));
    a.rewards_out[e] = (float)total;                                                 // storage.rewards[step].copy_ (:67)
    if (a.dones_out != nullptr) a.dones_out[e] = a.dones[e];
    if (a.reward_terms != nullptr) {
        a.reward_terms[(size_t)e * 4 + 0] = r_i;
        a.reward_terms[(size_t)e * 4 + 1] = r_us;
        a.reward_terms[(size_t)e * 4 + 2] = (float)r_ss;
        a.reward_terms[(size_t)e * 4 + 3] = r_t;
    }
}

extern "C" int qa_disc_reward(const QaDiscRewardArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->num_envs < 0) return QA_EINVAL;
    if (a->num_envs == 0) return 0;
    const void* need[] = {a->heads, a->obs, a->reward_t, a->rewards_out};
    for (const void* p : need) QA_CHECK_PTR(p);
    if (a->heads_pitch < 2 + QA_DIM_C || a->obs_width < QA_DIM_C + 1 || a->obs_pitch < a->obs_width) return QA_EINVAL;
    if (a->time_outs != nullptr) QA_CHECK_PTR(a->values);
    if (a->dones_out != nullptr) QA_CHECK_PTR(a->dones);
    k_disc_reward<<<(a->num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

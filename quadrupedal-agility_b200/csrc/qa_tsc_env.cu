// K16 / K17: the TSC (agility teacher) post-physics step, tsc/legged_gym/envs/base/legged_robot.py:226-298, as two
// kernels around the physics step that the reference takes inside reset_idx (:381-384).  See include/qa_b200.h for the
// split.  Warp per env: lanes = rigid bodies (contact-force norms, ballots), height-scan points, DOFs and row columns;
// the per-env scalar program (goals, yaw errors, termination, the 8 reward terms, reset draws) is warp-uniform.
// Compiled with -fmad=false and written in the reference's op order (masks and indices bit-exact, floats to fp32
// rounding).  HBM traffic per env-step: pre ~2.3 KB read + 0.7 KB written, post ~3.1 KB read + 8.4 KB written
// (obs 3200 B + obs_bbc 2684 B + history 2280 B + obs_disc 196 B).
#include "qa_b200.h"
#include "qa_common.cuh"

#define TSC_WARPS 4
#define TSC_THREADS (TSC_WARPS * 32)
#define TSC_PROP 57
#define TSC_HIST_W 570
#define TSC_BBC_W 671
#define SITE_TSC_RESET 40

struct TscWorkspace {
    double sums[QA_TSC_NUM_REWARDS];
    unsigned int reset_count;
    unsigned int ticket;
};

// torch.remainder(a, b) for b > 0 (c10: fmod, then shift negative results by b)
__device__ __forceinline__ float remainder_pos(float a, float b) {
    float m = fmodf(a, b);
    if (m != 0.f && m < 0.f) m = m + b;
    return m;
}

__device__ __forceinline__ float wrap_pi(float d) {               // (d + pi) % (2 pi) - pi   (:448-451, :1795)
    const float PI_F = 3.14159265358979323846f;
    return remainder_pos(d + PI_F, 2.f * PI_F) - PI_F;
}

__device__ __forceinline__ float tsc_height(const QaTerrain& t, Quat yq, const float* R, const float* hp) {
    // _get_heights (:1708-1755): quat_apply_yaw(point) + root pos, + border, / scale, .long(), clip, min of 3 samples
    const Vec3 p = quat_apply(yq, Vec3{hp[0], hp[1], hp[2]});
    float wx = p.x + R[0], wy = p.y + R[1];
    wx = wx + t.border_size;
    wy = wy + t.border_size;
    long long ix = (long long)(wx / t.horizontal_scale);
    long long iy = (long long)(wy / t.horizontal_scale);
    ix = ix < 0 ? 0 : (ix > t.rows - 2 ? t.rows - 2 : ix);
    iy = iy < 0 ? 0 : (iy > t.cols - 2 ? t.cols - 2 : iy);
    const int16_t* hs = t.height_samples;
    const int16_t h1 = __ldg(hs + ix * t.cols + iy), h2 = __ldg(hs + (ix + 1) * t.cols + iy), h3 = __ldg(hs + ix * t.cols + iy + 1);
    int16_t h = h1 < h2 ? h1 : h2;
    h = h < h3 ? h : h3;
    return (float)h * t.vertical_scale;
}

__global__ void __launch_bounds__(TSC_THREADS)
k_tsc_pre(const __grid_constant__ QaTscConst c, const __grid_constant__ QaTscStepArgs a) {
    __shared__ unsigned s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.x * TSC_WARPS + w;
    const int B = c.num_bodies, P = c.num_height_points, G = c.num_goals_total;
    TscWorkspace* ws = reinterpret_cast<TscWorkspace*>(a.workspace);
    if (e < a.num_envs) {
        float R[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) R[k] = a.root_states[(size_t)e * 13 + k];
        long long ep = a.episode_length_buf[e] + 1;                                          // :236
        const Quat q = {R[3], R[4], R[5], R[6]};
        const Vec3 blv = quat_rotate_sgn(q, Vec3{R[7], R[8], R[9]}, -1.f);                    // :241-243
        const Vec3 bav = quat_rotate_sgn(q, Vec3{R[10], R[11], R[12]}, -1.f);
        const Vec3 pg = quat_rotate_sgn(q, Vec3{0.f, 0.f, -1.f}, -1.f);
        float roll, pitch, yaw;
        {
            const float t0 = 2.0f * (q.w * q.x + q.y * q.z), t1 = 1.0f - 2.0f * (q.x * q.x + q.y * q.y);
            roll = atan2f(t0, t1);
            float t2 = 2.0f * (q.w * q.y - q.z * q.x);
            t2 = clampf(t2, -1.f, 1.f);
            pitch = asinf(t2);
            const float t3 = 2.0f * (q.w * q.z + q.x * q.y), t4 = 1.0f - 2.0f * (q.y * q.y + q.z * q.z);
            yaw = atan2f(t3, t4);
        }
        // contact-force norms, lanes = bodies (:247-249, :325, :1840)
        float nrm = 0.f;
        if (lane < B) {
            const float* f = a.contact_forces + ((size_t)e * B + lane) * 3;
            nrm = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        }
        const bool term_hit = (__ballot_sync(QA_FULL, nrm > 1.f) & c.termination_body_mask) != 0u;
        const float n_col = (float)__popc(__ballot_sync(QA_FULL, nrm > 0.1f) & c.penalised_body_mask);
        bool cfilt[4];
        uint8_t lcont[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) lcont[j] = a.last_contacts[e * 4 + j];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ct = __shfl_sync(QA_FULL, nrm, c.feet_indices[j]) > 2.f;
            cfilt[j] = ct || (lcont[j] != 0);
            if (lane == 0) {
                a.last_contacts[e * 4 + j] = ct ? 1 : 0;
                a.contact_filt[e * 4 + j] = cfilt[j] ? 1 : 0;
            }
        }
        // _update_goals (:204-224)
        float timer = a.reach_goal_timer[e];
        long long gidx = a.cur_goal_idx[e];
        if (timer > c.reach_goal_delay_steps) {
            gidx += 1;
            timer = 0.f;
        }
        const float gx = a.cur_goals[e * 3 + 0], gy = a.cur_goals[e * 3 + 1];
        const float ngx = a.next_goals[e * 3 + 0], ngy = a.next_goals[e * 3 + 1];
        const float relx = gx - R[0], rely = gy - R[1];
        const float dgx = R[0] - gx, dgy = R[1] - gy;
        const float dist = sqrtf(dgx * dgx + dgy * dgy);
        const bool reached = dist < c.next_goal_threshold, leave = dist > c.leave_goal_threshold;
        timer = timer + (reached ? 1.f : 0.f);
        const float nrel = sqrtf(relx * relx + rely * rely);
        const float tvx = relx / (nrel + 1e-5f), tvy = rely / (nrel + 1e-5f);
        const float target_yaw = atan2f(tvy, tvx);
        const float nrx = ngx - R[0], nry = ngy - R[1];
        const float nn = sqrtf(nrx * nrx + nry * nry);
        const float next_target_yaw = atan2f(nry / (nn + 1e-5f), nrx / (nn + 1e-5f));
        // 132-point height scan, lanes = points (:632-637, :1708-1755)
        if (a.global_counter % c.update_interval == 0) {
            const Quat yq = yaw_quat(q);
            for (int p = lane; p < P; p += 32)
                a.measured_heights[(size_t)e * P + p] = tsc_height(a.terrain, yq, R, a.height_points + ((size_t)e * P + p) * 3);
        }
        // current obstacle type (:255-258)
        long long ci = gidx < 0 ? 0 : (gidx > G - c.last_goal_repeat - 1 ? G - c.last_goal_repeat - 1 : gidx);
        const long long otype = a.obstacle_types[(size_t)e * c.num_obstacle_types + ci / c.num_goals_per_obstacle];
        // check_termination (:322-346)
        const bool reach_goal_cutoff = gidx >= (long long)(G - c.last_goal_repeat);
        bool reach_last_goal = false;
        if (c.use_camera) {
            const float* lg = a.env_goals + ((size_t)e * G + (G - c.last_goal_repeat)) * 3;
            const float lx = R[0] - lg[0], ly = R[1] - lg[1];
            reach_last_goal = sqrtf(lx * lx + ly * ly) < c.next_goal_threshold;
        }
        const bool time_out = ((float)ep > c.max_episode_length) || reach_goal_cutoff;
        const bool is_reset = term_hit || time_out || (fabsf(roll) > 1.5f) || (fabsf(pitch) > 1.5f) || (R[2] < -0.25f) ||
                              leave || reach_last_goal;
        // reward terms in dir() order (:412-430, :1779-1930)
        float rt[QA_TSC_NUM_REWARDS];
        rt[0] = 0.f;
        rt[3] = 0.f;
        if (a.action_hl_history_buf != nullptr) {                                            // :1847-1859
            const int H = c.hl_hist_len, A = c.hl_action_dim;
            const float* hl = a.action_hl_history_buf + (size_t)e * H * A;
            float d2 = 0.f;
            for (int k = lane; k < A; k += 32) {
                const float d = hl[(H - 2) * A + k] - hl[(H - 1) * A + k];
                d2 += d * d;
            }
            rt[0] = sqrtf(warp_sum(d2));
            rt[3] = 0.5f * (fabsf(hl[(H - 3) * A] - hl[(H - 1) * A]) + fabsf(hl[(H - 2) * A] - hl[(H - 1) * A]));
        }
        rt[1] = n_col;
        {                                                                                    // feet_edge :1899-1915
            bool at_edge = false;
            if (lane < 4) {
                const float* fp = a.rigid_body_state + ((size_t)e * B + c.feet_indices[lane]) * 13;
                long long fx = (long long)rintf((fp[0] + a.terrain.border_size) / a.terrain.horizontal_scale);
                long long fy = (long long)rintf((fp[1] + a.terrain.border_size) / a.terrain.horizontal_scale);
                fx = fx < 0 ? 0 : (fx > a.terrain.rows - 1 ? a.terrain.rows - 1 : fx);
                fy = fy < 0 ? 0 : (fy > a.terrain.cols - 1 ? a.terrain.cols - 1 : fy);
                bool cf_l = false;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (lane == j) cf_l = cfilt[j];
                at_edge = cf_l && (a.x_edge_mask[fx * a.terrain.cols + fy] != 0);
                a.feet_at_edge[e * 4 + lane] = at_edge ? 1 : 0;
            }
            rt[2] = (float)__popc(__ballot_sync(QA_FULL, at_edge) & 0xfu);
        }
        rt[4] = reached ? 1.f : 0.f;                                                         // :1921
        {                                                                                    // tracking_goal_vel :1779-1791
            const float proj = tvx * R[7] + tvy * R[8];
            const float z = a.commands[e * 5] * 0.f;
            const float tgt = (otype == 0 || otype == 4) ? 2.5f : c.target_lin_vel;
            rt[5] = fminf(proj, z + tgt) / (z + tgt + 1e-5f);
        }
        rt[6] = expf(-fabsf(wrap_pi(target_yaw - yaw)));                                     // tracking_yaw :1793-1797
        rt[7] = (is_reset && !time_out) ? 1.f : 0.f;                                         // termination :1917
        float rew = 0.f;
        float es[QA_TSC_NUM_REWARDS];
#pragma unroll
        for (int k = 0; k < QA_TSC_NUM_REWARDS; ++k) es[k] = a.episode_sums[(size_t)e * QA_TSC_NUM_REWARDS + k];
#pragma unroll
        for (int k = 0; k < QA_TSC_NUM_REWARDS - 1; ++k) {
            const float t = rt[k] * c.reward_scale[k];
            rew = rew + t;
            es[k] = es[k] + t;
        }
        if (c.only_positive_rewards) rew = fmaxf(rew, 0.f);
        if (c.reward_scale[QA_TSC_NUM_REWARDS - 1] != 0.f) {                                  // after the clip (:423-430)
            const float t = rt[QA_TSC_NUM_REWARDS - 1] * c.reward_scale[QA_TSC_NUM_REWARDS - 1];
            rew = rew + t;
            es[QA_TSC_NUM_REWARDS - 1] = es[QA_TSC_NUM_REWARDS - 1] + t;
        }
        if (is_reset) {                                                                      // reset_idx :348-410
            if (lane < QA_TSC_NUM_REWARDS) {
                float mine = 0.f;
#pragma unroll
                for (int k = 0; k < QA_TSC_NUM_REWARDS; ++k)
                    if (lane == k) mine = es[k];
                atomicAdd(&ws->sums[lane], (double)mine);
            }
            if (lane == 0) atomicAdd(&ws->reset_count, 1u);
#pragma unroll
            for (int k = 0; k < QA_TSC_NUM_REWARDS; ++k) es[k] = 0.f;
            gidx = 0;                                                                        // randomize_start False (:376)
            ep = 0;
            timer = 0.f;
            if (lane < 12) {                                                                 // _reset_dofs :798-804
                a.dof_state[((size_t)e * 12 + lane) * 2 + 0] = c.default_dof_pos[lane];
                a.dof_state[((size_t)e * 12 + lane) * 2 + 1] = 0.f;
            }
            float u_yaw, u_x, u_y;
            if (a.yaw_u != nullptr) {
                u_yaw = a.yaw_u[e], u_x = a.x_u[e], u_y = a.y_u[e];
            } else {
                const Philox4 r = philox4x32_10((uint32_t)e, SITE_TSC_RESET, (uint32_t)a.rng_step, (uint32_t)(a.rng_step >> 32),
                                                (uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
                u_yaw = u32_to_unit_f32(r.v[0]), u_x = u32_to_unit_f32(r.v[1]), u_y = u32_to_unit_f32(r.v[2]);
            }
            if (lane == 0) {
                if (a.obst_dof_state != nullptr) a.obst_dof_state[a.seesaw_dof_index[e] * 2] = c.seesaw_dof_pos;   // :825-829
                float* Rw = a.root_states + (size_t)e * 13;                                  // _reset_root_states :840-884
#pragma unroll
                for (int k = 0; k < 13; ++k) Rw[k] = c.base_init_state[k];
                const float* g0 = a.env_goals + (size_t)e * G * 3;
                const float rand_yaw = c.rand_yaw_range * (2.f * u_yaw + -1.f);
                const float root_yaw = rand_yaw + c.frame_ang0;
                const float cy = cosf(root_yaw * 0.5f), sy = sinf(root_yaw * 0.5f);
                Rw[0] = g0[0] + c.rand_x_range * (1.f * u_x + -1.f);
                Rw[1] = g0[1] + c.rand_y_range * (2.f * u_y + -1.f);
                Rw[3] = 0.f, Rw[4] = 0.f, Rw[5] = sy, Rw[6] = cy;                            // quat_from_euler_xyz(0, 0, yaw)
            }
        }
        if (lane == 0) {
            a.episode_length_buf[e] = ep;
            a.reach_goal_timer[e] = timer;
            a.cur_goal_idx[e] = gidx;
            a.base_lin_vel[e * 3 + 0] = blv.x, a.base_lin_vel[e * 3 + 1] = blv.y, a.base_lin_vel[e * 3 + 2] = blv.z;
            a.base_ang_vel[e * 3 + 0] = bav.x, a.base_ang_vel[e * 3 + 1] = bav.y, a.base_ang_vel[e * 3 + 2] = bav.z;
            a.projected_gravity[e * 3 + 0] = pg.x, a.projected_gravity[e * 3 + 1] = pg.y, a.projected_gravity[e * 3 + 2] = pg.z;
#pragma unroll
            for (int k = 0; k < 3; ++k) a.base_lin_acc[e * 3 + k] = (R[7 + k] - a.last_root_vel_in[e * 6 + k]) / c.dt;   // :244
            a.rpy[e * 3 + 0] = roll, a.rpy[e * 3 + 1] = pitch, a.rpy[e * 3 + 2] = yaw;
            a.target_yaw[e] = target_yaw;
            a.next_target_yaw[e] = next_target_yaw;
            a.cur_obstacle_types[e] = otype;
            a.reached_goal[e] = reached ? 1 : 0;
            a.reach_goal_cutoff[e] = reach_goal_cutoff ? 1 : 0;
            a.reset_buf[e] = is_reset ? 1 : 0;
            a.time_out_buf[e] = time_out ? 1 : 0;
            a.rew_buf[e] = rew;
        }
        if (lane < QA_TSC_NUM_REWARDS) {
            float mine = 0.f;
#pragma unroll
            for (int k = 0; k < QA_TSC_NUM_REWARDS; ++k)
                if (lane == k) mine = es[k];
            a.episode_sums[(size_t)e * QA_TSC_NUM_REWARDS + lane] = mine;
        }
    }
    // ---- last block finalises: episode reward means (:398-405), time-out latch (:408-410), obst_dof_vel[:] = 0 (:830) ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        const unsigned cnt = *reinterpret_cast<volatile unsigned*>(&ws->reset_count);
        if (cnt > 0) {
            if (threadIdx.x < QA_TSC_NUM_REWARDS) {
                const double s = *reinterpret_cast<volatile double*>(&ws->sums[threadIdx.x]);
                a.episode_rew_means[threadIdx.x] = (float)(s / (double)cnt) / c.episode_length_s;
            }
            const volatile uint8_t* src = a.time_out_buf;
            for (int i = threadIdx.x; i < a.num_envs; i += TSC_THREADS) a.time_outs_latched[i] = src[i];
            if (a.obst_dof_state != nullptr)
                for (long long i = threadIdx.x; i < a.num_obst_dofs; i += TSC_THREADS) a.obst_dof_state[i * 2 + 1] = 0.f;
        }
        __syncthreads();
        if (threadIdx.x < QA_TSC_NUM_REWARDS) ws->sums[threadIdx.x] = 0.0;
        if (threadIdx.x == 0) {
            *a.num_resets = (int)cnt;
            ws->reset_count = 0u;
            ws->ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(TSC_THREADS)
k_tsc_post(const __grid_constant__ QaTscConst c, const __grid_constant__ QaTscStepArgs a) {
    __shared__ float s_prop[TSC_WARPS][64];        // the 57-vector of this step (noise-free; TSC adds no noise, :102)
    __shared__ float s_mid[TSC_WARPS][36];         // priv_explicit 4 + priv_latent 29
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.x * TSC_WARPS + w;
    if (e >= a.num_envs) return;
    const int B = c.num_bodies, P = c.num_height_points, G = c.num_goals_total, L = c.contact_ring_len;
    const bool is_reset = a.reset_buf[e] != 0;
    float* hist = a.obs_history_buf + (size_t)e * TSC_HIST_W;
    float* ahist = a.action_history_buf + (size_t)e * QA_ACT_HIST_LEN * 12;
    if (is_reset) {                                                                          // :386-396
        if (lane < 4) a.feet_air_time[e * 4 + lane] = 0.f;
        for (int i = lane; i < TSC_HIST_W; i += 32) hist[i] = 0.f;
        for (int i = lane; i < L * 4; i += 32) a.contact_buf[(size_t)e * L * 4 + i] = 0.f;
        for (int i = lane; i < QA_ACT_HIST_LEN * 12; i += 32) ahist[i] = 0.f;
        __syncwarp();
    }
    const long long gidx = a.cur_goal_idx[e];                                                // :271-272
    if (lane < 3) {
        a.cur_goals[e * 3 + lane] = a.env_goals[((size_t)e * G + gidx) * 3 + lane];
        a.next_goals[e * 3 + lane] = a.env_goals[((size_t)e * G + gidx + 1) * 3 + lane];
    }
    // ---- compute_observations (:432-515) ----
    float R[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) R[k] = a.root_states[(size_t)e * 13 + k];
    const float* mh = a.measured_heights + (size_t)e * P;
    const float root_h = R[2] - mh[P / 2 + 1];
    const float roll = a.rpy[e * 3 + 0], pitch = a.rpy[e * 3 + 1], yaw = a.rpy[e * 3 + 2];
    float dyaw, dnyaw;
    if (a.global_counter % c.update_interval == 0) {                                         // :446-451
        dyaw = wrap_pi(a.target_yaw[e] - yaw);
        dnyaw = wrap_pi(a.next_target_yaw[e] - yaw);
        if (lane == 0) a.delta_yaw[e] = dyaw, a.delta_next_yaw[e] = dnyaw;
    } else {
        dyaw = a.delta_yaw[e], dnyaw = a.delta_next_yaw[e];
    }
    float* prop = s_prop[w];
    float* mid = s_mid[w];
    float* disc = a.obs_disc_buf + (size_t)e * QA_NUM_OBS_DISC;
    const float blv_l = lane < 3 ? a.base_lin_vel[e * 3 + lane] : 0.f;
    const float bav_l = lane < 3 ? a.base_ang_vel[e * 3 + lane] : 0.f;
    if (lane < 12) {
        const float dof_pos = a.dof_state[((size_t)e * 12 + lane) * 2], dof_vel = a.dof_state[((size_t)e * 12 + lane) * 2 + 1];
        const float dq = (dof_pos - c.default_dof_pos[lane]) * c.s_dof_pos, dv = dof_vel * c.s_dof_vel;
        prop[5 + lane] = dq;
        prop[17 + lane] = dv;
        prop[29 + lane] = ahist[(QA_ACT_HIST_LEN - 1) * 12 + lane];
        disc[9 + lane] = dq;
        disc[21 + lane] = dv;
        a.last_actions[(size_t)e * 12 + lane] = a.actions[(size_t)e * 12 + lane];            // :278-280
        a.last_dof_vel[(size_t)e * 12 + lane] = dof_vel;
        a.last_torques_org[(size_t)e * 12 + lane] = a.torques_org[(size_t)e * 12 + lane];
        mid[9 + lane] = a.motor_strength[(size_t)e * 12 + lane] - 1.f;
        mid[21 + lane] = a.motor_strength[(size_t)a.num_envs * 12 + (size_t)e * 12 + lane] - 1.f;
    }
    if (lane < 6) a.last_root_vel[e * 6 + lane] = R[7 + lane];                               // :281
    if (lane < 4) {
        // compute_flat_key_pos (:1925-1947) on the post-reset root and the refreshed rigid-body tensor
        const Quat hq = heading_quat_inv(Quat{R[3], R[4], R[5], R[6]});
        const float* kp = a.rigid_body_state + ((size_t)e * B + c.feet_indices[lane]) * 13;
        const Vec3 o = quat_rotate_sgn(hq, Vec3{kp[0] - R[0], kp[1] - R[1], kp[2] - R[2]}, 1.f);
        const float cf = a.contact_filt[e * 4 + lane] ? 1.f : 0.f;
        disc[33 + lane * 3 + 0] = o.x * c.s_key_pos, disc[33 + lane * 3 + 1] = o.y * c.s_key_pos, disc[33 + lane * 3 + 2] = o.z * c.s_key_pos;
        disc[45 + lane] = cf * c.s_foot_contact;
        prop[41 + lane] = cf - 0.5f;
        prop[45 + lane * 3 + 0] = o.x * 0.f, prop[45 + lane * 3 + 1] = o.y * 0.f, prop[45 + lane * 3 + 2] = o.z * 0.f;
        a.contact_buf[((size_t)e * L + a.contact_ring_head) * 4 + lane] = clampf(cf, -c.clip_obs, c.clip_obs);   // :509, :514
        mid[4 + lane] = a.mass_params[e * 4 + lane];
    }
    if (lane < 3) {
        prop[2 + lane] = bav_l * c.s_ang_vel;
        disc[3 + lane] = blv_l * c.s_lin_vel_dist;
        disc[6 + lane] = bav_l * c.s_ang_vel_dist;
        mid[1 + lane] = blv_l * c.s_lin_vel;
    }
    if (lane == 0) {
        prop[0] = roll, prop[1] = pitch;
        disc[0] = roll, disc[1] = pitch, disc[2] = root_h;
        mid[0] = c.root_height_obs ? root_h : 0.f;
        mid[9 - 1] = a.friction_coeffs[e];        // slot 8: [root_h | lin vel 3 | mass 4 | friction 1 | ms_p 12 | ms_d 12]
    }
    __syncwarp();
    const float cl = c.clip_obs;
    float* obs = a.obs_buf + (size_t)e * QA_TSC_OBS;
    float* bbc = a.obs_bbc_buf + (size_t)e * TSC_BBC_W;
    const long long otype = a.cur_obstacle_types[e];
    // obs = [57 | delta_yaws 2 | obstacle one-hot 6 | heights 132 | priv explicit 4 | priv latent 29 | history 570]
    for (int i = lane; i < TSC_PROP; i += 32) {
        const float v = clampf(prop[i], -cl, cl);
        obs[i] = v;
        bbc[i] = v;
    }
    if (lane == 0) obs[57] = clampf(dyaw, -cl, cl), obs[58] = clampf(dnyaw, -cl, cl);
    if (lane < c.num_obstacle_types) obs[59 + lane] = (otype == lane) ? 1.f : 0.f;
    const int off_h = TSC_PROP + QA_TSC_AUX;
    for (int p = lane; p < P; p += 32) obs[off_h + p] = clampf(clampf(R[2] - 0.3f - mh[p], -1.f, 1.f), -cl, cl);
    for (int i = lane; i < 33; i += 32) {
        const float v = clampf(mid[i], -cl, cl);
        obs[off_h + P + i] = v;
        bbc[TSC_PROP + i] = v;
    }
    // history BEFORE the update goes into both rows (:484-507); then fill / shift (:499-507), staged through registers
    constexpr int NH = (TSC_HIST_W + 31) / 32;
    float hv[NH];
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        const int i = k * 32 + lane;
        hv[k] = i < TSC_HIST_W ? hist[i] : 0.f;
    }
    const bool fill = a.episode_length_buf[e] <= 1;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        const int i = k * 32 + lane;
        if (i < TSC_HIST_W) {
            const float v = clampf(hv[k], -cl, cl);
            obs[off_h + P + 33 + i] = v;
            bbc[TSC_PROP + 33 + i] = v;
            if (!fill && i >= TSC_PROP) hist[i - TSC_PROP] = v;                              // stored values are clipped already
        }
    }
    if (fill) {
        for (int i = lane; i < TSC_HIST_W; i += 32) hist[i] = clampf(prop[i % TSC_PROP], -cl, cl);
    } else {
        for (int i = lane; i < TSC_PROP; i += 32) hist[TSC_HIST_W - TSC_PROP + i] = clampf(prop[i], -cl, cl);
    }
    if (lane < 11) {
        const float v = lane < 5 ? a.commands[e * 5 + lane] : (lane == 5 ? a.latent_eps[e] : a.latent_c[e * 5 + lane - 6]);
        bbc[TSC_PROP + 33 + TSC_HIST_W + lane] = clampf(v, -cl, cl);
    }
}

static int tsc_validate(const QaTscConst* c, const QaTscStepArgs* a) {
    if (c == nullptr || a == nullptr) return QA_EINVAL;
    if (a->num_envs < 0) return QA_EINVAL;
    if (c->num_bodies <= 0 || c->num_bodies > 32) return QA_ERANGE;
    for (int j = 0; j < 4; ++j)
        if (c->feet_indices[j] < 0 || c->feet_indices[j] >= c->num_bodies) return QA_ERANGE;
    if (c->update_interval <= 0 || c->num_goals_per_obstacle <= 0 || c->num_goals_total <= c->last_goal_repeat) return QA_EINVAL;
    if (c->num_height_points <= 0 || c->num_obstacle_types <= 0 || c->num_obstacle_types > 32) return QA_ERANGE;
    if (c->contact_ring_len <= 0) return QA_EINVAL;
    return 0;
}

extern "C" int qa_post_physics_tsc_pre(const QaTscConst* c, const QaTscStepArgs* a, void* stream) {
    int rc = tsc_validate(c, a);
    if (rc != 0) return rc;
    if (a->num_envs == 0) return 0;
    const void* need[] = {a->root_states, a->dof_state, a->rigid_body_state, a->contact_forces, a->terrain.height_samples,
                          a->x_edge_mask, a->height_points, a->env_goals, a->obstacle_types, a->episode_length_buf,
                          a->last_root_vel_in, a->last_contacts, a->reach_goal_timer, a->cur_goal_idx, a->cur_goals,
                          a->next_goals, a->commands, a->episode_sums, a->measured_heights, a->base_lin_vel, a->base_ang_vel,
                          a->projected_gravity, a->base_lin_acc, a->rpy, a->contact_filt, a->target_yaw, a->next_target_yaw,
                          a->cur_obstacle_types, a->reached_goal, a->reach_goal_cutoff, a->feet_at_edge, a->reset_buf,
                          a->time_out_buf, a->time_outs_latched, a->rew_buf, a->episode_rew_means, a->num_resets, a->workspace};
    for (const void* p : need) QA_CHECK_PTR(p);
    if (a->terrain.rows < 2 || a->terrain.cols < 2) return QA_EINVAL;
    if (a->obst_dof_state != nullptr) QA_CHECK_PTR(a->seesaw_dof_index);
    if (a->action_hl_history_buf != nullptr && (c->hl_hist_len < 3 || c->hl_action_dim <= 0)) return QA_EINVAL;
    const bool any_u = a->yaw_u || a->x_u || a->y_u;
    if (any_u && !(a->yaw_u && a->x_u && a->y_u)) return QA_EINVAL;
    k_tsc_pre<<<(a->num_envs + TSC_WARPS - 1) / TSC_WARPS, TSC_THREADS, 0, (cudaStream_t)stream>>>(*c, *a);
    QA_LAUNCH_RET();
}

extern "C" int qa_post_physics_tsc_post(const QaTscConst* c, const QaTscStepArgs* a, void* stream) {
    int rc = tsc_validate(c, a);
    if (rc != 0) return rc;
    if (a->num_envs == 0) return 0;
    const void* need[] = {a->root_states, a->dof_state, a->rigid_body_state, a->env_goals, a->mass_params, a->friction_coeffs,
                          a->motor_strength, a->episode_length_buf, a->cur_goal_idx, a->cur_goals, a->next_goals, a->actions,
                          a->torques_org, a->last_actions, a->last_dof_vel, a->last_torques_org, a->last_root_vel, a->commands,
                          a->latent_eps, a->latent_c, a->feet_air_time, a->obs_history_buf, a->action_history_buf,
                          a->contact_buf, a->measured_heights, a->delta_yaw, a->delta_next_yaw, a->base_lin_vel,
                          a->base_ang_vel, a->rpy, a->contact_filt, a->target_yaw, a->next_target_yaw, a->cur_obstacle_types,
                          a->reset_buf, a->obs_buf, a->obs_bbc_buf, a->obs_disc_buf};
    for (const void* p : need) QA_CHECK_PTR(p);
    if (a->contact_ring_head < 0 || a->contact_ring_head >= c->contact_ring_len) return QA_ERANGE;
    k_tsc_post<<<(a->num_envs + TSC_WARPS - 1) / TSC_WARPS, TSC_THREADS, 0, (cudaStream_t)stream>>>(*c, *a);
    QA_LAUNCH_RET();
}

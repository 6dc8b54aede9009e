// K2: fused BBC post-physics step (sm_100a).
//
// Replaces LeggedRobot.post_physics_step and everything it calls
// (bbc/legged_gym/envs/base/legged_robot.py:124-166, 168-176, 178-240, 242-259, 261-331,
//  449-540, 598-612, 660-687, 1248-1335, 1377-1396) with ONE kernel.
//
// Mapping: one warp per env, QA_K2_ENVS_PER_CTA envs per CTA.  Per env the kernel moves ~3 KB in and
// ~8 KB out (SURVEY.md 8d byte table) and does a few hundred flops, so it is HBM/latency bound:
//   * all per-env inputs are fetched with coalesced warp loads into a per-warp shared-memory
//     staging area up front (one round trip), scalars are then broadcast-read from smem;
//   * per-DOF quantities live in lanes 0..11 and the nine 12-wide reward sums are warp-shuffle
//     butterflies;
//   * the 671-float observation row is assembled in a CTA-contiguous smem tile (history rows 1..9 are
//     loaded straight into their shifted position), noise + clip are applied in one pass, and the
//     tile goes out either with coalesced warp stores or, when the CTA's rows are contiguous and
//     16B aligned, with two TMA bulk stores (cp.async.bulk.global.shared::cta -> UBLKCP);
//   * reset envs (warp-uniform branch) resample commands/latents, blend their mocap frame from the
//     L2-resident clip table and rewrite simulator state in place;
//   * episode statistics of the reset envs are reduced with fp64 atomics and finalised by the last
//     CTA (threadfence ticket), which also latches extras["time_outs"].
//
// Compiled with -fmad=false: the arithmetic is written in the reference's op order so that the
// quantised terrain index and all masks are bit-exact against the oracle.

#include "qa_k2_common.cuh"

#define K2_ENVS 4                       // envs (= warps) per CTA
#define K2_THREADS (K2_ENVS * 32)

template <bool BULK>
__global__ void __launch_bounds__(K2_THREADS)
k_post_physics_bbc(const __grid_constant__ QaBbcConst c, const __grid_constant__ QaBbcStepArgs a_in) {
    const K2Step a(a_in);
    __shared__ __align__(16) float s_tile[K2_ENVS * ROW];          // obs rows of this CTA, contiguous
    __shared__ __align__(16) float s_stage[K2_ENVS][S_TOTAL];
    __shared__ float s_noise[ROW];
    __shared__ int s_last;

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int e = blockIdx.x * K2_ENVS + wid;
    const bool active = e < a.num_envs;
    const int B = c.num_bodies;

    for (int i = threadIdx.x; i < ROW; i += K2_THREADS) s_noise[i] = c.add_noise ? a.noise_scale_vec[i] : 0.f;
    __syncthreads();

    float* row = s_tile + wid * ROW;
    float* st = s_stage[wid];
    bool is_reset = false;

    if (active) {
        // ---------------- stage 1: issue every per-env load ------------------------------------
        const float* g_root = a.root_states + (size_t)e * 13;
        if (lane < 13) st[S_ROOT + lane] = g_root[lane];
        if (lane < 5) st[S_CMD + lane] = a.commands[e * 5 + lane];
        if (lane == 5) st[S_CMD + 5] = a.latent_eps[e];
        if (lane >= 6 && lane < 11) st[S_CMD + lane] = a.latent_c[e * QA_DIM_C + lane - 6];
        if (lane < 12) {
            const int j = lane / 3, k = lane - j * 3;
            st[S_KEY + lane] = a.rigid_body_state[((size_t)e * B + c.feet_indices[j]) * 13 + k];
        }
        if (lane < 4) st[S_MISC + lane] = a.mass_params[e * 4 + lane];
        if (lane == 4) st[S_MISC + 4] = a.friction_coeffs[e];
        for (int i = lane; i < B * 3; i += 32) st[S_CF + i] = a.contact_forces[(size_t)e * B * 3 + i];

        float dof_pos = 0.f, dof_vel = 0.f, act = 0.f, lact = 0.f, tq = 0.f, ltq = 0.f, ldv = 0.f, msp = 0.f,
              msd = 0.f, hist_last_action = 0.f;
        const size_t nd = (size_t)a.num_envs * QA_NUM_DOF;
        if (lane < QA_NUM_DOF) {
            const float2 pv = reinterpret_cast<const float2*>(a.dof_state)[e * QA_NUM_DOF + lane];
            dof_pos = pv.x;
            dof_vel = pv.y;
            act = a.actions[e * QA_NUM_DOF + lane];
            lact = a.last_actions[e * QA_NUM_DOF + lane];
            tq = a.torques_org[e * QA_NUM_DOF + lane];
            ltq = a.last_torques_org[e * QA_NUM_DOF + lane];
            ldv = a.last_dof_vel[e * QA_NUM_DOF + lane];
            msp = a.motor_strength[e * QA_NUM_DOF + lane];
            msd = a.motor_strength[nd + e * QA_NUM_DOF + lane];
            hist_last_action =
                a.action_history_buf[(size_t)e * QA_ACT_HIST_LEN * QA_NUM_DOF + (QA_ACT_HIST_LEN - 1) * QA_NUM_DOF + lane];
        }
        float epsum = 0.f;
        if (lane < QA_NUM_REWARDS) epsum = a.episode_sums[(size_t)e * QA_EPSUM_PITCH + lane];
        const long long ep_in = a.episode_length_buf[e];
        const uint32_t lc_bits = reinterpret_cast<const uint32_t*>(a.last_contacts)[e];
        // history rows 1..9 land directly in their shifted position inside the obs row
        {
            const float* gh = a.obs_history_buf + (size_t)e * HIST_W + QA_NUM_PROP;
            for (int i = lane; i < HIST_W - QA_NUM_PROP; i += 32) row[HIST_OFF + i] = gh[i];
        }
        __syncwarp();

        // ---------------- stage 2: base-frame quantities (:133-146) --------------------------------
        long long ep = ep_in + 1;
        const Quat q = {st[S_ROOT + 3], st[S_ROOT + 4], st[S_ROOT + 5], st[S_ROOT + 6]};
        const Vec3 blv = quat_rotate_sgn(q, Vec3{st[S_ROOT + 7], st[S_ROOT + 8], st[S_ROOT + 9]}, -1.f);
        const Vec3 bav = quat_rotate_sgn(q, Vec3{st[S_ROOT + 10], st[S_ROOT + 11], st[S_ROOT + 12]}, -1.f);
        const Vec3 pg = quat_rotate_sgn(q, Vec3{0.f, 0.f, -1.f}, -1.f);
        float roll, pitch, yaw;
        {
            const float t0 = 2.0f * (q.w * q.x + q.y * q.z);
            const float t1 = 1.0f - 2.0f * (q.x * q.x + q.y * q.y);
            roll = atan2f(t0, t1);
            float t2 = 2.0f * (q.w * q.y - q.z * q.x);
            t2 = clampf(t2, -1.f, 1.f);
            pitch = asinf(t2);
            const float t3 = 2.0f * (q.w * q.z + q.x * q.y);
            const float t4 = 1.0f - 2.0f * (q.y * q.y + q.z * q.z);
            yaw = atan2f(t3, t4);
        }
        float feet_force[4];
        bool contact[4], cfilt[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* f = st + S_CF + c.feet_indices[j] * 3;
            feet_force[j] = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
            contact[j] = feet_force[j] > 2.f;
            cfilt[j] = contact[j] || (((lc_bits >> (8 * j)) & 0xffu) != 0u);
        }

        // ---------------- stage 3: callback (:449-472) ---------------------------------------------
        if (ep % (long long)c.resample_period == 0) {
            const K2Draw d = draw_site(c, a, e, SITE_RS0, a.rs_eps_u, a.rs_c_idx, a.rs_cmd_u);
            resample_env(c, d, st + S_CMD, lane);
        }
        float center_h = 0.f;
        if (c.measure_heights) {
            const Quat yq = yaw_quat(q);
            center_h = terrain_center_height(a.terrain, yq, st[S_ROOT + 0], st[S_ROOT + 1], c.center_px, c.center_py);
        }
        float root_vx = st[S_ROOT + 7], root_vy = st[S_ROOT + 8];
        if (a.do_push) {                                                       // _push_robots :682-687
            float u0, u1;
            if (a.push_u != nullptr) {
                u0 = a.push_u[e * 2 + 0];
                u1 = a.push_u[e * 2 + 1];
            } else {
                Philox4 r = philox4x32_10((uint32_t)e, SITE_PUSH, (uint32_t)a.rng_step, (uint32_t)(a.rng_step >> 32),
                                          (uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
                u0 = u32_to_unit_f32(r.v[0]);
                u1 = u32_to_unit_f32(r.v[1]);
            }
            const float span = c.max_push_vel_xy - (-c.max_push_vel_xy);
            root_vx = span * u0 + (-c.max_push_vel_xy);
            root_vy = span * u1 + (-c.max_push_vel_xy);
        }
        const float root_z_pre = st[S_ROOT + 2];

        // ---------------- stage 4: termination (:168-176) ------------------------------------------
        bool term = false;
        for (int b = lane; b < B; b += 32) {
            if ((c.termination_body_mask >> b) & 1u) {
                const float* f = st + S_CF + b * 3;
                term = term || (sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 1.f);
            }
        }
        term = __any_sync(QA_FULL, term);
        const bool time_out = ((float)ep > c.max_episode_length) || (root_z_pre < -6.0f);
        is_reset = term || time_out;

        // ---------------- stage 5: rewards (:242-259, :1248-1335) ----------------------------------
        const float* cmd = st + S_CMD;
        const float root_h_pre = root_z_pre - center_h;
        float r_terms[QA_NUM_REWARDS];
        {
            const bool dl = lane < QA_NUM_DOF;
            const int d_ = dl ? lane : 0;
            float v;
            v = lact - act;
            r_terms[0] = warp_sum(dl ? v * v : 0.f);                                  // action_rate
            int ncol = 0;
            for (int b = lane; b < B; b += 32) {
                if ((c.penalised_body_mask >> b) & 1u) {
                    const float* f = st + S_CF + b * 3;
                    ncol += (sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 0.1f) ? 1 : 0;
                }
            }
            r_terms[1] = warp_sum((float)ncol);                                       // collision
            v = tq - ltq;
            r_terms[2] = warp_sum(dl ? v * v : 0.f);                                  // delta_torques
            v = (ldv - dof_vel) / c.dt;
            r_terms[3] = warp_sum(dl ? v * v : 0.f);                                  // dof_acc
            const float dq0 = dof_pos - c.default_dof_pos[d_];
            r_terms[4] = warp_sum(dl ? dq0 * dq0 : 0.f);                              // dof_error
            v = -fminf(dof_pos - c.dof_pos_lower[d_], 0.f);
            v = v + fmaxf(dof_pos - c.dof_pos_upper[d_], 0.f);
            r_terms[5] = warp_sum(dl ? v : 0.f);                                      // dof_pos_limits
            v = clampf(fabsf(dof_vel) - c.dof_vel_limits[d_] * c.soft_dof_vel_limit, 0.f, 1.f);
            r_terms[6] = warp_sum(dl ? v : 0.f);                                      // dof_vel_limits
            r_terms[7] = warp_sum((dl && ((c.hip_dof_mask >> d_) & 1u)) ? dq0 * dq0 : 0.f);   // hip_pos
            {
                const float err = sqrtf((cmd[3] - root_h_pre) * (cmd[3] - root_h_pre));
                r_terms[8] = ((err < 0.05f) && (cmd[3] >= c.jump_height_lo)) ? c.jump_goal : 0.f;   // jump_up_height
            }
            {
                const float err = sqrtf((cmd[4] - root_h_pre) * (cmd[4] - root_h_pre));
                const float rl = expf(-10.0f * (err * err) / c.tracking_sigma);
                r_terms[9] = (!(cmd[3] > c.jump_height_lo)) ? rl : 0.f;                // locomotion_height
            }
            v = fmaxf(fabsf(tq) - c.torque_limits[d_] * c.soft_torque_limit, 0.f);
            r_terms[10] = warp_sum(dl ? v : 0.f);                                     // torque_limits
            r_terms[11] = warp_sum(dl ? tq * tq : 0.f);                               // torques
            {
                const float dw = cmd[2] - bav.z;
                r_terms[12] = expf(-(dw * dw) / c.tracking_sigma);                    // tracking_ang_vel
                const float dx = cmd[0] - blv.x, dy = cmd[1] - blv.y;
                r_terms[13] = expf(-(dx * dx + dy * dy) / c.tracking_sigma);          // tracking_lin_vel
            }
        }
        float rew = 0.f, my_term = 0.f;
#pragma unroll
        for (int k = 0; k < QA_NUM_REWARDS; ++k) {
            const float t = r_terms[k] * c.reward_scale[k];
            rew = rew + t;
            if (lane == k) my_term = t;
        }
        if (c.only_positive_rewards) rew = fmaxf(rew, 0.f);
        epsum = epsum + my_term;

        // ---------------- stage 6: reset_idx (:178-240) --------------------------------------------
        float root_out[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) root_out[k] = st[S_ROOT + k];
        root_out[7] = root_vx;
        root_out[8] = root_vy;
        float feet_air_zero = 0.f;
        (void)feet_air_zero;
        if (is_reset) {
            K2Workspace* ws = reinterpret_cast<K2Workspace*>(a.workspace);
            if (lane < QA_NUM_REWARDS) atomicAdd(&ws->sums[lane], (double)epsum);
            if (lane == 0) atomicAdd(&ws->reset_count, 1u);
            epsum = 0.f;

            const K2Draw d = draw_site(c, a, e, SITE_RT0, a.rt_eps_u, a.rt_c_idx, a.rt_cmd_u);
            resample_env(c, d, st + S_CMD, lane);

            // mocap reference-state initialisation
            int clip;
            double time_u;
            if (a.mocap_clip_idx != nullptr) {
                clip = a.mocap_clip_idx[e];
                time_u = a.mocap_time_u[e];
            } else {
                Philox4 r = philox4x32_10((uint32_t)e, SITE_MOCAP, (uint32_t)a.rng_step, (uint32_t)(a.rng_step >> 32),
                                          (uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
                const double cu = u64_to_unit_f64(r.v[0], r.v[1]);
                time_u = u64_to_unit_f64(r.v[2], r.v[3]);
                const int lo = a.mocap.mode_offset[d.c_idx], hi = a.mocap.mode_offset[d.c_idx + 1];
                int j = lo;
                while (j < hi - 1 && a.mocap.mode_cdf[j] <= cu) ++j;
                clip = a.mocap.mode_clips[j];
            }
            clip = min(max(clip, 0), a.mocap.num_clips - 1);
            const MocapBlendIdx bi = mocap_blend_index(a.mocap, clip, time_u, c.time_between_frames, c.disc_obs_len);
            const float* f0 = a.mocap.frames + (size_t)bi.row_lo * QA_MOCAP_W;
            const float* f1 = a.mocap.frames + (size_t)bi.row_hi * QA_MOCAP_W;
            const float bl = bi.blend;
            const Quat qs = slerp_ref(Quat{f0[3], f0[4], f0[5], f0[6]}, Quat{f1[3], f1[4], f1[5], f1[6]}, bl);
            if (lane < QA_NUM_DOF) {
                dof_pos = mocap_lerp(f0[7 + lane], f1[7 + lane], bl);                  // :607
                dof_vel = mocap_lerp(f0[37 + lane], f1[37 + lane], bl);                // :608
                reinterpret_cast<float2*>(a.dof_state)[e * QA_NUM_DOF + lane] = make_float2(dof_pos, dof_vel);
            }
            const Vec3 lin = quat_rotate_sgn(
                qs, Vec3{mocap_lerp(f0[31], f1[31], bl), mocap_lerp(f0[32], f1[32], bl), mocap_lerp(f0[33], f1[33], bl)}, 1.f);
            const Vec3 ang = quat_rotate_sgn(
                qs, Vec3{mocap_lerp(f0[34], f1[34], bl), mocap_lerp(f0[35], f1[35], bl), mocap_lerp(f0[36], f1[36], bl)}, 1.f);
            root_out[0] = mocap_lerp(f0[0], f1[0], bl) + a.env_origins[e * 3 + 0];     // :668-670
            root_out[1] = mocap_lerp(f0[1], f1[1], bl) + a.env_origins[e * 3 + 1];
            root_out[2] = mocap_lerp(f0[2], f1[2], bl) + a.env_origins[e * 3 + 2];
            root_out[3] = qs.x;
            root_out[4] = qs.y;
            root_out[5] = qs.z;
            root_out[6] = qs.w;
            root_out[7] = lin.x;
            root_out[8] = lin.y;
            root_out[9] = lin.z;
            root_out[10] = ang.x;
            root_out[11] = ang.y;
            root_out[12] = ang.z;
            ep = 0;
            hist_last_action = 0.f;
            float* gah = a.action_history_buf + (size_t)e * QA_ACT_HIST_LEN * QA_NUM_DOF;
            for (int i = lane; i < QA_ACT_HIST_LEN * QA_NUM_DOF; i += 32) gah[i] = 0.f;    // :227
            if (lane < 4) a.feet_air_time[e * 4 + lane] = 0.f;                              // :224
        }
        if (is_reset || a.do_push) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 13; ++k)
                if (lane == k) v = root_out[k];
            if (lane < 13) a.root_states[(size_t)e * 13 + lane] = v;
        }

        // ---------------- stage 7: observations (:261-331) -----------------------------------------
        const float root_h = root_out[2] - center_h;
        float key_local[3] = {0.f, 0.f, 0.f};
        {
            // compute_flat_key_pos (:1377-1396): lanes 0..3 rotate one foot each
            const Quat hq = heading_quat_inv(Quat{root_out[3], root_out[4], root_out[5], root_out[6]});
            const int j = lane & 3;
            const Vec3 local = {st[S_KEY + j * 3 + 0] - root_out[0], st[S_KEY + j * 3 + 1] - root_out[1],
                                st[S_KEY + j * 3 + 2] - root_out[2]};
            const Vec3 o = quat_rotate_sgn(hq, local, 1.f);
            key_local[0] = o.x;
            key_local[1] = o.y;
            key_local[2] = o.z;
        }
        const float dq = (dof_pos - c.default_dof_pos[lane < QA_NUM_DOF ? lane : 0]) * c.s_dof_pos;
        const float dv = dof_vel * c.s_dof_vel;
        float* disc = st + S_DISC;
        if (lane < QA_NUM_DOF) {
            row[5 + lane] = dq;
            row[17 + lane] = dv;
            row[29 + lane] = hist_last_action;
            row[45 + lane] = 0.f;                                  // flat_local_key_pos * 0
            row[66 + lane] = msp - 1.f;
            row[78 + lane] = msd - 1.f;
            disc[9 + lane] = dq;
            disc[21 + lane] = dv;
        }
        if (lane < 4) {
            const float cf_ = cfilt[0] * (lane == 0) + cfilt[1] * (lane == 1) + cfilt[2] * (lane == 2) + cfilt[3] * (lane == 3);
            row[41 + lane] = cf_ - 0.5f;
            row[61 + lane] = st[S_MISC + lane];
            disc[45 + lane] = cf_ * c.s_foot_contact;
            disc[33 + lane * 3 + 0] = key_local[0] * c.s_key_pos;
            disc[33 + lane * 3 + 1] = key_local[1] * c.s_key_pos;
            disc[33 + lane * 3 + 2] = key_local[2] * c.s_key_pos;
        }
        if (lane == 0) {
            row[0] = roll;
            row[1] = pitch;
            row[2] = bav.x * c.s_ang_vel;
            row[3] = bav.y * c.s_ang_vel;
            row[4] = bav.z * c.s_ang_vel;
            row[57] = c.root_height_obs ? root_h : 0.f;
            row[58] = blv.x * c.s_lin_vel;
            row[59] = blv.y * c.s_lin_vel;
            row[60] = blv.z * c.s_lin_vel;
            row[65] = st[S_MISC + 4];
            disc[0] = roll;
            disc[1] = pitch;
            disc[2] = root_h;
            disc[3] = blv.x * c.s_lin_vel_dist;
            disc[4] = blv.y * c.s_lin_vel_dist;
            disc[5] = blv.z * c.s_lin_vel_dist;
            disc[6] = bav.x * c.s_ang_vel_dist;
            disc[7] = bav.y * c.s_ang_vel_dist;
            disc[8] = bav.z * c.s_ang_vel_dist;
        }
        if (lane < 11) row[CMD_OFF + lane] = st[S_CMD + lane];
        __syncwarp();
        // history (:302-312): fill with the current 57-vector when ep <= 1, else the shift loaded above
        if (ep <= 1) {
            for (int i = lane; i < HIST_W; i += 32) row[HIST_OFF + i] = row[i % QA_NUM_PROP];
        } else {
            for (int i = lane; i < QA_NUM_PROP; i += 32) row[HIST_OFF + HIST_W - QA_NUM_PROP + i] = row[i];
        }
        __syncwarp();

        // ---------------- stage 8: outputs ---------------------------------------------------------
        {
            float* g_obs = a.obs_buf + (size_t)e * a.obs_pitch;
            float* g_priv = a.privileged_obs_buf + (size_t)e * a.obs_pitch;
            float* g_hist = a.obs_history_buf + (size_t)e * HIST_W;
            const float clipv = c.clip_obs;
            const uint32_t slo = (uint32_t)a.rng_step, shi = (uint32_t)(a.rng_step >> 32);
            const uint32_t k0 = (uint32_t)a.rng_seed, k1 = (uint32_t)(a.rng_seed >> 32);
            for (int i = lane; i < ROW; i += 32) {
                float v = row[i];
                if (i >= HIST_OFF && i < CMD_OFF) __stcs(g_hist + (i - HIST_OFF), clampf(v, -clipv, clipv));
                const float ns = s_noise[i];
                if (ns != 0.f) {
                    float u;
                    if (a.noise_u != nullptr) {
                        u = a.noise_u[(size_t)e * ROW + i];
                    } else {
                        Philox4 r = philox4x32_10((uint32_t)e, SITE_NOISE0 + (i >> 2), slo, shi, k0, k1);
                        u = u32_to_unit_f32(r.v[i & 3]);
                    }
                    v = v + (2.f * u - 1.f) * ns;
                }
                v = clampf(v, -clipv, clipv);
                if (BULK) {
                    row[i] = v;
                } else {
                    __stcs(g_obs + i, v);
                    if (g_priv != g_obs) __stcs(g_priv + i, v);
                }
            }
            float* g_disc = a.obs_disc_buf + (size_t)e * QA_NUM_OBS_DISC;
            for (int i = lane; i < QA_NUM_OBS_DISC; i += 32) g_disc[i] = disc[i];
        }
        if (lane < QA_NUM_DOF) {                                                     // :158-161
            a.last_actions[e * QA_NUM_DOF + lane] = act;
            a.last_dof_vel[e * QA_NUM_DOF + lane] = dof_vel;
            a.last_torques_org[e * QA_NUM_DOF + lane] = tq;
        }
        if (lane < QA_NUM_REWARDS) a.episode_sums[(size_t)e * QA_EPSUM_PITCH + lane] = epsum;
        if (lane < 11) {
            const float v = st[S_CMD + lane];
            if (lane < 5) a.commands[e * 5 + lane] = v;
            else if (lane == 5) a.latent_eps[e] = v;
            else a.latent_c[e * QA_DIM_C + lane - 6] = v;
        }
        {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (lane == k) v = root_out[7 + k];
            if (lane < 6) a.last_root_vel[e * 6 + lane] = v;
        }
        if (lane < 4) {
            const float ff = feet_force[0] * (lane == 0) + feet_force[1] * (lane == 1) + feet_force[2] * (lane == 2) +
                             feet_force[3] * (lane == 3);
            const bool ct = (lane == 0 && contact[0]) || (lane == 1 && contact[1]) || (lane == 2 && contact[2]) ||
                            (lane == 3 && contact[3]);
            const bool cfl = (lane == 0 && cfilt[0]) || (lane == 1 && cfilt[1]) || (lane == 2 && cfilt[2]) ||
                             (lane == 3 && cfilt[3]);
            a.feet_forces[e * 4 + lane] = ff;
            a.last_contacts[e * 4 + lane] = ct ? 1 : 0;
            a.contact_filt[e * 4 + lane] = cfl ? 1 : 0;
            if (a.contact_buf)
                a.contact_buf[((size_t)e * a.contact_ring_len + a.contact_ring_head) * 4 + lane] = cfl ? 1.f : 0.f;
            if (a.contact_force_buf)
                a.contact_force_buf[((size_t)e * a.contact_ring_len + a.contact_ring_head) * 4 + lane] =
                    clampf(ff, -c.clip_obs, c.clip_obs);
        }
        if (lane < 3) {
            a.base_lin_vel[e * 3 + lane] = lane == 0 ? blv.x : (lane == 1 ? blv.y : blv.z);
            a.base_ang_vel[e * 3 + lane] = lane == 0 ? bav.x : (lane == 1 ? bav.y : bav.z);
            a.projected_gravity[e * 3 + lane] = lane == 0 ? pg.x : (lane == 1 ? pg.y : pg.z);
            a.rpy[e * 3 + lane] = lane == 0 ? roll : (lane == 1 ? pitch : yaw);
        }
        if (lane == 0) {
            a.rew_buf[e] = rew;
            a.reset_buf[e] = is_reset ? 1 : 0;
            a.time_out_buf[e] = time_out ? 1 : 0;
            a.episode_length_buf[e] = ep;
            a.root_h[e] = root_h_pre;
        }
    }

    if (BULK) {
        // all rows of this CTA are final in smem: hand the tile to the TMA engine
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const int first = blockIdx.x * K2_ENVS;
            const int rows = min(K2_ENVS, a.num_envs - first);
            const unsigned bytes = (unsigned)(rows * ROW * sizeof(float));
            bulk_store_tile(a.obs_buf + (size_t)first * ROW, s_tile, bytes);
            if (a.privileged_obs_buf != a.obs_buf) bulk_store_tile(a.privileged_obs_buf + (size_t)first * ROW, s_tile, bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }

    // ---------------- epilogue: reset statistics, last CTA finalises ----------------------------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        K2Workspace* ws = reinterpret_cast<K2Workspace*>(a.workspace);
        const unsigned t = atomicAdd(&ws->ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        K2Workspace* ws = reinterpret_cast<K2Workspace*>(a.workspace);
        const unsigned cnt = *reinterpret_cast<volatile unsigned*>(&ws->reset_count);
        if (cnt > 0) {
            if (threadIdx.x < QA_NUM_REWARDS) {
                const double s = *reinterpret_cast<volatile double*>(&ws->sums[threadIdx.x]);
                const float mean = (float)(s / (double)cnt);
                a.episode_rew_means[threadIdx.x] = mean / c.episode_length_s;
            }
            // extras["time_outs"] = time_out_buf, only on steps with >= 1 reset (:239-240)
            const volatile uint8_t* src = a.time_out_buf;
            for (int i = threadIdx.x; i < a.num_envs; i += K2_THREADS) a.time_outs_latched[i] = src[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            *a.num_resets = (int)cnt;
            if (a.step_state != nullptr) a.step_state[0] = (long long)a.rng_step;
            ws->reset_count = 0u;
            ws->ticket = 0u;
        }
        if (threadIdx.x < QA_NUM_REWARDS) ws->sums[threadIdx.x] = 0.0;
    }
    if (BULK) {
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

int qa_k2_try_launch_tiled(const QaBbcConst* c, const QaBbcStepArgs* a, cudaStream_t stream, int* launched);

extern "C" int qa_post_physics_bbc(const QaBbcConst* c, const QaBbcStepArgs* a, void* stream) {
    QA_CHECK_PTR(c);
    QA_CHECK_PTR(a);
    if (a->num_envs < 0) return QA_EINVAL;
    if (c->num_bodies <= 0 || c->num_bodies > QA_MAX_BODIES) return QA_ERANGE;
    for (int j = 0; j < 4; ++j)
        if (c->feet_indices[j] < 0 || c->feet_indices[j] >= c->num_bodies) return QA_ERANGE;
    if (a->obs_pitch < QA_OBS_WIDTH) return QA_EINVAL;
    if (c->resample_period <= 0) return QA_EINVAL;
    QA_CHECK_PTR(a->root_states);
    QA_CHECK_PTR(a->dof_state);
    QA_CHECK_PTR(a->rigid_body_state);
    QA_CHECK_PTR(a->contact_forces);
    QA_CHECK_PTR(a->motor_strength);
    QA_CHECK_PTR(a->mass_params);
    QA_CHECK_PTR(a->friction_coeffs);
    QA_CHECK_PTR(a->env_origins);
    QA_CHECK_PTR(a->noise_scale_vec);
    QA_CHECK_PTR(a->mocap.frames);
    QA_CHECK_PTR(a->mocap.clip_start);
    QA_CHECK_PTR(a->mocap.clip_nframes);
    QA_CHECK_PTR(a->mocap.clip_len_s);
    QA_CHECK_PTR(a->mocap.clip_frame_dur);
    QA_CHECK_PTR(a->mocap.mode_offset);
    QA_CHECK_PTR(a->mocap.mode_clips);
    QA_CHECK_PTR(a->mocap.mode_cdf);
    if (c->measure_heights) {
        QA_CHECK_PTR(a->terrain.height_samples);
        if (a->terrain.rows < 2 || a->terrain.cols < 2) return QA_EINVAL;
    }
    QA_CHECK_PTR(a->episode_length_buf);
    QA_CHECK_PTR(a->last_contacts);
    QA_CHECK_PTR(a->commands);
    QA_CHECK_PTR(a->latent_eps);
    QA_CHECK_PTR(a->latent_c);
    QA_CHECK_PTR(a->actions);
    QA_CHECK_PTR(a->last_actions);
    QA_CHECK_PTR(a->torques_org);
    QA_CHECK_PTR(a->last_torques_org);
    QA_CHECK_PTR(a->last_dof_vel);
    QA_CHECK_PTR(a->last_root_vel);
    QA_CHECK_PTR(a->action_history_buf);
    QA_CHECK_PTR(a->obs_history_buf);
    QA_CHECK_PTR(a->episode_sums);
    QA_CHECK_PTR(a->feet_air_time);
    QA_CHECK_PTR(a->obs_buf);
    QA_CHECK_PTR(a->privileged_obs_buf);
    QA_CHECK_PTR(a->obs_disc_buf);
    QA_CHECK_PTR(a->rew_buf);
    QA_CHECK_PTR(a->reset_buf);
    QA_CHECK_PTR(a->time_out_buf);
    QA_CHECK_PTR(a->base_lin_vel);
    QA_CHECK_PTR(a->base_ang_vel);
    QA_CHECK_PTR(a->projected_gravity);
    QA_CHECK_PTR(a->rpy);
    QA_CHECK_PTR(a->feet_forces);
    QA_CHECK_PTR(a->contact_filt);
    QA_CHECK_PTR(a->root_h);
    QA_CHECK_PTR(a->episode_rew_means);
    QA_CHECK_PTR(a->time_outs_latched);
    QA_CHECK_PTR(a->num_resets);
    QA_CHECK_PTR(a->workspace);
    // parity-mode draws come as a complete set or not at all
    const bool any_draw = a->noise_u || a->rs_eps_u || a->rs_c_idx || a->rs_cmd_u || a->rt_eps_u || a->rt_c_idx ||
                          a->rt_cmd_u || a->push_u || a->mocap_clip_idx || a->mocap_time_u;
    const bool all_draw = a->noise_u && a->rs_eps_u && a->rs_c_idx && a->rs_cmd_u && a->rt_eps_u && a->rt_c_idx &&
                          a->rt_cmd_u && a->push_u && a->mocap_clip_idx && a->mocap_time_u;
    if (any_draw && !all_draw) return QA_EINVAL;
    if (a->contact_ring_len > 0 && (a->contact_ring_head < 0 || a->contact_ring_head >= a->contact_ring_len))
        return QA_EINVAL;
    if (a->num_envs == 0) return 0;

    {
        int launched = 0;
        const int rc = qa_k2_try_launch_tiled(c, a, (cudaStream_t)stream, &launched);
        if (rc != 0 || launched) return rc;
    }
    const int grid = (a->num_envs + K2_ENVS - 1) / K2_ENVS;
    const bool aligned = (((uintptr_t)a->obs_buf | (uintptr_t)a->privileged_obs_buf) & 15u) == 0;
    const bool bulk = (a->flags & QA_K2_BULK_STORE) && a->obs_pitch == QA_OBS_WIDTH && aligned &&
                      (a->num_envs % K2_ENVS == 0);
    if (bulk)
        k_post_physics_bbc<true><<<grid, K2_THREADS, 0, (cudaStream_t)stream>>>(*c, *a);
    else
        k_post_physics_bbc<false><<<grid, K2_THREADS, 0, (cudaStream_t)stream>>>(*c, *a);
    QA_LAUNCH_RET();
}

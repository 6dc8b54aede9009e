// Small per-step env kernels of the BBC LeggedRobot path (sm_100a):
//   K0 qa_action_push   legged_robot.py:84-98
//   K1 qa_pd_torques    legged_robot.py:547-579
//   K3 qa_height_scan   legged_robot.py:1190-1228
//   K4 qa_mocap_blend   motion_loader.py:410-447 + utils.py:126-159
//      qa_compact_resets legged_robot.py:153-154 (reset_buf.nonzero() + obs_disc_buf[env_ids])
// All are HBM/latency bound element-wise kernels: one thread per (env, dof) or per sample point,
// coalesced along the fastest dimension, no shared-memory reuse to exploit.
#include "qa_b200.h"
#include "qa_common.cuh"
#include "qa_mocap.cuh"

// ------------------------------------------------------------------------------------------
// K0: action history push + delayed select + clip.  One thread per (env, dof); the 8-slot
// history of that dof is shifted through registers.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_action_push(QaActionPushArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.num_envs * QA_NUM_DOF) return;
    const int e = i / QA_NUM_DOF, d = i - e * QA_NUM_DOF;
    float* h = a.action_history_buf + (size_t)e * QA_ACT_HIST_LEN * QA_NUM_DOF + d;
    float v[QA_ACT_HIST_LEN];
#pragma unroll
    for (int s = 0; s < QA_ACT_HIST_LEN - 1; ++s) v[s] = h[(s + 1) * QA_NUM_DOF];
    v[QA_ACT_HIST_LEN - 1] = a.actions_in[i];
    float sel = v[QA_ACT_HIST_LEN - 1];
#pragma unroll
    for (int s = 0; s < QA_ACT_HIST_LEN; ++s) {
        h[s * QA_NUM_DOF] = v[s];
        if (s == QA_ACT_HIST_LEN - 1 - a.delay) sel = v[s];
    }
    a.actions_out[i] = clampf(sel, -a.clip, a.clip);
}

extern "C" int qa_action_push(const QaActionPushArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->num_envs == 0) return 0;          // empty input: nothing to do (pointers may be null)
    QA_CHECK_PTR(a->actions_in);
    QA_CHECK_PTR(a->action_history_buf);
    QA_CHECK_PTR(a->actions_out);
    if (a->num_envs < 0) return QA_EINVAL;
    if (a->delay < 0 || a->delay >= QA_ACT_HIST_LEN) return QA_ERANGE;
    if (a->num_envs == 0) return 0;
    const int n = a->num_envs * QA_NUM_DOF;
    k_action_push<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K1: PD torques.  One thread per (env, dof); dof_state read as float2 (pos, vel).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pd_torques(QaTorqueArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.num_envs * QA_NUM_DOF;
    if (i >= n) return;
    const int d = i % QA_NUM_DOF;
    const float2 pv = reinterpret_cast<const float2*>(a.dof_state)[i];
    float as = a.actions[i] * a.action_scale;
    if (d % 3 == 0) as *= a.hip_scale_reduction;          // DOF 0,3,6,9
    const float msp = a.motor_strength[i], msd = a.motor_strength[n + i];
    const float t = msp * a.p_gains[d] * (as + a.default_dof_pos[d] - pv.x) - msd * a.d_gains[d] * pv.y;
    a.torques_org[i] = t;
    const float lim = a.torque_limits[d];
    a.torques[i] = fminf(fmaxf(t, -lim), lim);
}

extern "C" int qa_pd_torques(const QaTorqueArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->num_envs == 0) return 0;
    QA_CHECK_PTR(a->actions);
    QA_CHECK_PTR(a->dof_state);
    QA_CHECK_PTR(a->motor_strength);
    QA_CHECK_PTR(a->p_gains);
    QA_CHECK_PTR(a->d_gains);
    QA_CHECK_PTR(a->default_dof_pos);
    QA_CHECK_PTR(a->torque_limits);
    QA_CHECK_PTR(a->torques);
    QA_CHECK_PTR(a->torques_org);
    if (a->num_envs < 0) return QA_EINVAL;
    if (a->num_envs == 0) return 0;
    const int n = a->num_envs * QA_NUM_DOF;
    k_pd_torques<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K3: terrain height scan, one thread per (env, point).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float terrain_height_at(const QaTerrain& t, Quat yawq, float bx, float by, float bz,
                                                    float hx, float hy, float hz) {
    Vec3 p = quat_apply(yawq, Vec3{hx, hy, hz});
    float wx = p.x + bx, wy = p.y + by;
    (void)bz;
    wx = wx + t.border_size;
    wy = wy + t.border_size;
    long long ix = (long long)(wx / t.horizontal_scale);      // .long(): truncation toward zero
    long long iy = (long long)(wy / t.horizontal_scale);
    ix = ix < 0 ? 0 : (ix > t.rows - 2 ? t.rows - 2 : ix);
    iy = iy < 0 ? 0 : (iy > t.cols - 2 ? t.cols - 2 : iy);
    const int16_t* hs = t.height_samples;
    const int16_t h1 = __ldg(hs + ix * t.cols + iy);
    const int16_t h2 = __ldg(hs + (ix + 1) * t.cols + iy);
    const int16_t h3 = __ldg(hs + ix * t.cols + iy + 1);
    int16_t h = h1 < h2 ? h1 : h2;
    h = h < h3 ? h : h3;
    return (float)h * t.vertical_scale;
}

__global__ void __launch_bounds__(256) k_height_scan(QaHeightScanArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.num_envs * a.num_points) return;
    const int e = i / a.num_points, p = i - e * a.num_points;
    const float* r = a.root_states + (size_t)e * 13;
    Quat q = {r[3], r[4], r[5], r[6]};
    Quat yq = yaw_quat(q);
    a.measured_heights[i] = terrain_height_at(a.terrain, yq, r[0], r[1], r[2], a.height_points[p * 3 + 0],
                                              a.height_points[p * 3 + 1], a.height_points[p * 3 + 2]);
}

extern "C" int qa_height_scan(const QaHeightScanArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->num_envs == 0) return 0;
    QA_CHECK_PTR(a->root_states);
    QA_CHECK_PTR(a->height_points);
    QA_CHECK_PTR(a->terrain.height_samples);
    QA_CHECK_PTR(a->measured_heights);
    if (a->num_envs < 0 || a->num_points <= 0 || a->terrain.rows < 2 || a->terrain.cols < 2) return QA_EINVAL;
    if (a->num_envs == 0) return 0;
    const long long n = (long long)a->num_envs * a->num_points;
    k_height_scan<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// K4: mocap frame blend, one warp per output row (49 floats -> lanes 0..31 take col, col+32).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mocap_blend(QaMocapBlendArgs a) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= a.num) return;
    int clip = a.clip_idx[w];
    clip = min(max(clip, 0), a.table.num_clips - 1);
    const MocapBlendIdx bi = mocap_blend_index(a.table, clip, a.time_u[w], a.time_between_frames, a.disc_obs_len);
    const float* f0 = a.table.frames + (size_t)bi.row_lo * QA_MOCAP_W;
    const float* f1 = a.table.frames + (size_t)bi.row_hi * QA_MOCAP_W;
    Quat q0 = {f0[3], f0[4], f0[5], f0[6]}, q1 = {f1[3], f1[4], f1[5], f1[6]};
    const Quat qs = slerp_ref(q0, q1, bi.blend);
    float* out = a.frames_out + (size_t)w * QA_MOCAP_W;
    for (int c = lane; c < QA_MOCAP_W; c += 32) {
        float v;
        if (c >= 3 && c < 7)
            v = c == 3 ? qs.x : (c == 4 ? qs.y : (c == 5 ? qs.z : qs.w));
        else
            v = mocap_lerp(f0[c], f1[c], bi.blend);
        out[c] = v;
    }
}

extern "C" int qa_mocap_blend(const QaMocapBlendArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->num == 0) return 0;
    QA_CHECK_PTR(a->table.frames);
    QA_CHECK_PTR(a->table.clip_start);
    QA_CHECK_PTR(a->table.clip_nframes);
    QA_CHECK_PTR(a->table.clip_len_s);
    QA_CHECK_PTR(a->table.clip_frame_dur);
    QA_CHECK_PTR(a->clip_idx);
    QA_CHECK_PTR(a->time_u);
    QA_CHECK_PTR(a->frames_out);
    if (a->num < 0 || a->table.num_clips <= 0) return QA_EINVAL;
    if (a->num == 0) return 0;
    const int warps_per_block = 4;
    k_mocap_blend<<<(a->num + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                    (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// Reset-mask compaction: ascending env ids (== reset_buf.nonzero()) + gather of the terminal
// discriminator states.  Single CTA: N <= a few 10^4 bytes of mask; ballot + warp prefix, then a
// block prefix over the warp counts, repeated over chunks of 1024 envs (order preserving).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_compact_resets(QaCompactArgs a) {
    __shared__ int warp_cnt[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int start = 0; start < a.num_envs; start += 1024) {
        const int e = start + tid;
        const bool flag = e < a.num_envs && a.reset_buf[e] != 0;
        const unsigned bal = __ballot_sync(QA_FULL, flag);
        const int within = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int wbase = 0, total = 0;
        for (int k = 0; k < 32; ++k) {
            const int c = warp_cnt[k];
            if (k < wid) wbase += c;
            total += c;
        }
        const int base = base_s;
        if (flag) {
            const int pos = base + wbase + within;
            a.reset_env_ids[pos] = e;
            if (a.reset_env_ids_i32) a.reset_env_ids_i32[pos] = e;
        }
        __syncthreads();
        if (tid == 0) base_s = base + total;
        __syncthreads();
    }
    const int count = base_s;
    if (tid == 0) *a.count = count;
    if (a.terminal_disc_states && a.prev_obs_disc_buf) {
        __syncthreads();   // ids written by this CTA are visible to it after the barrier
        for (int i = tid; i < count * QA_NUM_OBS_DISC; i += blockDim.x) {
            const int r = i / QA_NUM_OBS_DISC, c = i - r * QA_NUM_OBS_DISC;
            a.terminal_disc_states[i] = a.prev_obs_disc_buf[(size_t)a.reset_env_ids[r] * QA_NUM_OBS_DISC + c];
        }
    }
}

extern "C" int qa_compact_resets(const QaCompactArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    QA_CHECK_PTR(a->reset_buf);
    QA_CHECK_PTR(a->reset_env_ids);
    QA_CHECK_PTR(a->count);
    if (a->num_envs < 0) return QA_EINVAL;
    k_compact_resets<<<1, 1024, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

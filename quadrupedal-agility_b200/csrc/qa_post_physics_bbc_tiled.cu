// K2 "tiled" variant: the fused BBC post-physics step with 8-env CTA tiles staged entirely by TMA.
//
// Same function as k_post_physics_bbc (qa_post_physics_bbc.cu; reference lines cited there), restructured after
// the round-1 profile showed the warp-per-env kernel to be instruction/latency bound (3150 warp-instructions per
// env, 32x redundant scalar math, serialised global->shared copies), not HBM bound:
//
//   P0  every load of the tile is put in flight before the first barrier: the two scalar warps fetch what their
//       thread-per-env programs need (root tile, feet positions, commands, counters; ~1.5 KB) with plain 16-B loads,
//       8 threads of 8 different warps issue 13 TMA bulk loads (cp.async.bulk + mbarrier complete_tx; every per-env
//       array's 8-env slice is contiguous and 16 B aligned) -- the 12 small tiles on one mbarrier, the 18 KB history
//       tile on its own -- and the env warps draw this step's Philox noise while they wait;
//   P2a (scalar warp A, lane = (env, role), starts as soon as ITS loads land): four roles per env -- base linear velocity |
//       base angular velocity | projected gravity + euler angles | centre terrain height, periodic resampling, push -- so the
//       dependent chain is the longest role, not their sum;  (scalar warp B, lane = (env, foot)): key-body positions;
//   P1  after the small tiles: scalar warp B (lane = (env, leg), 3 DOFs per lane) forms the nine 12-wide reward sums,
//       the env warps (lanes = bodies) the contact-force norms (ballots) -> per-env scalars; named barrier 1 hands them to
//   P2b (scalar warp A): contacts, termination, reward total, episode sums, reset decision -- while the env warps run
//   P3a: the row lanes that depend on loaded inputs only and, once the history tile has landed, the history shift;
//   P3b (env warps, after barrier 2): reset envs blend their mocap frame (lanes = frame columns) and rewrite simulator
//       state; the P2-dependent row lanes, newest history slot, noise on the <= 64 noisy lanes, clip, disc obs;
//       meanwhile scalar warp A writes the ~14 small per-env outputs with plain 16-B stores;
//   P4  four threads of four warps issue the TMA bulk stores of the big tiles (obs, privileged obs, history, disc obs;
//       root / dof state only after a reset or push); the CTA that takes the last ticket finalises the reset statistics,
//       latches extras["time_outs"] and advances the device step counter.
//
// Per env this is ~1000 warp-instructions, and all bulk global traffic is full-line copies.
// profiles/r1_k2_phase_trace.txt has the per-phase clock trace that led here (tools/k2_trace.py).
#include "qa_k2_common.cuh"

#define T2_ENVS 8
#define T2_THREADS (T2_ENVS * 32 + 64)   // 8 env warps + 2 scalar warps
#define W_SA T2_ENVS                     // scalar warp A: base-frame quantities, terrain, rewards, reset decision
#define W_SB (T2_ENVS + 1)               // scalar warp B: key-body positions

// per-env scalar slots (floats) exchanged between the phases
enum {
    SC_SUM0 = 0,        // 9 DOF sums: action_rate, delta_torques, dof_acc, dof_error, dof_pos_limits, dof_vel_limits,
                        //             hip_pos, torque_limits, torques
    SC_NCOL = 9,
    SC_TERM = 10,
    SC_FF = 11,         // 4 feet force norms
    SC_RESET = 15,
    SC_CH = 16,         // centre terrain height
    SC_MODE = 17,       // behaviour mode drawn at reset (int bits)
    SC_KEY = 18,        // 4 x 3 key-body (feet) positions in the heading frame
    SC_N = 32
};

// slots of T2Smem::rst
enum { RST_MP = 0, RST_MV = 12, RST_Q = 24, RST_LIN = 28, RST_ANG = 31, RST_POS = 34, RST_KEY = 37, RST_N = 52 };

struct __align__(16) T2Smem {
    float obs[T2_ENVS * ROW];
    float hist[T2_ENVS * HIST_W];
    float root[T2_ENVS * 13];
    float dof[T2_ENVS * 24];
    float cf[T2_ENVS * QA_MAX_BODIES * 3];
    float act[T2_ENVS * 12], lact[T2_ENVS * 12], tq[T2_ENVS * 12], ltq[T2_ENVS * 12], ldv[T2_ENVS * 12];
    float msp[T2_ENVS * 12], msd[T2_ENVS * 12];
    float cmd[T2_ENVS * 5], eps[T2_ENVS], lc[T2_ENVS * 5];
    float mass[T2_ENVS * 4], fric[T2_ENVS];
    float epsum[T2_ENVS * QA_EPSUM_PITCH];
    long long ep[T2_ENVS];
    uint8_t lcont[T2_ENVS * 4];
    float key[T2_ENVS * 12];
    float disc[T2_ENVS * QA_NUM_OBS_DISC];
    float rew[T2_ENVS], rooth[T2_ENVS];
    float blv[T2_ENVS * 3], bav[T2_ENVS * 3], pg[T2_ENVS * 3], rpy[T2_ENVS * 3];
    float ff[T2_ENVS * 4];
    uint8_t cont_out[T2_ENVS * 4], cfilt_out[T2_ENVS * 4];
    float scal[T2_ENVS][SC_N];
    float rootB[T2_ENVS * 13];           // scalar warp B's private copy of the root tile
    float rst[T2_ENVS][RST_N];           // a resetting env's precomputed mocap frame (reset_precompute)
    uint64_t bar, bar_small;
    unsigned last;
};

// Optional phase trace (tools/k2_trace.py builds a -DQA_K2_TRACE variant of the library; never in the product build):
// per-CTA clock64 stamps at the phase boundaries, slots 0..15, plus %globaltimer at entry / exit in slots 16, 17.
#ifdef QA_K2_TRACE
#define K2_TRACE_CTAS 1024
__device__ long long g_k2_trace[K2_TRACE_CTAS][24];
__device__ __forceinline__ long long k2_gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define STAMP(slot, who)                                                                   \
    do {                                                                                   \
        if (threadIdx.x == (who) && blockIdx.x < K2_TRACE_CTAS) g_k2_trace[blockIdx.x][slot] = clock64(); \
    } while (0)
#define GSTAMP(slot, who)                                                                  \
    do {                                                                                   \
        if (threadIdx.x == (who) && blockIdx.x < K2_TRACE_CTAS) g_k2_trace[blockIdx.x][slot] = k2_gtime(); \
    } while (0)
// attribution experiments (trace build only): bit 0 = the big output tiles (obs, privileged obs, history, disc obs) are NOT stored
__device__ int g_k2_dbg_mode = 0;
extern "C" int qa_k2_trace_set_mode(int mode) { return (int)cudaMemcpyToSymbol(g_k2_dbg_mode, &mode, sizeof(int)); }
#define K2_DBG_NO_TILE_STORES (g_k2_dbg_mode & 1)
extern "C" int qa_k2_trace_dump(long long* host, int max_ctas) {
    const int n = max_ctas < K2_TRACE_CTAS ? max_ctas : K2_TRACE_CTAS;
    return (int)cudaMemcpyFromSymbol(host, g_k2_trace, sizeof(long long) * 24 * n);
}
#else
#define STAMP(slot, who) do { } while (0)
#define GSTAMP(slot, who) do { } while (0)
#define K2_DBG_NO_TILE_STORES 0
#endif
#define T_ENV 0                          // first env warp, lane 0
#define T_SA (W_SA * 32)                 // scalar warp A, lane 0
#define T_SB (W_SB * 32)                 // scalar warp B, lane 0

// thread-level resampler (same arithmetic as resample_env)
__device__ __forceinline__ void resample_thread(const QaBbcConst& c, const K2Draw& d, float* cmd, float* eps, float* lc) {
    const int m = d.c_idx;
    float n0 = (c.lin_vel_x[m][1] - c.lin_vel_x[m][0]) * d.cmd_u[0] + c.lin_vel_x[m][0];
    float n1 = (c.lin_vel_y[m][1] - c.lin_vel_y[m][0]) * d.cmd_u[1] + c.lin_vel_y[m][0];
    float n2 = (c.ang_vel_yaw[m][1] - c.ang_vel_yaw[m][0]) * d.cmd_u[2] + c.ang_vel_yaw[m][0];
    const float jump = (m == QA_DIM_C - 1) ? 1.f : 0.f;
    cmd[3] = (c.jump_h_span * d.cmd_u[3] + c.jump_h_lo) * jump;
    cmd[4] = (c.loco_h_span * d.cmd_u[4] + c.loco_h_lo) * (1.f - jump);
    n0 *= (fabsf(n0) > c.lin_vel_x_clip) ? 1.f : 0.f;
    n1 *= (fabsf(n1) > c.lin_vel_y_clip) ? 1.f : 0.f;
    n2 *= (fabsf(n2) > c.ang_vel_yaw_clip) ? 1.f : 0.f;
    cmd[0] = n0;
    cmd[1] = n1;
    cmd[2] = n2;
    eps[0] = (float)(d.eps_u * 2. - 1.);
#pragma unroll
    for (int k = 0; k < QA_DIM_C; ++k) lc[k] = (k == m) ? 1.f : 0.f;
}

// sum over the low 16 lanes (lanes 12..15 hold zeros): the 16-offset step of warp_sum would only add zeros
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(QA_FULL, v, o);
    return v;
}

// 64-bit counters that in practice fit 32 bits: take the short hardware-divide path (the 64-bit software modulo was
// ~10 % of the kernel's instructions)
__device__ __forceinline__ bool divisible_by(long long v, int period) {
    const unsigned long long u = (unsigned long long)v;
    if ((u >> 32) == 0ull) return ((unsigned)u % (unsigned)period) == 0u;
    return (v % (long long)period) == 0;
}

__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void copy16(void* gdst, const void* ssrc, int n16, int lane) {
    for (int i = lane; i < n16; i += 32) reinterpret_cast<float4*>(gdst)[i] = reinterpret_cast<const float4*>(ssrc)[i];
}

// Everything of reset_idx (:178-240, :598-612) that does NOT depend on this step's scalar program: the mocap clip / time draw,
// the blended frame (slerp of the root quaternion, lerps of root position / velocities / DOF state) and the key-body positions
// on the new root.  The draws are functions of (env, site, step) only and the reset DECISION needs just the contact ballot,
// the episode counter and the root height, so the env warp runs this ~4 us chain of dependent global loads BEFORE barrier 2,
// next to the scalar warps' P2a / P2b, instead of after it (profiles/r2_k2_phase_trace_reset_path.txt: the 8.6 us reset path
// behind barrier 2 set the kernel's tail -- ~11 % of the CTAs hold a resetting env).  Results go to `rst`; P3b commits them.
// Same arithmetic, op for op, as the in-place version it replaces.
__device__ __forceinline__ void reset_precompute(const QaBbcConst& c, const K2Step& a, int e, int lane, int B, float* rst,
                                                 int mode_off) {
    // `mode_off`: lane 18 + k holds mocap.mode_offset[k] (loaded at kernel entry; production draws only)
    int clip;
    double time_u;
    // clip metadata of the chosen clip (mocap_blend_index's operands)
    double len_s, frame_dur, nframes;
    int start;
    bool have_meta = false;
    if (a.mocap_clip_idx != nullptr) {
        clip = a.mocap_clip_idx[e];
        time_u = a.mocap_time_u[e];
    } else {
        Philox4 rr = philox4x32_10((uint32_t)e, SITE_MOCAP, (uint32_t)a.rng_step, (uint32_t)(a.rng_step >> 32),
                                   (uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
        const double cu = u64_to_unit_f64(rr.v[0], rr.v[1]);
        time_u = u64_to_unit_f64(rr.v[2], rr.v[3]);
        // the behaviour mode P2b is going to draw for this reset (same site, same stream)
        int m;                                                          // == draw_site(..., SITE_RT0, ...).c_idx
        if (a.rt_eps_u != nullptr) {
            m = a.rt_c_idx[e];
        } else {
            Philox4 r0 = philox4x32_10((uint32_t)e, SITE_RT0, (uint32_t)a.rng_step, (uint32_t)(a.rng_step >> 32),
                                       (uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
            m = pick_mode(c, a.prior_cdf, u32_to_unit_f32(r0.v[2]));
        }
        m = min(max(m, 0), QA_DIM_C - 1);
        const int lo = __shfl_sync(QA_FULL, mode_off, 18 + m), hi = __shfl_sync(QA_FULL, mode_off, 18 + m + 1);
        // The sequential scan `j = lo; while (j < hi - 1 && cdf[j] <= cu) ++j` ends at the first j in [lo, hi - 1) with
        // !(cdf[j] <= cu), else at hi - 1.  32 candidates per round, one per lane; every lane also fetches ITS candidate's clip
        // id and metadata, so the winner's operands arrive with two dependent loads instead of four.
        clip = -1;
        for (int base = lo; base <= hi - 1; base += 32) {
            const int jj = base + lane;
            const bool in = jj <= hi - 1;
            const double cj = (in && jj < hi - 1) ? a.mocap.mode_cdf[jj] : 0.0;
            const int cl = in ? min(max(a.mocap.mode_clips[jj], 0), a.mocap.num_clips - 1) : 0;
            const double c_len = a.mocap.clip_len_s[cl], c_dur = a.mocap.clip_frame_dur[cl], c_n = a.mocap.clip_nframes[cl];
            const int c_start = a.mocap.clip_start[cl];
            const bool stop = in && (jj == hi - 1 || !(cj <= cu));
            const unsigned hit = __ballot_sync(QA_FULL, stop);
            if (hit != 0u) {
                const int w = __ffs((int)hit) - 1;
                clip = __shfl_sync(QA_FULL, cl, w);
                len_s = __shfl_sync(QA_FULL, c_len, w), frame_dur = __shfl_sync(QA_FULL, c_dur, w);
                nframes = __shfl_sync(QA_FULL, c_n, w), start = __shfl_sync(QA_FULL, c_start, w);
                have_meta = true;
                break;
            }
        }
        if (clip < 0) clip = a.mocap.mode_clips[max(lo, hi - 1)];             // empty mode: what the scan would read
    }
    clip = min(max(clip, 0), a.mocap.num_clips - 1);
    STAMP(21, T_ENV);
    if (!have_meta) {
        len_s = a.mocap.clip_len_s[clip], frame_dur = a.mocap.clip_frame_dur[clip], nframes = a.mocap.clip_nframes[clip];
        start = a.mocap.clip_start[clip];
    }
    const MocapBlendIdx bi = mocap_blend_index_meta(len_s, frame_dur, nframes, start, time_u, c.time_between_frames, c.disc_obs_len);
    const float* f0 = a.mocap.frames + (size_t)bi.row_lo * QA_MOCAP_W;
    const float* f1 = a.mocap.frames + (size_t)bi.row_hi * QA_MOCAP_W;
    const float bl = bi.blend;
    // every global operand of the frame first (independent loads, one latency), then the arithmetic
    const bool dl = lane < QA_NUM_DOF;
    const int d_ = dl ? lane : 0;
    const float p0 = f0[7 + d_], p1 = f1[7 + d_], w0 = f0[37 + d_], w1 = f1[37 + d_];
    float h0[7], h1[7], u0[6], u1[6];
#pragma unroll
    for (int k = 0; k < 7; ++k) h0[k] = f0[k], h1[k] = f1[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) u0[k] = f0[31 + k], u1[k] = f1[31 + k];
    const float ox = a.env_origins[e * 3 + 0], oy = a.env_origins[e * 3 + 1], oz = a.env_origins[e * 3 + 2];
    float kx = 0.f, ky = 0.f, kz = 0.f;
    if (lane < 4) {
        const float* kp = a.rigid_body_state + ((size_t)e * B + c.feet_indices[lane]) * 13;
        kx = kp[0], ky = kp[1], kz = kp[2];
    }
    STAMP(22, T_ENV);
    const Quat qs = slerp_ref(Quat{h0[3], h0[4], h0[5], h0[6]}, Quat{h1[3], h1[4], h1[5], h1[6]}, bl);
    STAMP(23, T_ENV);
    if (dl) {
        rst[RST_MP + lane] = mocap_lerp(p0, p1, bl);
        rst[RST_MV + lane] = mocap_lerp(w0, w1, bl);
    }
    // root linear (lane 0) and angular (lane 1) velocity: the same rotation on two lanes
    const int vo = lane == 1 ? 3 : 0;
    const Vec3 vel = quat_rotate_sgn(
        qs, Vec3{mocap_lerp(vo ? u0[3] : u0[0], vo ? u1[3] : u1[0], bl), mocap_lerp(vo ? u0[4] : u0[1], vo ? u1[4] : u1[1], bl),
                 mocap_lerp(vo ? u0[5] : u0[2], vo ? u1[5] : u1[2], bl)}, 1.f);
    const float px = mocap_lerp(h0[0], h1[0], bl) + ox, py = mocap_lerp(h0[1], h1[1], bl) + oy, pz = mocap_lerp(h0[2], h1[2], bl) + oz;
    if (lane == 0) {
        rst[RST_Q + 0] = qs.x, rst[RST_Q + 1] = qs.y, rst[RST_Q + 2] = qs.z, rst[RST_Q + 3] = qs.w;
        rst[RST_POS + 0] = px, rst[RST_POS + 1] = py, rst[RST_POS + 2] = pz;
    }
    if (lane < 2) rst[RST_LIN + vo + 0] = vel.x, rst[RST_LIN + vo + 1] = vel.y, rst[RST_LIN + vo + 2] = vel.z;   // RST_ANG = RST_LIN + 3
    if (lane < 4) {                                                         // key-body positions on the mocap root
        const Quat hq = heading_quat_inv(qs);
        const Vec3 o = quat_rotate_sgn(hq, Vec3{kx - px, ky - py, kz - pz}, 1.f);
        rst[RST_KEY + lane * 3 + 0] = o.x * c.s_key_pos, rst[RST_KEY + lane * 3 + 1] = o.y * c.s_key_pos;
        rst[RST_KEY + lane * 3 + 2] = o.z * c.s_key_pos;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(T2_THREADS, 4)
k_post_physics_bbc_tiled(const __grid_constant__ QaBbcConst c, const __grid_constant__ QaBbcStepArgs a_in) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T2Smem& S = *reinterpret_cast<T2Smem*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int e0 = blockIdx.x * T2_ENVS;
    const int B = c.num_bodies;
    const QaBbcStepArgs& r = a_in;
    GSTAMP(16, T_ENV);
    STAMP(0, T_ENV);
    // Programmatic dependent launch (QA_K2_PDL): this grid may have been scheduled while its predecessor in the stream was
    // still draining -- nothing global is read before the predecessor has completed and flushed (`wait`); our own dependents
    // may be scheduled right away (`launch_dependents`): they park in `wait` until this grid is done, so the launch latency of
    // back-to-back steps overlaps the tail of the previous one.  Both instructions are no-ops under a plain launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long step_before = k2_load_step(a_in);     // device step counter: in flight before anything else

    // ---------------- P0: every load of the tile is put in flight before the first barrier -----------------------
    // scalar warps fetch what their thread-per-env programs need with plain 16-B loads (they start computing as soon
    // as those ~1 KB land, long before the 18 KB history tile); env warps fetch their last action
    float4 ld0 = {0.f, 0.f, 0.f, 0.f}, ld1 = {0.f, 0.f, 0.f, 0.f};
    float4* st1 = nullptr;
    float k0 = 0.f, k1 = 0.f, k2 = 0.f, alast = 0.f, qc = 0.f;
    long long ep_pre = 0;
    int mode_off = 0;
    if (wid == W_SA) {
        if (lane < 26) ld0 = reinterpret_cast<const float4*>(r.root_states + (size_t)e0 * 13)[lane];
        const float4* g = nullptr;
        if (lane < 10) g = reinterpret_cast<const float4*>(r.commands + (size_t)e0 * 5) + lane, st1 = reinterpret_cast<float4*>(S.cmd) + lane;
        else if (lane < 20) g = reinterpret_cast<const float4*>(r.latent_c + (size_t)e0 * 5) + (lane - 10), st1 = reinterpret_cast<float4*>(S.lc) + (lane - 10);
        else if (lane < 22) g = reinterpret_cast<const float4*>(r.latent_eps + e0) + (lane - 20), st1 = reinterpret_cast<float4*>(S.eps) + (lane - 20);
        else if (lane < 26) g = reinterpret_cast<const float4*>(r.episode_length_buf + e0) + (lane - 22), st1 = reinterpret_cast<float4*>(S.ep) + (lane - 22);
        else if (lane < 28) g = reinterpret_cast<const float4*>(r.last_contacts + (size_t)e0 * 4) + (lane - 26), st1 = reinterpret_cast<float4*>(S.lcont) + (lane - 26);
        if (g != nullptr) ld1 = *g;
    } else if (wid == W_SB) {
        if (lane < 26) ld0 = reinterpret_cast<const float4*>(r.root_states + (size_t)e0 * 13)[lane];
        // feet positions: 8 envs x 4 feet x 3, three per lane
        const int el = lane >> 2, j = lane & 3;
        const float* kp = r.rigid_body_state + ((size_t)(e0 + el) * B + c.feet_indices[j]) * 13;
        k0 = kp[0], k1 = kp[1], k2 = kp[2];
    } else {
        if (lane < QA_NUM_DOF)
            alast = r.action_history_buf[(size_t)(e0 + wid) * QA_ACT_HIST_LEN * QA_NUM_DOF + (QA_ACT_HIST_LEN - 1) * QA_NUM_DOF + lane];
        // root height and episode counter: with the contact ballot of P1 they decide the reset (:168-176) before P2b says so
        else if (lane == 16) qc = r.root_states[(size_t)(e0 + wid) * 13 + 2];
        else if (lane == 17) ep_pre = r.episode_length_buf[e0 + wid];
        else if (lane >= 18 && lane < 18 + QA_DIM_C + 1 && r.mocap_clip_idx == nullptr) mode_off = r.mocap.mode_offset[lane - 18];
    }
    if (tid == 0) {
        mbar_init(&S.bar_small, 12);
        mbar_init(&S.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (wid < T2_ENVS && lane == 0) {
        // 13 TMA bulk loads, issued from 8 warps in parallel (two rounds): the 12 small tiles complete on bar_small,
        // the history tile -- 60 % of the bytes, needed last -- on its own barrier
        const unsigned n12 = T2_ENVS * 12 * 4;
        const size_t nd_ = (size_t)r.num_envs * QA_NUM_DOF;
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            void* dst = nullptr;
            const void* src = nullptr;
            unsigned bytes = 0;
            switch (wid + round * 8) {
                case 0: dst = S.dof, src = r.dof_state + (size_t)e0 * 24, bytes = T2_ENVS * 24 * 4; break;
                case 1: dst = S.cf, src = r.contact_forces + (size_t)e0 * B * 3, bytes = (unsigned)(T2_ENVS * B * 3 * 4); break;
                case 2: dst = S.act, src = r.actions + (size_t)e0 * 12, bytes = n12; break;
                case 3: dst = S.lact, src = r.last_actions + (size_t)e0 * 12, bytes = n12; break;
                case 4: dst = S.tq, src = r.torques_org + (size_t)e0 * 12, bytes = n12; break;
                case 5: dst = S.ltq, src = r.last_torques_org + (size_t)e0 * 12, bytes = n12; break;
                case 6: dst = S.ldv, src = r.last_dof_vel + (size_t)e0 * 12, bytes = n12; break;
                case 7: dst = S.msp, src = r.motor_strength + (size_t)e0 * 12, bytes = n12; break;
                case 8: dst = S.msd, src = r.motor_strength + nd_ + (size_t)e0 * 12, bytes = n12; break;
                case 9: dst = S.mass, src = r.mass_params + (size_t)e0 * 4, bytes = T2_ENVS * 4 * 4; break;
                case 10: dst = S.fric, src = r.friction_coeffs + e0, bytes = T2_ENVS * 4; break;
                case 11: dst = S.epsum, src = r.episode_sums + (size_t)e0 * QA_EPSUM_PITCH, bytes = T2_ENVS * QA_EPSUM_PITCH * 4; break;
                case 12: dst = S.hist, src = r.obs_history_buf + (size_t)e0 * HIST_W, bytes = T2_ENVS * HIST_W * 4; break;
                default: break;
            }
            if (dst != nullptr) {
                uint64_t* bar = (wid + round * 8 == 12) ? &S.bar : &S.bar_small;
                mbar_expect_tx(bar, bytes);
                bulk_load_tile(dst, src, bytes, bar);
            }
        }
    }
    const K2Step a(a_in, step_before);     // per-step scalars derived from the device counter

    bool any_state_write = a.do_push != 0;

    if (wid < T2_ENVS) {
        const int el = wid, e = e0 + el;
        const bool dl = lane < QA_NUM_DOF;
        const int d_ = dl ? lane : 0;
        STAMP(1, T_ENV);
        mbar_wait(&S.bar_small, 0);
        STAMP(2, T_ENV);
        // ---------------- P1 (env-warp half): contact-force norms, lanes = bodies --------------------------------
        // (the nine 12-wide DOF sums run on scalar warp B with lane = (env, leg): 8x fewer warp instructions)
        const float dof_pos = S.dof[el * 24 + 2 * d_], dof_vel = S.dof[el * 24 + 2 * d_ + 1];
        bool early_reset = false;
        {
            float nrm = 0.f;
            if (lane < B) {
                const float* f = S.cf + (el * B + lane) * 3;
                nrm = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
            }
            const unsigned hit_term = __ballot_sync(QA_FULL, nrm > 1.f) & c.termination_body_mask;
            {
                const long long epn = __shfl_sync(QA_FULL, ep_pre, 17) + 1;                          // :133
                const float rz = __shfl_sync(QA_FULL, qc, 16);
                early_reset = (hit_term != 0u) || ((float)epn > c.max_episode_length) || (rz < -6.0f);   // == P2b's is_reset
            }
            const unsigned hit_col = __ballot_sync(QA_FULL, nrm > 0.1f) & c.penalised_body_mask;
            float ff[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ff[j] = __shfl_sync(QA_FULL, nrm, c.feet_indices[j]);
            float out = 0.f;
            if (lane == SC_NCOL) out = (float)__popc(hit_col);
            if (lane == SC_TERM) out = hit_term ? 1.f : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (lane == SC_FF + j) out = ff[j];
            if (lane >= SC_NCOL && lane < SC_RESET) S.scal[el][lane] = out;
            // pure copies leave straight from the staged tiles (:158, :161)
            if (dl) {
                a.last_actions[(size_t)e * 12 + lane] = S.act[el * 12 + lane];
                a.last_torques_org[(size_t)e * 12 + lane] = S.tq[el * 12 + lane];
            }
        }
        bar_arrive(1, T2_THREADS);          // P1 results are in S.scal: the scalar warp may run P2b while this warp goes on
        if (early_reset) reset_precompute(c, a, e, lane, B, S.rst[el], mode_off);
        // noise of this step (:318-319) needs nothing but (env, lane, step): drawn after P1, in this warp's slack before
        // barrier 2 (a resetting env starts its precompute that much earlier)
        float nz[QA_MAX_NOISE_LANES / 32];
#pragma unroll
        for (int kk = 0; kk < QA_MAX_NOISE_LANES / 32; ++kk) {
            const int k = kk * 32 + lane;
            nz[kk] = 0.f;
            if (k < c.num_noise) {
                const int i = c.noise_idx[k];
                float u;
                if (a.noise_u != nullptr) {
                    u = a.noise_u[(size_t)e * ROW + i];
                } else {
                    Philox4 rr = philox4x32_10((uint32_t)e, SITE_NOISE0 + (i >> 2), (uint32_t)a.rng_step,
                                               (uint32_t)(a.rng_step >> 32), (uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
                    u = u32_to_unit_f32(rr.v[i & 3]);
                }
                nz[kk] = (2.f * u - 1.f) * c.noise_scale[k];
            }
        }
        STAMP(3, T_ENV);
        // ---------------- P3a: everything of the row that depends on loaded inputs only ---------------------
        // (the reset path of P3b redoes the DOF lanes for the ~1.5 % reset envs)
        float* row = S.obs + el * ROW;
        float* disc = S.disc + el * QA_NUM_OBS_DISC;
        if (dl) {
            const float dq = (dof_pos - c.default_dof_pos[lane]) * c.s_dof_pos;
            const float dv = dof_vel * c.s_dof_vel;
            row[5 + lane] = dq;
            row[17 + lane] = dv;
            row[29 + lane] = alast;
            row[45 + lane] = 0.f;
            row[66 + lane] = S.msp[el * 12 + lane] - 1.f;
            row[78 + lane] = S.msd[el * 12 + lane] - 1.f;
            disc[9 + lane] = dq;
            disc[21 + lane] = dv;
        }
        if (lane >= 12 && lane < 16) row[61 + lane - 12] = S.mass[el * 4 + lane - 12];
        if (lane == 16) row[65] = S.fric[el];
        STAMP(4, T_ENV);
        mbar_wait(&S.bar, 0);
        STAMP(5, T_ENV);
        {
            // history shift (:302-312, non-fill case), staged through registers: hazard-free in-place shift of the tile.
            // Every stored history value was clamped when it entered the buffer, so the shifted slots need no clip.
            float* h = S.hist + el * HIST_W;
            constexpr int NSH = (HIST_W - QA_NUM_PROP + 31) / 32;
            float hv[NSH];
#pragma unroll
            for (int k = 0; k < NSH; ++k) {
                const int i = k * 32 + lane;
                hv[k] = (i < HIST_W - QA_NUM_PROP) ? h[i + QA_NUM_PROP] : 0.f;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NSH; ++k) {
                const int i = k * 32 + lane;
                if (i < HIST_W - QA_NUM_PROP) {
                    const float v = clampf(hv[k], -c.clip_obs, c.clip_obs);
                    row[HIST_OFF + i] = v;
                    h[i] = v;
                }
            }
        }
        STAMP(6, T_ENV);
        // Early store of the history tile (29 % of the CTA's output bytes): nine of its ten slots are final now.  The TMA engine
        // writes the tile out while the scalar chain (P2b) and P3b run; the newest slot -- and the whole history of the ~1.5 %
        // envs that reset or refill -- are patched with plain stores in P4, after this bulk store has COMPLETED.  What the engine
        // reads out of the lanes P3b is still going to write does not matter: exactly those lanes are patched.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bar_sync(2, T2_ENVS * 32);                                             // the eight env warps: every shifted row is in S.hist
        if (wid == 2 && lane == 0 && !K2_DBG_NO_TILE_STORES) {
            bulk_store_bytes(a.obs_history_buf + (size_t)e0 * HIST_W, S.hist, T2_ENVS * HIST_W * 4);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        any_state_write = __syncthreads_or(any_state_write ? 1 : 0) != 0;      // barrier 2: P2b results are in
        STAMP(7, T_ENV);

        // ---------------- P3b: reset write (rare), P2-dependent row lanes, last history slot, noise, clip --------------
        float* sc = S.scal[el];
        float* R = S.root + el * 13;
        const bool is_reset = sc[SC_RESET] != 0.f;
        float dv_out = dof_vel;
        if (is_reset) {
            float* rs = S.rst[el];
            if (!early_reset) reset_precompute(c, a, e, lane, B, rs, mode_off);             // (the early decision is P2b's; kept for safety)
            if (dl) {
                const float mp = rs[RST_MP + lane], mv = rs[RST_MV + lane];
                S.dof[el * 24 + 2 * lane] = mp;
                S.dof[el * 24 + 2 * lane + 1] = mv;
                const float dq = (mp - c.default_dof_pos[lane]) * c.s_dof_pos;
                const float dv = mv * c.s_dof_vel;
                row[5 + lane] = dq;
                row[17 + lane] = dv;
                row[29 + lane] = 0.f;                                           // action history is cleared (:227)
                disc[9 + lane] = dq;
                disc[21 + lane] = dv;
                dv_out = mv;
            }
            if (lane < 13) R[lane] = lane < 3 ? rs[RST_POS + lane] : (lane < 7 ? rs[RST_Q + lane - 3] : rs[RST_LIN + lane - 7]);
            float* gah = a.action_history_buf + (size_t)e * QA_ACT_HIST_LEN * QA_NUM_DOF;
            for (int i = lane; i < QA_ACT_HIST_LEN * QA_NUM_DOF; i += 32) gah[i] = 0.f;
            if (lane < 4) a.feet_air_time[e * 4 + lane] = 0.f;
            if (lane < 12) disc[33 + lane] = rs[RST_KEY + lane];                // key-body positions on the mocap root
            __syncwarp();
        }
        if (dl) a.last_dof_vel[(size_t)e * 12 + lane] = dv_out;                    // :159 (post-reset dof_vel)
        // observations (:261-331): the P2-dependent row / disc lanes were written by the scalar warps (A: angles, base
        // velocities, root height, contacts, commands; B: key-body positions); a reset env overrides the lanes that see
        // its post-reset root
        if (is_reset && lane == 0) {
            const float root_h = R[2] - sc[SC_CH];
            row[57] = c.root_height_obs ? root_h : 0.f;
            disc[2] = root_h;
        }
        if (lane >= 8 && lane < 14) a.last_root_vel[(size_t)e * 6 + lane - 8] = R[7 + lane - 8];      // :160
        __syncwarp();
        {
            float* h = S.hist + el * HIST_W;
            if (S.ep[el] <= 1) {                                                   // fill: all 10 slots = current 57-vector
                int k = lane;                                                      // i % QA_NUM_PROP, carried
                for (int i = lane; i < HIST_W; i += 32) {
                    const float v = clampf(row[k], -c.clip_obs, c.clip_obs);
                    row[HIST_OFF + i] = v;
                    h[i] = v;
                    k += 32;
                    if (k >= QA_NUM_PROP) k -= QA_NUM_PROP;
                }
            } else {                                                               // newest slot only (shift done in P3a)
                for (int i = lane; i < QA_NUM_PROP; i += 32) {
                    const float v = clampf(row[i], -c.clip_obs, c.clip_obs);
                    row[HIST_OFF + HIST_W - QA_NUM_PROP + i] = v;
                    h[HIST_W - QA_NUM_PROP + i] = v;
                }
            }
        }
        __syncwarp();
        // noise on the noisy lanes only (:318-319) -- all of them lie in [0, HIST_OFF), checked at launch -- then the
        // clip (:326-328) of the only lanes that are not clamped yet
#pragma unroll
        for (int kk = 0; kk < QA_MAX_NOISE_LANES / 32; ++kk) {
            const int k = kk * 32 + lane;
            if (k < c.num_noise) {
                const int i = c.noise_idx[k];
                row[i] = row[i] + nz[kk];
            }
        }
        __syncwarp();
        for (int i = lane; i < HIST_OFF; i += 32) row[i] = clampf(row[i], -c.clip_obs, c.clip_obs);
        STAMP(8, T_ENV);
    } else if (wid == W_SA) {
        // ---------------- P2a (scalar warp A, lane = (env, role)): the thread-per-env scalar program of round 1 split four ways
        // so that the dependent chain is the longest ROLE, not their sum (same arithmetic per quantity, bit for bit):
        //   role 0  base linear velocity            role 2  projected gravity (the euler angles run on scalar warp B)
        //   role 1  base angular velocity           role 3  centre terrain height, periodic resample, push
        if (lane < 26) reinterpret_cast<float4*>(S.root)[lane] = ld0;
        if (st1 != nullptr) *st1 = ld1;
        __syncwarp();
        STAMP(12, T_SA);
        const int el = lane >> 2, role = lane & 3, e = e0 + el;
        float* R = S.root + el * 13;
        float* sc = S.scal[el];
        float* cmd = S.cmd + el * 5;
        float* row = S.obs + el * ROW;
        float* disc = S.disc + el * QA_NUM_OBS_DISC;
        const Quat q = {R[3], R[4], R[5], R[6]};
        long long ep = S.ep[el] + 1;                                                    // :133
        const float root_z_pre = R[2];                                                 // before any push / reset write
        Vec3 vec = {0.f, 0.f, 0.f};                                                    // role 0: blv, role 1: bav
        float center_h = 0.f;
        if (role == 0) {
            vec = quat_rotate_sgn(q, Vec3{R[7], R[8], R[9]}, -1.f);                     // :138
            S.blv[el * 3 + 0] = vec.x, S.blv[el * 3 + 1] = vec.y, S.blv[el * 3 + 2] = vec.z;
            row[58] = vec.x * c.s_lin_vel, row[59] = vec.y * c.s_lin_vel, row[60] = vec.z * c.s_lin_vel;
            disc[3] = vec.x * c.s_lin_vel_dist, disc[4] = vec.y * c.s_lin_vel_dist, disc[5] = vec.z * c.s_lin_vel_dist;
        } else if (role == 1) {
            vec = quat_rotate_sgn(q, Vec3{R[10], R[11], R[12]}, -1.f);                  // :139
            S.bav[el * 3 + 0] = vec.x, S.bav[el * 3 + 1] = vec.y, S.bav[el * 3 + 2] = vec.z;
            row[2] = vec.x * c.s_ang_vel, row[3] = vec.y * c.s_ang_vel, row[4] = vec.z * c.s_ang_vel;
            disc[6] = vec.x * c.s_ang_vel_dist, disc[7] = vec.y * c.s_ang_vel_dist, disc[8] = vec.z * c.s_ang_vel_dist;
        } else if (role == 2) {
            const Vec3 pg = quat_rotate_sgn(q, Vec3{0.f, 0.f, -1.f}, -1.f);             // :140 (euler angles: env warps)
            S.pg[el * 3 + 0] = pg.x, S.pg[el * 3 + 1] = pg.y, S.pg[el * 3 + 2] = pg.z;
        } else {
            if (c.measure_heights) center_h = terrain_center_height(a.terrain, yaw_quat(q), R[0], R[1], c.center_px, c.center_py);
            const float root_h_pre = root_z_pre - center_h;                             // pre-reset quantities (:263-291)
            row[57] = c.root_height_obs ? root_h_pre : 0.f;
            disc[2] = root_h_pre;
            if (divisible_by(ep, c.resample_period)) {                                 // :454-462
                const K2Draw d = draw_site(c, a, e, SITE_RS0, a.rs_eps_u, a.rs_c_idx, a.rs_cmd_u);
                resample_thread(c, d, cmd, S.eps + el, S.lc + el * 5);
            }
            if (a.do_push) {                                                           // :682-687
                float u0, u1;
                if (a.push_u != nullptr) {
                    u0 = a.push_u[e * 2 + 0];
                    u1 = a.push_u[e * 2 + 1];
                } else {
                    Philox4 rr = philox4x32_10((uint32_t)e, SITE_PUSH, (uint32_t)a.rng_step, (uint32_t)(a.rng_step >> 32),
                                               (uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
                    u0 = u32_to_unit_f32(rr.v[0]);
                    u1 = u32_to_unit_f32(rr.v[1]);
                }
                const float span = c.max_push_vel_xy - (-c.max_push_vel_xy);
                R[7] = span * u0 + (-c.max_push_vel_xy);
                R[8] = span * u1 + (-c.max_push_vel_xy);
            }
        }
        // role 0 runs P2b and needs the other roles' results: registers through shuffles, commands through shared memory
        const Vec3 blv = vec;                                                          // valid on role 0
        Vec3 bav;
        bav.x = __shfl_down_sync(QA_FULL, vec.x, 1);
        bav.y = __shfl_down_sync(QA_FULL, vec.y, 1);
        bav.z = __shfl_down_sync(QA_FULL, vec.z, 1);
        center_h = __shfl_down_sync(QA_FULL, center_h, 3);
        __syncwarp();                                                                  // role 3's command / root writes are visible
        STAMP(13, T_SA);
        mbar_wait(&S.bar_small, 0);          // episode sums tile (TMA) visible to this warp
        bar_sync(1, T2_THREADS);             // P1 sums / force norms of the env warps are in S.scal
        STAMP(14, T_SA);
        if (role == 0) {
            // ---------------- P2b: contacts, termination, reward total, episode sums, reset decision --------------------
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                              // :143-146
                const float f = sc[SC_FF + j];
                const bool ct = f > 2.f;
                S.ff[el * 4 + j] = f;
                S.cont_out[el * 4 + j] = ct ? 1 : 0;
                const bool cfl = ct || S.lcont[el * 4 + j] != 0;
                S.cfilt_out[el * 4 + j] = cfl ? 1 : 0;
                const float cf_ = cfl ? 1.f : 0.f;                                       // :285, :275, :323-324
                S.obs[el * ROW + 41 + j] = cf_ - 0.5f;
                S.disc[el * QA_NUM_OBS_DISC + 45 + j] = cf_ * c.s_foot_contact;
                if (a.contact_buf) a.contact_buf[((size_t)e * a.contact_ring_len + a.contact_ring_head) * 4 + j] = cf_;
                if (a.contact_force_buf)
                    a.contact_force_buf[((size_t)e * a.contact_ring_len + a.contact_ring_head) * 4 + j] =
                        clampf(f, -c.clip_obs, c.clip_obs);
            }
            const bool time_out = ((float)ep > c.max_episode_length) || (root_z_pre < -6.0f);   // :168-176
            const bool is_reset = (sc[SC_TERM] != 0.f) || time_out;
            const float root_h_pre = root_z_pre - center_h;
            float rt[QA_NUM_REWARDS];                                                  // :1248-1335, dir() order
            rt[0] = sc[SC_SUM0 + 0];
            rt[1] = sc[SC_NCOL];
            rt[2] = sc[SC_SUM0 + 1];
            rt[3] = sc[SC_SUM0 + 2];
            rt[4] = sc[SC_SUM0 + 3];
            rt[5] = sc[SC_SUM0 + 4];
            rt[6] = sc[SC_SUM0 + 5];
            rt[7] = sc[SC_SUM0 + 6];
            {
                const float err = sqrtf((cmd[3] - root_h_pre) * (cmd[3] - root_h_pre));
                rt[8] = ((err < 0.05f) && (cmd[3] >= c.jump_height_lo)) ? c.jump_goal : 0.f;
            }
            {
                const float err = sqrtf((cmd[4] - root_h_pre) * (cmd[4] - root_h_pre));
                const float rl = expf(-10.0f * (err * err) / c.tracking_sigma);
                rt[9] = (!(cmd[3] > c.jump_height_lo)) ? rl : 0.f;
            }
            rt[10] = sc[SC_SUM0 + 7];
            rt[11] = sc[SC_SUM0 + 8];
            {
                const float dw = cmd[2] - bav.z;
                rt[12] = expf(-(dw * dw) / c.tracking_sigma);
                const float dx = cmd[0] - blv.x, dy = cmd[1] - blv.y;
                rt[13] = expf(-(dx * dx + dy * dy) / c.tracking_sigma);
            }
            float rew = 0.f;
            float* es = S.epsum + el * QA_EPSUM_PITCH;
#pragma unroll
            for (int k = 0; k < QA_NUM_REWARDS; ++k) {
                const float t = rt[k] * c.reward_scale[k];
                rew = rew + t;
                es[k] = es[k] + t;
            }
            if (c.only_positive_rewards) rew = fmaxf(rew, 0.f);
            int mode = 0;
            if (is_reset) {                                                            // :178-240 (scalar half)
                K2Workspace* ws = reinterpret_cast<K2Workspace*>(a.workspace);
#pragma unroll
                for (int k = 0; k < QA_NUM_REWARDS; ++k) {
                    atomicAdd(&ws->sums[k], (double)es[k]);
                    es[k] = 0.f;
                }
                atomicAdd(&ws->reset_count, 1u);
                const K2Draw d = draw_site(c, a, e, SITE_RT0, a.rt_eps_u, a.rt_c_idx, a.rt_cmd_u);
                resample_thread(c, d, cmd, S.eps + el, S.lc + el * 5);
                mode = d.c_idx;
                ep = 0;
            }
            {
                float* rowc = S.obs + el * ROW + CMD_OFF;                                // [commands 5 | eps 1 | c 5] (:314-316)
#pragma unroll
                for (int k = 0; k < 5; ++k) rowc[k] = clampf(cmd[k], -c.clip_obs, c.clip_obs);
                rowc[5] = clampf(S.eps[el], -c.clip_obs, c.clip_obs);
#pragma unroll
                for (int k = 0; k < 5; ++k) rowc[6 + k] = clampf(S.lc[el * 5 + k], -c.clip_obs, c.clip_obs);
            }
            S.ep[el] = ep;
            S.rew[el] = rew;
            S.rooth[el] = root_h_pre;
            sc[SC_RESET] = is_reset ? 1.f : 0.f;
            sc[SC_CH] = center_h;
            sc[SC_MODE] = __int_as_float(mode);
            a.reset_buf[e] = is_reset ? 1 : 0;
            a.time_out_buf[e] = time_out ? 1 : 0;
            any_state_write = any_state_write || is_reset;
        }
        STAMP(15, T_SA);
        any_state_write = __syncthreads_or(any_state_write ? 1 : 0) != 0;      // barrier 2
        // the per-env scalar outputs leave with plain 16-B stores while the env warps assemble the rows
        copy16(a.episode_sums + (size_t)e0 * QA_EPSUM_PITCH, S.epsum, T2_ENVS * QA_EPSUM_PITCH / 4, lane);
        copy16(a.commands + (size_t)e0 * 5, S.cmd, T2_ENVS * 5 / 4, lane);
        copy16(a.latent_c + (size_t)e0 * 5, S.lc, T2_ENVS * 5 / 4, lane);
        copy16(a.latent_eps + e0, S.eps, T2_ENVS / 4, lane);
        copy16(a.episode_length_buf + e0, S.ep, T2_ENVS * 2 / 4, lane);
        copy16(a.last_contacts + (size_t)e0 * 4, S.cont_out, T2_ENVS / 4, lane);
        copy16(a.contact_filt + (size_t)e0 * 4, S.cfilt_out, T2_ENVS / 4, lane);
        copy16(a.feet_forces + (size_t)e0 * 4, S.ff, T2_ENVS, lane);
        copy16(a.rew_buf + e0, S.rew, T2_ENVS / 4, lane);
        copy16(a.root_h + e0, S.rooth, T2_ENVS / 4, lane);
        copy16(a.base_lin_vel + (size_t)e0 * 3, S.blv, T2_ENVS * 3 / 4, lane);
        copy16(a.base_ang_vel + (size_t)e0 * 3, S.bav, T2_ENVS * 3 / 4, lane);
        copy16(a.projected_gravity + (size_t)e0 * 3, S.pg, T2_ENVS * 3 / 4, lane);
        copy16(a.rpy + (size_t)e0 * 3, S.rpy, T2_ENVS * 3 / 4, lane);
        STAMP(18, T_SA);
    } else {
        // ---------------- P2a (scalar warp B): key-body positions in the heading frame (:1377-1396) -----------------------
        if (lane < 26) reinterpret_cast<float4*>(S.rootB)[lane] = ld0;
        S.key[lane * 3 + 0] = k0, S.key[lane * 3 + 1] = k1, S.key[lane * 3 + 2] = k2;
        __syncwarp();
        {
            // lane = (env, foot): the heading quaternion is computed 4x redundantly, the rotations run 32 wide
            const int el = lane >> 2, j = lane & 3;
            const float* R = S.rootB + el * 13;
            const Quat hq = heading_quat_inv(Quat{R[3], R[4], R[5], R[6]});
            const Vec3 local = {k0 - R[0], k1 - R[1], k2 - R[2]};
            const Vec3 o = quat_rotate_sgn(hq, local, 1.f);
            float* disc = S.disc + el * QA_NUM_OBS_DISC;                                 // :272-274, on the pre-reset root
            disc[33 + j * 3 + 0] = o.x * c.s_key_pos, disc[33 + j * 3 + 1] = o.y * c.s_key_pos, disc[33 + j * 3 + 2] = o.z * c.s_key_pos;
        }
        STAMP(19, T_SB);
        // ---------------- P1 (DOF half): the nine 12-wide reward sums, lane = (env, leg), 3 DOFs per lane -----------------
        mbar_wait(&S.bar_small, 0);
        {
            const int el = lane >> 2, leg = lane & 3;
            float s[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) s[k] = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int d = leg * 3 + j;
                const float dof_pos = S.dof[el * 24 + 2 * d], dof_vel = S.dof[el * 24 + 2 * d + 1];
                const float act = S.act[el * 12 + d], lact = S.lact[el * 12 + d];
                const float tq = S.tq[el * 12 + d], ltq = S.ltq[el * 12 + d], ldv = S.ldv[el * 12 + d];
                float v;
                v = lact - act;
                s[0] = s[0] + v * v;                                                       // action_rate
                v = tq - ltq;
                s[1] = s[1] + v * v;                                                       // delta_torques
                v = (ldv - dof_vel) / c.dt;
                s[2] = s[2] + v * v;                                                       // dof_acc
                const float dq0 = dof_pos - c.default_dof_pos[d];
                s[3] = s[3] + dq0 * dq0;                                                   // dof_error
                v = -fminf(dof_pos - c.dof_pos_lower[d], 0.f);
                v = v + fmaxf(dof_pos - c.dof_pos_upper[d], 0.f);
                s[4] = s[4] + v;                                                           // dof_pos_limits
                v = clampf(fabsf(dof_vel) - c.dof_vel_limits[d] * c.soft_dof_vel_limit, 0.f, 1.f);
                s[5] = s[5] + v;                                                           // dof_vel_limits
                if ((c.hip_dof_mask >> d) & 1u) s[6] = s[6] + dq0 * dq0;                   // hip_pos
                v = fmaxf(fabsf(tq) - c.torque_limits[d] * c.soft_torque_limit, 0.f);
                s[7] = s[7] + v;                                                           // torque_limits
                s[8] = s[8] + tq * tq;                                                     // torques
            }
            float* sc = S.scal[el];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                s[k] += __shfl_xor_sync(QA_FULL, s[k], 1);
                s[k] += __shfl_xor_sync(QA_FULL, s[k], 2);
                if (leg == 0) sc[SC_SUM0 + k] = s[k];
            }
        }
        STAMP(15 + 5, T_SB);
        bar_arrive(1, T2_THREADS);
        {
            // euler angles of the pre-reset root (:141, get_euler_xyz), lane = (env, j): j = 0 roll and j = 2 yaw run the two
            // atan2f side by side, j = 1 the asinf -- in this warp's slack before barrier 2 (they were the longest role of scalar
            // warp A's chain: ~2 us of dependent issue).  Same expressions, op for op, as the scalar program had.
            const int el = lane >> 2, j = lane & 3;
            const float* R = S.rootB + el * 13;
            const Quat q = {R[3], R[4], R[5], R[6]};
            float ang = 0.f;
            if (j == 0 || j == 2) {
                const float num = j == 0 ? 2.0f * (q.w * q.x + q.y * q.z) : 2.0f * (q.w * q.z + q.x * q.y);
                const float den = j == 0 ? 1.0f - 2.0f * (q.x * q.x + q.y * q.y) : 1.0f - 2.0f * (q.y * q.y + q.z * q.z);
                ang = atan2f(num, den);
            } else if (j == 1) {
                float t2 = 2.0f * (q.w * q.y - q.z * q.x);
                t2 = clampf(t2, -1.f, 1.f);
                ang = asinf(t2);
            }
            if (j < 3) S.rpy[el * 3 + j] = ang;
            if (j < 2) {
                S.obs[el * ROW + j] = ang;
                S.disc[el * QA_NUM_OBS_DISC + j] = ang;
            }
        }
        any_state_write = __syncthreads_or(any_state_write ? 1 : 0) != 0;      // barrier 2
    }

    // ---------------- P4: the four big tiles leave through the TMA engine, one issuing thread each ------------------
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    STAMP(9, T_ENV);
    if (wid == 2 && !K2_DBG_NO_TILE_STORES) {
        // history: the early bulk store has to be complete before its stale lanes are overwritten (write-after-write on the
        // same addresses from two proxies); then the newest slot of every env, everything of a reset / refilled env
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll 1
        for (int el = 0; el < T2_ENVS; ++el) {
            float* gh = a.obs_history_buf + (size_t)(e0 + el) * HIST_W;
            const float* h = S.hist + el * HIST_W;
            const int lo = (S.ep[el] <= 1) ? 0 : HIST_W - QA_NUM_PROP;
            for (int i = lo + lane; i < HIST_W; i += 32) gh[i] = h[i];
        }
    }
    if (lane == 0 && wid < 6 && wid != 2) {
        bool issued = true;
        switch (K2_DBG_NO_TILE_STORES && wid < 4 ? 99 : wid) {
            case 99: issued = false; break;
            case 0: bulk_store_bytes(a.obs_buf + (size_t)e0 * ROW, S.obs, T2_ENVS * ROW * 4); break;
            case 1:
                if (a.privileged_obs_buf != a.obs_buf) bulk_store_bytes(a.privileged_obs_buf + (size_t)e0 * ROW, S.obs, T2_ENVS * ROW * 4);
                else issued = false;
                break;
            case 3: bulk_store_bytes(a.obs_disc_buf + (size_t)e0 * QA_NUM_OBS_DISC, S.disc, T2_ENVS * QA_NUM_OBS_DISC * 4); break;
            case 4:                                     // simulator memory is only written on reset / push
                if (any_state_write) bulk_store_bytes(a.root_states + (size_t)e0 * 13, S.root, T2_ENVS * 13 * 4);
                else issued = false;
                break;
            default:
                if (any_state_write) bulk_store_bytes(a.dof_state + (size_t)e0 * 24, S.dof, T2_ENVS * 24 * 4);
                else issued = false;
                break;
        }
        STAMP(10, T_ENV);
        if (issued) {
            // the tiles become visible to the next kernel at kernel end; the shared-memory source must outlive the reads
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    STAMP(11, T_ENV);

    // ---------------- epilogue: the CTA that takes the last ticket finalises the step (no extra kernel launch) --------
    // Every CTA's direct stores (time_out_buf) and reset-statistics atomics precede barrier 3; thread 0's fence after that
    // barrier is cumulative, so whoever observes the full ticket count also observes them.  The TMA tile stores are not
    // needed by the finaliser.
    if (tid == 0) {
        K2Workspace* ws = reinterpret_cast<K2Workspace*>(a.workspace);
        __threadfence();
        S.last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (S.last) {
        __threadfence();
        K2Workspace* ws = reinterpret_cast<K2Workspace*>(a.workspace);
        const unsigned cnt = *reinterpret_cast<volatile unsigned*>(&ws->reset_count);
        if (cnt > 0) {
            if (tid < QA_NUM_REWARDS) {                                     // :230-234
                const double sum = *reinterpret_cast<volatile double*>(&ws->sums[tid]);
                a.episode_rew_means[tid] = (float)(sum / (double)cnt) / c.episode_length_s;
            }
            // extras["time_outs"] = time_out_buf, only on steps with >= 1 reset (:239-240); num_envs % 8 == 0
            const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.time_out_buf);
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.time_outs_latched);
            for (int i = tid; i < a.num_envs / 8; i += T2_THREADS) dst[i] = __ldcg(src + i);
        }
        __syncthreads();
        if (tid < QA_NUM_REWARDS) ws->sums[tid] = 0.0;
        if (tid == 0) {
            *a.num_resets = (int)cnt;
            if (a.step_state != nullptr) a.step_state[0] = (long long)a.rng_step;
            ws->reset_count = 0u;
            ws->ticket = 0u;
        }
    }
    GSTAMP(17, T_ENV);
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// Returns 1 if the tiled kernel was launched, 0 if the arguments do not qualify (caller falls back), <0 / >0 error.
int qa_k2_try_launch_tiled(const QaBbcConst* c, const QaBbcStepArgs* a, cudaStream_t stream, int* launched) {
    *launched = 0;
    if (!(a->flags & QA_K2_TILED)) return 0;
    if (a->num_envs % T2_ENVS != 0 || a->obs_pitch != QA_OBS_WIDTH || c->num_bodies > 32) return 0;
    if (c->num_noise < 0 || c->num_noise > QA_MAX_NOISE_LANES) return QA_ERANGE;
    for (int k = 0; k < c->num_noise; ++k)                       // the tiled kernel clamps history / command lanes as it writes them
        if (c->noise_idx[k] < 0 || c->noise_idx[k] >= HIST_OFF) return 0;
    const void* ptrs[] = {a->obs_buf, a->privileged_obs_buf, a->obs_history_buf, a->obs_disc_buf, a->root_states,
                          a->dof_state, a->contact_forces, a->actions, a->last_actions, a->torques_org,
                          a->last_torques_org, a->last_dof_vel, a->last_root_vel, a->motor_strength, a->commands,
                          a->latent_c, a->latent_eps, a->mass_params, a->friction_coeffs, a->episode_sums,
                          a->episode_length_buf, a->last_contacts, a->contact_filt, a->feet_forces, a->rew_buf,
                          a->root_h, a->base_lin_vel, a->base_ang_vel, a->projected_gravity, a->rpy, a->time_out_buf,
                          a->time_outs_latched};
    for (const void* p : ptrs)
        if (!aligned16(p)) return 0;
    if ((((size_t)a->num_envs * QA_NUM_DOF * 4) & 15u) != 0) return 0;       // second motor_strength plane
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_post_physics_bbc_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(T2Smem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    *launched = 1;
    if (a->flags & QA_K2_PDL) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(a->num_envs / T2_ENVS), cfg.blockDim = dim3(T2_THREADS);
        cfg.dynamicSmemBytes = sizeof(T2Smem), cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_post_physics_bbc_tiled, *c, *a);
        return e == cudaSuccess ? 0 : (int)e;
    }
    k_post_physics_bbc_tiled<<<a->num_envs / T2_ENVS, T2_THREADS, sizeof(T2Smem), stream>>>(*c, *a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

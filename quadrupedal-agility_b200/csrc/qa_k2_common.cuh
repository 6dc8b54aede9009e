// Shared pieces of the two K2 (fused BBC post-physics) kernels: row layout, Philox sites, the dense-draw /
// in-kernel-RNG switch, command resampling, centre terrain height, reset-statistics workspace, TMA bulk helpers.
#pragma once
#include "qa_b200.h"
#include "qa_common.cuh"
#include "qa_mocap.cuh"

#define ROW QA_OBS_WIDTH                // 671
#define HIST_OFF 90                     // 57 + 4 + 29
#define HIST_W (QA_HIST_LEN * QA_NUM_PROP)   // 570
#define CMD_OFF (HIST_OFF + HIST_W)     // 660

// per-warp staging layout (floats)
#define S_ROOT 0                        // 13 (+3)
#define S_CMD 16                        // commands 5, eps 1, c 5 (+1)
#define S_KEY 28                        // 12 feet positions
#define S_DISC 40                       // 49 (+3) disc obs
#define S_MISC 92                       // mass 4, friction 1 (+3)
#define S_CF 100                        // up to 32*3 contact forces
#define S_TOTAL 196

// Philox sites (perf-mode RNG); counter = (env, site, step_lo, step_hi), key = seed
#define SITE_RS0 8
#define SITE_RS1 9
#define SITE_RT0 10
#define SITE_RT1 11
#define SITE_PUSH 12
#define SITE_MOCAP 13
#define SITE_NOISE0 16                  // + element/4

// Step arguments with the per-step scalars resolved: either the host-supplied fields or, under CUDA-graph
// replay, values derived from the device-resident step counter (see QaBbcStepArgs::step_state).
__device__ __forceinline__ long long k2_mod(long long v, int m) {      // counters fit 32 bits in practice
    const unsigned long long u = (unsigned long long)v;
    if ((u >> 32) == 0ull) return (long long)((unsigned)u % (unsigned)m);
    return v % (long long)m;
}
__device__ __forceinline__ long long k2_load_step(const QaBbcStepArgs& in) {
    return in.step_state != nullptr ? *reinterpret_cast<const volatile long long*>(in.step_state) : 0ll;
}
struct K2Step : QaBbcStepArgs {
    __device__ __forceinline__ K2Step(const QaBbcStepArgs& in, long long before) : QaBbcStepArgs(in) {
        if (in.step_state != nullptr) {
            const long long cnt = before + 1;
            rng_step = (uint64_t)cnt;
            do_push = (in.push_interval > 0 && k2_mod(cnt, in.push_interval) == 0) ? 1 : 0;
            contact_ring_head = in.contact_ring_len > 0 ? (int)k2_mod(before, in.contact_ring_len) : 0;
        }
    }
    __device__ __forceinline__ explicit K2Step(const QaBbcStepArgs& in) : K2Step(in, k2_load_step(in)) {}
};

struct K2Draw {
    double eps_u;
    int c_idx;
    float cmd_u[5];
};

// mode ~ Categorical(softmax(prior / T)) by CDF inversion; the CDF is the live device copy when the host supplies one
__device__ __forceinline__ int pick_mode(const QaBbcConst& c, const float* cdf_dev, float u) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < QA_DIM_C - 1; ++i) k += (u >= (cdf_dev != nullptr ? __ldcg(cdf_dev + i) : c.prior_cdf[i])) ? 1 : 0;
    return k;
}

__device__ __forceinline__ K2Draw draw_site(const QaBbcConst& c, const QaBbcStepArgs& a, int e, int site0,
                                            const double* eps_u, const int32_t* c_idx, const float* cmd_u) {
    K2Draw d;
    if (eps_u != nullptr) {
        d.eps_u = eps_u[e];
        d.c_idx = c_idx[e];
#pragma unroll
        for (int k = 0; k < 5; ++k) d.cmd_u[k] = cmd_u[e * 5 + k];
    } else {
        const uint32_t slo = (uint32_t)a.rng_step, shi = (uint32_t)(a.rng_step >> 32);
        const uint32_t k0 = (uint32_t)a.rng_seed, k1 = (uint32_t)(a.rng_seed >> 32);
        Philox4 r0 = philox4x32_10((uint32_t)e, site0, slo, shi, k0, k1);
        Philox4 r1 = philox4x32_10((uint32_t)e, site0 + 1, slo, shi, k0, k1);
        d.eps_u = u64_to_unit_f64(r0.v[0], r0.v[1]);
        d.c_idx = pick_mode(c, a.prior_cdf, u32_to_unit_f32(r0.v[2]));
        d.cmd_u[0] = u32_to_unit_f32(r0.v[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) d.cmd_u[k + 1] = u32_to_unit_f32(r1.v[k]);
    }
    d.c_idx = min(max(d.c_idx, 0), QA_DIM_C - 1);
    return d;
}

// _resample_latent_eps / _resample_latent_c / _resample_commands (:474-540) for one env.
// cmd = smem pointer to [commands 5 | eps 1 | c 5]; every lane computes, lane 0 writes.
__device__ __forceinline__ void resample_env(const QaBbcConst& c, const K2Draw& d, float* cmd, int lane) {
    const int m = d.c_idx;
    const float eps_new = (float)(d.eps_u * 2. - 1.);
    float n0 = (c.lin_vel_x[m][1] - c.lin_vel_x[m][0]) * d.cmd_u[0] + c.lin_vel_x[m][0];
    float n1 = (c.lin_vel_y[m][1] - c.lin_vel_y[m][0]) * d.cmd_u[1] + c.lin_vel_y[m][0];
    float n2 = (c.ang_vel_yaw[m][1] - c.ang_vel_yaw[m][0]) * d.cmd_u[2] + c.ang_vel_yaw[m][0];
    const float jump = (m == QA_DIM_C - 1) ? 1.f : 0.f;
    const float n3 = (c.jump_h_span * d.cmd_u[3] + c.jump_h_lo) * jump;
    const float n4 = (c.loco_h_span * d.cmd_u[4] + c.loco_h_lo) * (1.f - jump);
    n0 *= (fabsf(n0) > c.lin_vel_x_clip) ? 1.f : 0.f;
    n1 *= (fabsf(n1) > c.lin_vel_y_clip) ? 1.f : 0.f;
    n2 *= (fabsf(n2) > c.ang_vel_yaw_clip) ? 1.f : 0.f;
    __syncwarp();
    if (lane == 0) {
        cmd[0] = n0;
        cmd[1] = n1;
        cmd[2] = n2;
        cmd[3] = n3;
        cmd[4] = n4;
        cmd[5] = eps_new;
#pragma unroll
        for (int k = 0; k < QA_DIM_C; ++k) cmd[6 + k] = (k == m) ? 1.f : 0.f;
    }
    __syncwarp();
}

__device__ __forceinline__ float terrain_center_height(const QaTerrain& t, Quat yawq, float bx, float by, float hx,
                                                        float hy) {
    Vec3 p = quat_apply(yawq, Vec3{hx, hy, 0.f});
    float wx = (p.x + bx) + t.border_size;
    float wy = (p.y + by) + t.border_size;
    long long ix = (long long)(wx / t.horizontal_scale);
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

struct K2Workspace {
    double sums[QA_NUM_REWARDS];   // 112 B
    unsigned int reset_count;      // 112
    unsigned int ticket;           // 116
};

__device__ __forceinline__ void bulk_store_tile(float* gdst, const float* ssrc, unsigned bytes) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
                 : "memory");
}


// ---- mbarrier + TMA bulk load helpers (cp.async.bulk, sm_90+) ---------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_load_tile(void* sdst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(sdst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store_bytes(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}

// Mocap frame blending shared by the standalone K4 kernel and the reset path of K2.
//
// Follows MotionLoader.get_full_frame_at_time_batch / traj_time_sample_batch
// (bbc/rsl_rl/datasets/motion_loader.py:333-341, 410-447) and quaternion_slerp
// (bbc/rsl_rl/utils/utils.py:126-159): frame index math in float64 (numpy), values in fp32.
#pragma once
#include "qa_b200.h"
#include "qa_common.cuh"

struct MocapBlendIdx {
    int row_lo, row_hi;   // rows of the flat (F,49) table
    float blend;          // fp32(p*n - floor(p*n))
};

// traj_time_sample_batch + the index half of get_full_frame_at_time_batch
// operands: the chosen clip's length [s], frame duration [s], frame count and first row
__device__ __forceinline__ MocapBlendIdx mocap_blend_index_meta(double len_s, double frame_dur, double n, int start, double time_u,
                                                                double time_between_frames, int disc_obs_len) {
    const double subst = time_between_frames * (double)disc_obs_len + frame_dur;
    double t = (len_s - subst) * time_u;
    t = fmax(0.0 + 1e-7, t);
    const double p = t / len_s;
    const double pn = p * n;
    const double lo = floor(pn), hi = ceil(pn);
    MocapBlendIdx r;
    int ilo = (int)lo, ihi = (int)hi;
    // the reference would raise on an out-of-range row; clamp so that a bad draw cannot fault
    const int last = (int)n - 1;
    ilo = min(max(ilo, 0), last);
    ihi = min(max(ihi, 0), last);
    r.row_lo = start + ilo;
    r.row_hi = start + ihi;
    r.blend = (float)(pn - lo);
    return r;
}

__device__ __forceinline__ MocapBlendIdx mocap_blend_index(const QaMocapTable& tb, int clip, double time_u,
                                                           double time_between_frames, int disc_obs_len) {
    return mocap_blend_index_meta(tb.clip_len_s[clip], tb.clip_frame_dur[clip], tb.clip_nframes[clip], tb.clip_start[clip], time_u,
                                  time_between_frames, disc_obs_len);
}

// quaternion_slerp, utils.py:126-159 (spin=0, shortestpath=True).  NB the reference scales by
// 1/angle (":154"), not 1/sin(angle); reproduced as is.
__device__ __forceinline__ Quat slerp_ref(Quat q0, Quat q1, float f) {
    const float EPSF = (float)(2.220446049250313e-16 * 4.0);
    // torch.isclose(fraction, 0) / (fraction, 1) with rtol=1e-5, atol=1e-8 evaluated in fp32
    const bool zero_mask = (f == 0.f) || (fabsf(f - 0.f) <= 1e-8f + fabsf(1e-5f * 0.f));
    const bool ones_mask = (f == 1.f) || (fabsf(f - 1.f) <= 1e-8f + fabsf(1e-5f * 1.f));
    float d = q0.x * q1.x + q0.y * q1.y + q0.z * q1.z + q0.w * q1.w;
    const bool dist_mask = fabsf(fabsf(d) - 1.0f) < EPSF;
    const Quat q1_in = q1;   // out[ones_mask] = q1 is taken BEFORE the shortest-path sign flip
    if (d < 0.f) {
        d = -d;
        q1.x = -q1.x;
        q1.y = -q1.y;
        q1.z = -q1.z;
        q1.w = -q1.w;
    }
    d = clampf(d, -1.f, 1.f);
    const float angle = acosf(d);
    const bool angle_mask = fabsf(angle) < EPSF;
    if (angle_mask || dist_mask) return q0;
    if (ones_mask) return q1_in;
    if (zero_mask) return q0;
    const float isin = 1.0f / angle;
    const float s0 = sinf((1.0f - f) * angle) * isin;
    const float s1 = sinf(f * angle) * isin;
    Quat o;
    o.x = q0.x * s0 + q1.x * s1;
    o.y = q0.y * s0 + q1.y * s1;
    o.z = q0.z * s0 + q1.z * s1;
    o.w = q0.w * s0 + q1.w * s1;
    return o;
}

// Blended value of column `col` (0..48) of the 49-float frame; `rot` is the slerped quaternion
// (computed once per env by the caller).
__device__ __forceinline__ float mocap_lerp(float a, float b, float blend) { return (1.0f - blend) * a + blend * b; }

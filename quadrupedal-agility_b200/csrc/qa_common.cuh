// Shared device helpers for the qa_b200 kernels (sm_100a).
//
// Arithmetic note: the env kernels are compiled with -fmad=false and written in the same
// operation order as the reference's PyTorch eager ops (each aten op rounds separately), so
// that the quantised terrain indices and the bool masks are bit-exact against the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define QA_WARP 32
#define QA_FULL 0xffffffffu

#define QA_CHECK_PTR(p)            \
    do {                           \
        if ((p) == nullptr) return QA_EINVAL; \
    } while (0)

#define QA_LAUNCH_RET()                                   \
    do {                                                  \
        cudaError_t e__ = cudaGetLastError();             \
        return e__ == cudaSuccess ? 0 : (int)e__;         \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(QA_FULL, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(QA_FULL, v, o);
    return v;
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// streaming loads/stores: state rows are touched once per step
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

// ---- quaternion helpers (xyzw), isaacgym.torch_utils definitions ----------
struct Vec3 {
    float x, y, z;
};
struct Quat {
    float x, y, z, w;
};

__device__ __forceinline__ Vec3 cross3(Vec3 a, Vec3 b) {
    Vec3 r;
    r.x = a.y * b.z - a.z * b.y;
    r.y = a.z * b.x - a.x * b.z;
    r.z = a.x * b.y - a.y * b.x;
    return r;
}

// quat_rotate / quat_rotate_inverse: a +/- b + c with
//   a = v (2 w^2 - 1), b = 2 w (q_v x v), c = 2 q_v (q_v . v)
__device__ __forceinline__ Vec3 quat_rotate_sgn(Quat q, Vec3 v, float sgn) {
    float s = 2.0f * (q.w * q.w) - 1.0f;
    Vec3 a = {v.x * s, v.y * s, v.z * s};
    Vec3 cr = cross3(Vec3{q.x, q.y, q.z}, v);
    Vec3 b = {cr.x * q.w * 2.0f, cr.y * q.w * 2.0f, cr.z * q.w * 2.0f};
    float d = q.x * v.x + q.y * v.y + q.z * v.z;
    Vec3 c = {q.x * d * 2.0f, q.y * d * 2.0f, q.z * d * 2.0f};
    Vec3 r;
    if (sgn > 0.f) {
        r.x = a.x + b.x + c.x;
        r.y = a.y + b.y + c.y;
        r.z = a.z + b.z + c.z;
    } else {
        r.x = a.x - b.x + c.x;
        r.y = a.y - b.y + c.y;
        r.z = a.z - b.z + c.z;
    }
    return r;
}

// quat_apply(a, b) = b + w t + xyz x t, t = 2 (xyz x b)
__device__ __forceinline__ Vec3 quat_apply(Quat q, Vec3 b) {
    Vec3 xyz = {q.x, q.y, q.z};
    Vec3 t = cross3(xyz, b);
    t.x *= 2.f;
    t.y *= 2.f;
    t.z *= 2.f;
    Vec3 u = cross3(xyz, t);
    Vec3 r = {b.x + q.w * t.x + u.x, b.y + q.w * t.y + u.y, b.z + q.w * t.z + u.z};
    return r;
}

// quat_apply_yaw's quaternion: zero x,y then normalize (torch_jit_utils.py:117-122)
__device__ __forceinline__ Quat yaw_quat(Quat q) {
    float n = sqrtf(q.z * q.z + q.w * q.w);
    n = fmaxf(n, 1e-9f);
    Quat r = {0.f / n, 0.f / n, q.z / n, q.w / n};
    return r;
}

// calc_heading_quat_inv (torch_jit_utils.py:64-75): rotation by -heading about z
__device__ __forceinline__ Quat heading_quat_inv(Quat q) {
    Vec3 ref = {1.f, 0.f, 0.f};
    Vec3 rd = quat_rotate_sgn(q, ref, 1.f);
    float heading = atan2f(rd.y, rd.x);
    float theta = (-heading) / 2.f;
    // normalize(axis) with axis = (0,0,1) is exact; xyz = axis * sin(theta)
    float s = sinf(theta), c = cosf(theta);
    Quat r = {0.f * s, 0.f * s, 1.f * s, c};
    float n = fmaxf(sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w), 1e-9f);
    r.x /= n;
    r.y /= n;
    r.z /= n;
    r.w /= n;
    return r;
}

// ---- Philox4x32-10 (counter based; perf-mode RNG) -----------------------------------------
struct Philox4 {
    uint32_t v[4];
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0;
    o.v[1] = c1;
    o.v[2] = c2;
    o.v[3] = c3;
    return o;
}

// torch.rand-style fp32 uniform in [0,1): 24 random mantissa bits
__device__ __forceinline__ float u32_to_unit_f32(uint32_t x) { return (float)(x & 0x00ffffffu) * (1.0f / 16777216.0f); }
// numpy-style fp64 uniform in [0,1): 53 random bits
__device__ __forceinline__ double u64_to_unit_f64(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

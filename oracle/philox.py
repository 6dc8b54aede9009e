"""CPU ORACLE (test infrastructure, not product code): Philox4x32-10 and the counter layout of the env kernels' production RNG
(`quadrupedal-agility_b200/csrc/qa_common.cuh philox4x32_10`, `qa_k2_common.cuh draw_site`, SITE_* constants), in numpy.

The reference draws from torch / numpy / multinomial generators on variable-length index sets (legged_robot.py:504-540,
motion_loader.py:311-341); the kernels' production mode replaces them by ONE counter-based stream
    value = Philox4x32-10(counter = (env, site, step_lo, step_hi), key = (seed_lo, seed_hi))
so that every kernel variant and every launch geometry draws the same numbers.  `k2_draws` materialises that stream as the
dense per-env arrays the oracle's parity mode consumes: feeding them to `oracle/bbc_env.post_physics_step` predicts the
production-mode kernel bit for bit.

Pinned by the published known-answer vectors of Philox4x32-10 (Random123 kat_vectors) in tests/test_philox_oracle.py.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
SITE_RS0, SITE_RT0, SITE_PUSH, SITE_MOCAP, SITE_NOISE0 = 8, 10, 12, 13, 16
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of uint32 (broadcast); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint64(int(k0) & 0xFFFFFFFF), np.uint64(int(k1) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + np.uint64(W0)) & MASK, (k1 + np.uint64(W1)) & MASK
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def u32_to_unit_f32(x):
    """24 random mantissa bits -> fp32 in [0,1) (qa_common.cuh)."""
    return ((x & np.uint32(0x00FFFFFF)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def u64_to_unit_f64(a, b):
    """53 random bits -> fp64 in [0,1), numpy-style (qa_common.cuh)."""
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def _site(env, site, step, seed):
    return philox4x32_10(env, np.uint32(site), np.uint32(step & 0xFFFFFFFF), np.uint32((step >> 32) & 0xFFFFFFFF),
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def pick_mode(prior_cdf_f32, u):
    """k = #{i < DIM_C-1 : u >= cdf[i]} (qa_k2_common.cuh pick_mode), cdf in fp32 as the kernel's constant block holds it."""
    cdf = np.asarray(prior_cdf_f32, dtype=np.float32)
    return (u[:, None] >= cdf[None, :-1]).sum(axis=1).astype(np.int32)


def prior_cdf(prior, temperature):
    """ops.bbc_const: softmax(prior / T) accumulated in python doubles, stored as fp32."""
    z = [float(p) / temperature for p in prior]
    mx = max(z)
    ez = [np.exp(v - mx) for v in z]
    tot, acc, out = sum(ez), 0.0, []
    for e in ez:
        acc += e / tot
        out.append(np.float32(acc))
    return np.asarray(out, dtype=np.float32)


def k2_draws(num_envs, obs_width, noise_lanes, prior_cdf_f32, mode_offset, mode_clips, mode_cdf, seed, step):
    """The dense parity-mode draws equivalent to the production stream of K2 at (seed, step):
    noise_u (N,W) f32 [only `noise_lanes` are drawn, the rest is 0.5 = zero noise], rs_/rt_ eps_u f64, c_idx i32, cmd_u (N,5)
    f32, push_u (N,2) f32, mocap_clip_idx (N,) i32 (picked within the mode drawn at the RESET site), mocap_time_u f64."""
    e = np.arange(num_envs, dtype=np.uint32)
    d = {}
    noise = np.full((num_envs, obs_width), 0.5, dtype=np.float32)
    for i in noise_lanes:
        v = _site(e, SITE_NOISE0 + (int(i) >> 2), step, seed)
        noise[:, int(i)] = u32_to_unit_f32(v[int(i) & 3])
    d["noise_u"] = noise
    for tag, site in (("rs", SITE_RS0), ("rt", SITE_RT0)):
        r0, r1 = _site(e, site, step, seed), _site(e, site + 1, step, seed)
        d[f"{tag}_eps_u"] = u64_to_unit_f64(r0[0], r0[1])
        d[f"{tag}_c_idx"] = pick_mode(prior_cdf_f32, u32_to_unit_f32(r0[2]))
        d[f"{tag}_cmd_u"] = np.stack([u32_to_unit_f32(r0[3])] + [u32_to_unit_f32(r1[k]) for k in range(4)], axis=1)
    p = _site(e, SITE_PUSH, step, seed)
    d["push_u"] = np.stack([u32_to_unit_f32(p[0]), u32_to_unit_f32(p[1])], axis=1)
    m = _site(e, SITE_MOCAP, step, seed)
    cu, d["mocap_time_u"] = u64_to_unit_f64(m[0], m[1]), u64_to_unit_f64(m[2], m[3])
    off, clips, cdf = (np.asarray(x) for x in (mode_offset, mode_clips, mode_cdf))
    clip = np.zeros(num_envs, dtype=np.int32)
    for n in range(num_envs):                       # j = lo; while (j < hi - 1 && cdf[j] <= cu) ++j   (K2 reset path)
        lo, hi = int(off[d["rt_c_idx"][n]]), int(off[d["rt_c_idx"][n] + 1])
        j = lo
        while j < hi - 1 and cdf[j] <= cu[n]:
            j += 1
        clip[n] = clips[j]
    d["mocap_clip_idx"] = clip
    return d


SITE_DEPTH_ENV, SITE_DEPTH_PIX = 32, 64


def depth_draws(num_envs, out_h, out_w, seed, step):
    """The production stream of K14 `qa_depth_update` (csrc/qa_depth.cu) at (seed, step) as the parity draws of
    `oracle/tsc_depth.update_depth_buffer`: noise_scale_u (N,), offset_u (N,), pixel_u (N,H,W) -- pixel p of an env comes from
    word p & 3 of the call at site SITE_DEPTH_PIX + (p >> 2)."""
    e = np.arange(num_envs, dtype=np.uint32)
    r = _site(e, SITE_DEPTH_ENV, step, seed)
    P = out_h * out_w
    groups = (P + 3) >> 2
    pix = np.zeros((num_envs, groups * 4), dtype=np.float32)
    g = np.arange(groups, dtype=np.uint32)
    v = philox4x32_10(e[:, None], (np.uint32(SITE_DEPTH_PIX) + g)[None, :], np.uint32(step & 0xFFFFFFFF),
                      np.uint32((step >> 32) & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    for k in range(4):
        pix[:, k::4] = u32_to_unit_f32(v[k])
    return dict(noise_scale_u=u32_to_unit_f32(r[0]), offset_u=u32_to_unit_f32(r[1]), pixel_u=pix[:, :P].reshape(num_envs, out_h, out_w))


SITE_TSC_RESET = 40


def tsc_draws(num_envs, seed, step):
    """The reset randomisation stream of K16 `qa_post_physics_tsc_pre` (csrc/qa_tsc_env.cu) at (seed, step) as the parity
    draws {yaw_u, x_u, y_u} (N,) f32 of `oracle/tsc_env.post_physics_pre`."""
    r = _site(np.arange(num_envs, dtype=np.uint32), SITE_TSC_RESET, step, seed)
    return dict(yaw_u=u32_to_unit_f32(r[0]), x_u=u32_to_unit_f32(r[1]), y_u=u32_to_unit_f32(r[2]))

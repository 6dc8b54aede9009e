"""The numpy restatement of the kernels' production RNG (oracle/philox.py): Philox4x32-10 against the published known-answer
vectors (Random123 `kat_vectors`, philox4x32 10 rounds), and the unit-interval conversions."""
import numpy as np

import philox as P


def test_philox4x32_10_known_answer_vectors():
    kat = [((0x00000000,) * 4, (0x00000000,) * 2, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = P.philox4x32_10(*(np.array([c], dtype=np.uint32) for c in ctr), *key)
        assert tuple(int(g[0]) for g in got) == want, (ctr, [hex(int(g[0])) for g in got])
    # vectorised over the first counter word, as the kernels use it (word 0 = env id)
    e = np.arange(5, dtype=np.uint32)
    v = P.philox4x32_10(e, np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert int(v[0][0]) == 0x6627e8d5 and len({int(x) for x in v[0]}) == 5


def test_unit_interval_conversions():
    x = np.array([0, 1, 0x00ffffff, 0xffffffff, 0x12345678], dtype=np.uint32)
    u = P.u32_to_unit_f32(x)
    assert u.dtype == np.float32 and u[0] == 0 and u[2] == u[3] == np.float32(1 - 2.0 ** -24) and u[1] == np.float32(2.0 ** -24)
    d = P.u64_to_unit_f64(np.array([0xffffffff, 0], dtype=np.uint32), np.array([0xffffffff, 0], dtype=np.uint32))
    assert d[0] == 1 - 2.0 ** -53 and d[1] == 0.0
    cdf = P.prior_cdf([0.2] * 5, 0.25)
    assert cdf.dtype == np.float32 and abs(float(cdf[-1]) - 1.0) < 1e-6
    assert P.pick_mode(cdf, np.array([0.0, 0.19, 0.21, 0.999], dtype=np.float32)).tolist() == [0, 0, 1, 4]


def test_k2_draws_shapes_and_ranges():
    off, clips = [0, 2, 3, 5, 6, 7], [0, 1, 2, 3, 4, 5, 6]
    cdf = [0.4, 1.0, 1.0, 0.5, 1.0, 1.0, 1.0]
    d = P.k2_draws(64, 671, [0, 1, 2, 58], P.prior_cdf([0.2] * 5, 0.25), off, clips, cdf, seed=7, step=4)
    assert d["noise_u"].shape == (64, 671) and (d["noise_u"][:, 5] == 0.5).all() and d["noise_u"][:, 58].std() > 0.1
    assert d["rs_cmd_u"].shape == (64, 5) and d["push_u"].shape == (64, 2) and d["mocap_time_u"].dtype == np.float64
    for n in range(64):
        m = int(d["rt_c_idx"][n])
        assert 0 <= m <= 4 and off[m] <= clips.index(int(d["mocap_clip_idx"][n])) < off[m + 1]
    d2 = P.k2_draws(64, 671, [0, 1, 2, 58], P.prior_cdf([0.2] * 5, 0.25), off, clips, cdf, seed=7, step=5)
    assert not np.array_equal(d["rs_eps_u"], d2["rs_eps_u"])


def test_depth_draws_layout():
    d = P.depth_draws(3, 58, 87, seed=77, step=1)
    assert d["pixel_u"].shape == (3, 58, 87) and d["noise_scale_u"].shape == (3,) and d["pixel_u"].dtype == np.float32
    # pixel p = word (p & 3) of the call at site 64 + (p >> 2), counter word 0 = env
    p, env = 4 * 1000 + 2, 2
    v = P.philox4x32_10(np.array([env], dtype=np.uint32), np.uint32(64 + 1000), np.uint32(1), np.uint32(0), 77, 0)
    assert d["pixel_u"].reshape(3, -1)[env, p] == P.u32_to_unit_f32(v[2])[0]
    assert 0.45 < float(d["pixel_u"].mean()) < 0.55

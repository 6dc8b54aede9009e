"""GPU parity tests of the BBC env kernels, called through the C ABI (`qa_b200.ops` -> libqa_b200.so).

Three anchors:
  * tests/golden/bbc_env_*.npz -- outputs of the UNMODIFIED reference `LeggedRobot` (oracle/gen_golden.py);
  * the CPU oracle (oracle/bbc_env.py) on seeded synthetic state at N = 4096 (BASELINE config 1);
  * size-independent properties (idempotence of noise-free lanes, ranges, determinism) for the in-kernel
    Philox path, which has no dense-draw counterpart in the reference.
Bar: masks / indices / counters bit-exact, floats within 1e-5 relative (helpers.RTOL, with helpers.ATOL as
the absolute floor for cancelling sums).
"""
import pytest
import torch

import bbc_env as O
from helpers import load_env_golden, mocap_table, assert_close, to_dev
from qa_b200 import config as C
from qa_b200 import ops, synthetic
from qa_b200.config import BbcEnvConfig
from qa_b200.legged_robot import LeggedRobot, RecordedPhysics

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

STATE_KEYS = ["rew_buf", "obs_buf", "privileged_obs_buf", "obs_disc_buf", "obs_history_buf", "commands", "latent_eps",
              "latent_c", "root_states", "dof_state", "last_actions", "last_dof_vel", "last_root_vel",
              "last_torques_org", "action_history_buf", "feet_air_time", "base_lin_vel", "base_ang_vel",
              "projected_gravity", "feet_forces"]
EXACT_KEYS = ["reset_buf", "time_out_buf", "episode_length_buf", "contact_filt", "last_contacts"]


VARIANTS = ["warp", "bulk", "tiled"]      # K2 kernel variants (warp stores / TMA row stores / 8-env TMA tiles)


def make_env(cfg, static, snap, draws, counter_before, bulk="tiled", table=None):
    table = table or mocap_table()
    s = to_dev({k: v.clone() for k, v in snap.items()}, DEV)
    phys = RecordedPhysics([s])
    variant = {False: "warp", True: "bulk"}.get(bulk, bulk)
    env = LeggedRobot(cfg, phys, static, table, device=DEV, seed=7, bulk_store=(variant != "warp"),
                      tiled=(variant == "tiled"))
    env.load_state(s)
    env.common_step_counter = counter_before
    if draws is not None:
        d = dict(draws)
        if "mocap_clip_idx" not in d:
            d["mocap_clip_idx"] = table.sample_clip(d["rt_c_idx"], d["mocap_clip_u"])
        env.set_parity_draws({k: d[k] for k in ("noise_u", "rs_eps_u", "rs_c_idx", "rs_cmd_u", "rt_eps_u", "rt_c_idx",
                                                "rt_cmd_u", "push_u", "mocap_clip_idx", "mocap_time_u")})
    return env


def check_against(env, want, n_reset_ids, terminal, label):
    for k in EXACT_KEYS:
        assert_close(f"{label}.{k}", getattr(env, k), want[k])
    for k in STATE_KEYS:
        assert_close(f"{label}.{k}", getattr(env, k), want[k])
    assert_close(f"{label}.roll", env.roll, want["roll"])
    assert_close(f"{label}.pitch", env.pitch, want["pitch"])
    assert_close(f"{label}.yaw", env.yaw, want["yaw"])
    es = torch.stack([env.episode_sums[k] for k in C.REWARD_NAMES])
    assert_close(f"{label}.episode_sums", es, want["episode_sums"])
    ids = want["reset_env_ids"]
    k = int(env._reset_count.item())
    assert k == ids.numel(), f"{label}: reset count {k} != {ids.numel()}"
    assert torch.equal(env._reset_ids[:k].cpu(), ids.cpu()), f"{label}: reset_env_ids"
    assert torch.equal(env._reset_ids_i32[:k].cpu().long(), ids.cpu())
    assert_close(f"{label}.terminal_disc_states", env._terminal_disc[:k], terminal)
    assert int(env._num_resets.item()) == k
    if k and want.get("episode_rew_means") is not None:
        assert_close(f"{label}.episode_rew_means", env._episode_rew_means, want["episode_rew_means"], atol=1e-7)
        assert_close(f"{label}.time_outs", env._time_outs_latched, want["time_out_buf"])


@pytest.mark.parametrize("name", ["n64a", "n64b_push"])
@pytest.mark.parametrize("bulk", VARIANTS)
def test_post_physics_matches_reference_golden(name, bulk):
    cfg, static, snap, draws, ref, meta = load_env_golden(name)
    env = make_env(cfg, static, snap, draws, int(meta["counter_before"]), bulk=bulk)
    env.post_physics_step()
    torch.cuda.synchronize()
    check_against(env, ref, None, ref["terminal_disc_states"], name)


@pytest.mark.parametrize("name", ["n64a", "n64b_push"])
def test_small_kernels_match_reference_golden(name):
    cfg, static, snap, draws, ref, meta = load_env_golden(name)
    st, s = to_dev(static, DEV), to_dev(snap, DEV)
    # K1 PD torques (legged_robot.py:547-579)
    tq, tq_org = torch.empty(cfg.num_envs, 12, device=DEV), torch.empty(cfg.num_envs, 12, device=DEV)
    ops.pd_torques(s["actions"], s["dof_state"], st["motor_strength"], st["p_gains"], st["d_gains"],
                   st["default_dof_pos"].contiguous(), st["torque_limits"], tq, tq_org, cfg.action_scale,
                   cfg.hip_scale_reduction)
    assert_close("torques", tq, ref["torques"])
    assert_close("torques_org", tq_org, ref["torques_org"])
    # K0 action push (legged_robot.py:84-98), delay = 1
    hist = s["action_history_buf"].clone()
    out = torch.empty(cfg.num_envs, 12, device=DEV)
    ops.action_push(s["actions"], hist, out, 1, cfg.clip_actions / cfg.action_scale)
    assert torch.equal(hist.cpu(), ref["act_hist_pushed"]) and torch.equal(out.cpu(), ref["actions_clipped"])
    # K3 full height scan (legged_robot.py:1190-1228) on the PRE-step root states
    mh = torch.empty(cfg.num_envs, cfg.num_height_points, device=DEV)
    ops.height_scan(s["root_states"], st["height_points"], st["height_samples"], cfg.border_size,
                    cfg.horizontal_scale, cfg.vertical_scale, mh)
    # every height is an int16 cell value * 0.005, so a mismatch can only be a different terrain cell: index work, bit-exact.
    # (K3 is compiled with -fmad=false and follows the op order of quat_apply_yaw / the .long() quantisation, :1190-1228;
    # a numpy fp32 emulation of the kernel's arithmetic reproduces torch-CPU's cell indices on all 765 952 points of a
    # 4096-env grid, tools/k3_probe.py)
    want = ref["measured_heights"]
    mism = (mh.cpu() != want)
    assert not bool(mism.any()), f"height scan: {int(mism.sum())} of {mism.numel()} cells differ"


def test_height_scan_is_bit_exact_against_the_oracle_at_4096():
    """K3 on all 187 points of 4096 envs (765 952 quantised terrain lookups) against oracle/bbc_env.get_heights."""
    cfg = BbcEnvConfig(num_envs=4096)
    st = synthetic.make_static(cfg, seed=5)
    snap = synthetic.make_snapshot(cfg, seed=5, step=0)
    want = O.get_heights(cfg, st, snap["root_states"])
    mh = torch.empty(cfg.num_envs, cfg.num_height_points, device=DEV)
    ops.height_scan(snap["root_states"].to(DEV), st["height_points"].to(DEV), st["height_samples"].to(DEV), cfg.border_size,
                    cfg.horizontal_scale, cfg.vertical_scale, mh)
    mism = (mh.cpu() != want)
    assert not bool(mism.any()), f"height scan: {int(mism.sum())} of {mism.numel()} cells differ"
    assert len(torch.unique(want)) > 10


def test_mocap_blend_matches_oracle():
    table = mocap_table()
    n = 4096
    g = torch.Generator().manual_seed(3)
    mode = torch.randint(0, C.DIM_C, (n,), generator=g, dtype=torch.int32)
    clip = table.sample_clip(mode, torch.rand(n, generator=g, dtype=torch.float64))
    tu = torch.rand(n, generator=g, dtype=torch.float64)
    tu[:8] = torch.tensor([0.0, 1.0 - 1e-16, 0.5, 1e-12, 0.25, 0.75, 0.999999, 1e-9], dtype=torch.float64)
    want = O.mocap_frames_dense(table, clip, tu, 0.02, 2)
    out = torch.empty(n, 49, device=DEV)
    ops.mocap_blend(table.to(DEV), clip.to(DEV), tu.to(DEV), 0.02, 2, out)
    assert_close("mocap frames", out, want)
    ops.mocap_blend(table.to(DEV), clip[:0].to(DEV), tu[:0].to(DEV), 0.02, 2, out[:0])      # empty input


@pytest.mark.parametrize("bulk", VARIANTS)
def test_post_physics_matches_oracle_at_4096(bulk):
    cfg = BbcEnvConfig(num_envs=4096)
    static = synthetic.make_static(cfg, seed=1234)
    snap = synthetic.make_snapshot(cfg, seed=1234, step=0)
    draws = synthetic.make_rng_draws(cfg, seed=1234, step=0)
    table = mocap_table()
    draws["mocap_clip_idx"] = table.sample_clip(draws["rt_c_idx"], draws["mocap_clip_u"])
    torch.set_num_threads(8)
    for counter_before in (3, 399):                      # without and with push (push_interval = 400)
        want = O.post_physics_step(cfg, static, snap, draws, table, counter_before + 1)
        env = make_env(cfg, static, snap, draws, counter_before, bulk=bulk, table=table)
        env.post_physics_step()
        torch.cuda.synchronize()
        check_against(env, want, None, want["terminal_disc_states"], f"n4096[{counter_before}]")
        assert int(want["reset_buf"].sum()) > 20


def test_bulk_store_variant_is_bit_identical():
    cfg = BbcEnvConfig(num_envs=4096)
    static = synthetic.make_static(cfg, seed=5)
    snap = synthetic.make_snapshot(cfg, seed=5, step=0)
    draws = synthetic.make_rng_draws(cfg, seed=5, step=0)
    outs = []
    for bulk in VARIANTS:
        env = make_env(cfg, static, snap, draws, 10, bulk=bulk)
        env.post_physics_step()
        torch.cuda.synchronize()
        outs.append((env.obs_buf.clone(), env.privileged_obs_buf.clone(), env.obs_history_buf.clone()))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert torch.equal(a, b)


def test_ragged_env_counts():
    """N not a multiple of the CTA tile (4) nor of the warp size: the tail CTA is partial."""
    for n in (1, 3, 61):
        cfg = BbcEnvConfig(num_envs=n)
        static = synthetic.make_static(cfg, seed=21, terrain_cells=1600)
        snap = synthetic.make_snapshot(cfg, seed=21, step=0, reset_frac=0.3, plant_frac=0.05)
        draws = synthetic.make_rng_draws(cfg, seed=21, step=0)
        table = mocap_table()
        draws["mocap_clip_idx"] = table.sample_clip(draws["rt_c_idx"], draws["mocap_clip_u"])
        want = O.post_physics_step(cfg, static, snap, draws, table, 1)
        env = make_env(cfg, static, snap, draws, 0, table=table)
        env.post_physics_step()
        torch.cuda.synchronize()
        check_against(env, want, None, want["terminal_disc_states"], f"n{n}")


def test_philox_mode_properties():
    """In-kernel RNG (production mode): everything that does not depend on a draw equals the oracle;
    what does is inside its range, reproducible for (seed, step) and different across steps."""
    cfg = BbcEnvConfig(num_envs=4096)
    static = synthetic.make_static(cfg, seed=9)
    snap = synthetic.make_snapshot(cfg, seed=9, step=0)
    draws = synthetic.make_rng_draws(cfg, seed=9, step=0)
    table = mocap_table()
    draws["mocap_clip_idx"] = table.sample_clip(draws["rt_c_idx"], draws["mocap_clip_u"])
    want = O.post_physics_step(cfg, static, snap, draws, table, 4)

    def run(step_counter):
        env = make_env(cfg, static, snap, None, step_counter, table=table)
        env.post_physics_step()
        torch.cuda.synchronize()
        return env

    e1, e2, e3 = run(3), run(3), run(4)
    # the Philox stream is a function of (seed, step, env, site) only: every kernel variant draws the same numbers
    for variant in ("warp", "bulk"):
        ev = make_env(cfg, static, snap, None, 3, bulk=variant, table=table)
        ev.post_physics_step()
        torch.cuda.synchronize()
        for k in ("obs_buf", "commands", "latent_eps", "latent_c", "root_states", "dof_state", "obs_history_buf"):
            assert torch.equal(getattr(ev, k), getattr(e1, k)), (variant, k)
    assert torch.equal(e1.obs_buf, e2.obs_buf) and torch.equal(e1.commands, e2.commands)
    assert not torch.equal(e1.obs_buf, e3.obs_buf)
    for k in EXACT_KEYS:
        assert_close(k, getattr(e1, k), want[k])
    keep = ~(want["reset_buf"] | (want["episode_length_buf"] % cfg.resample_period == 0))
    # rewards use the pre-reset state and no noise, but the commands of envs resampled in the callback differ
    assert_close("rew_buf", e1.rew_buf.cpu()[keep], want["rew_buf"][keep])
    ns = static["noise_scale_vec"]
    quiet = (ns == 0)
    got, ref_obs = e1.obs_buf.cpu(), want["obs_buf"]
    # noise-free lanes of envs that neither reset nor resampled are deterministic
    assert_close("quiet lanes", got[keep][:, quiet], ref_obs[keep][:, quiet])
    # noisy lanes: |obs - noise_free| <= scale, and the noise is roughly uniform
    noise_free = ref_obs - (2 * draws["noise_u"] - 1) * ns
    dev_ = (got - noise_free)[keep][:, ~quiet] / ns[~quiet]
    assert float(dev_.abs().max()) <= 1.0 + 1e-4
    assert abs(float(dev_.mean())) < 0.01 and abs(float(dev_.std()) - 3 ** -0.5) < 0.01
    # resampled commands are inside the configured ranges and consistent with the drawn mode
    cmd, lc = e1.commands.cpu(), e1.latent_c.cpu()
    assert bool((lc.sum(dim=1) == 1).all())
    m = lc.argmax(dim=1)
    lo = torch.tensor(cfg.lin_vel_x)[m, 0]
    hi = torch.tensor(cfg.lin_vel_x)[m, 1]
    assert bool(((cmd[:, 0] == 0) | ((cmd[:, 0] >= lo - 1e-6) & (cmd[:, 0] <= hi + 1e-6))).all())
    assert bool((e1.latent_eps.cpu().abs() <= 1).all())
    # reset envs carry a mocap pose: joint angles inside the table's range, unit-ish quaternion
    r = want["reset_buf"]
    q = e1.root_states.cpu()[r][:, 3:7]
    assert bool(((q.norm(dim=-1) > 0.99) & (q.norm(dim=-1) < 1.001)).all())
    modes_hit = lc[r].argmax(dim=1).unique().numel()
    assert modes_hit >= 4


def test_multi_step_rollout_matches_oracle_chain():
    """3 consecutive env.step() calls through the public VecEnv API == 3 chained oracle steps."""
    N, T = 256, 3
    cfg = BbcEnvConfig(num_envs=N)
    static = synthetic.make_static(cfg, seed=31)
    table = mocap_table()
    snaps = [synthetic.make_snapshot(cfg, seed=31, step=t, reset_frac=0.1, plant_frac=0.02) for t in range(T)]
    carried = {k: v.clone() for k, v in snaps[0].items()}
    phys = RecordedPhysics([to_dev({k: v.clone() for k, v in s.items()}, DEV) for s in snaps])
    env = LeggedRobot(cfg, phys, static, table, device=DEV, seed=3)
    env.load_state(to_dev(snaps[0], DEV))
    env.global_counter = 1                      # past the delay-schedule pop at 0; delay stays 0
    g = torch.Generator().manual_seed(77)
    for t in range(T):
        draws = synthetic.make_rng_draws(cfg, seed=31, step=t)
        draws["mocap_clip_idx"] = table.sample_clip(draws["rt_c_idx"], draws["mocap_clip_u"])
        env.set_parity_draws({k: v for k, v in draws.items() if k != "mocap_clip_u"})
        actions = torch.randn(N, 12, generator=g)
        # oracle: step() front half, torques on the CURRENT dof_state, then post-physics on snapshot t
        hist, act = O.action_push(cfg, carried["action_history_buf"], actions, delay=0)
        s = dict(snaps[t])
        for k in ("last_actions", "last_torques_org", "last_dof_vel", "last_root_vel", "obs_history_buf",
                  "episode_length_buf", "last_contacts", "commands", "latent_eps", "latent_c", "episode_sums",
                  "feet_air_time", "obs_disc_buf"):
            s[k] = carried[k]
        s["action_history_buf"], s["actions"] = hist, act
        cur_dof = snaps[t - 1]["dof_state"] if t > 0 else snaps[0]["dof_state"]
        if t > 0:
            cur_dof = prev_out["dof_state"]
        _, s["torques_org"] = O.compute_torques(cfg, {**static, "dof_state": cur_dof}, act.clone())
        want = O.post_physics_step(cfg, static, s, draws, table, t + 1)
        obs, priv, rew, reset, extras, ids, term = env.step(actions.to(DEV))
        assert_close(f"t{t}.obs", obs, want["obs_buf"])
        assert_close(f"t{t}.priv", priv, want["privileged_obs_buf"])
        assert_close(f"t{t}.rew", rew, want["rew_buf"])
        assert_close(f"t{t}.reset", reset, want["reset_buf"])
        assert torch.equal(ids.cpu(), want["reset_env_ids"])
        assert_close(f"t{t}.terminal", term, want["terminal_disc_states"])
        assert_close(f"t{t}.disc", env.get_disc_observations(), want["obs_disc_buf"])
        if ids.numel():
            assert_close(f"t{t}.time_outs", extras["time_outs"], want["time_out_buf"])
        for k in ("last_actions", "last_torques_org", "last_dof_vel", "last_root_vel", "obs_history_buf",
                  "episode_length_buf", "last_contacts", "commands", "latent_eps", "latent_c", "episode_sums",
                  "feet_air_time", "obs_disc_buf", "action_history_buf"):
            carried[k] = want[k]
        prev_out = want


def test_device_step_counter_matches_host_counters():
    """CUDA-graph mode: the Philox counter, push schedule and contact-ring head derived on the device from
    step_state are the ones the host path passes as scalars (5 steps across a push at counter 400)."""
    N = 512
    cfg = BbcEnvConfig(num_envs=N)
    static = synthetic.make_static(cfg, seed=41)
    table = mocap_table()
    snaps = [synthetic.make_snapshot(cfg, seed=41, step=t, reset_frac=0.1, plant_frac=0.02) for t in range(5)]
    envs = []
    for dev_counter in (False, True):
        phys = RecordedPhysics([to_dev({k: v.clone() for k, v in s.items()}, DEV) for s in snaps])
        env = LeggedRobot(cfg, phys, static, table, device=DEV, seed=5)
        env.load_state(to_dev(snaps[0], DEV))
        env.common_step_counter = 397
        env._ring_head = 396 % 100
        if dev_counter:
            env.use_device_step_counter(True)
        envs.append(env)
    pushed = False
    for t in range(5):
        for env in envs:
            env.post_physics_step()
        torch.cuda.synchronize()
        a, b = envs
        for k in ("obs_buf", "commands", "root_states", "rew_buf", "reset_buf", "obs_history_buf"):
            assert torch.equal(getattr(a, k), getattr(b, k)), (t, k)
        assert torch.equal(a._contact_ring, b._contact_ring), t
        pushed = pushed or (a.common_step_counter % 400 == 0)
    assert pushed
    envs[1].common_step_counter = -1
    envs[1].sync_step_counter()
    assert envs[1].common_step_counter == envs[0].common_step_counter == 402


def test_reset_and_following_steps_match_the_reference():
    """`env.reset()` (:67-76 = reset_idx(all envs) + a zero-action step) and the two `step()`s after it against the UNMODIFIED
    reference driven on the same state and draws (oracle/gen_golden_reset.py -> tests/golden/bbc_env_reset_n64.npz): every
    carried buffer, the latched extras and the returned observations; masks / counters bit-exact."""
    import numpy as np
    from helpers import GOLD
    z = np.load(f"{GOLD}/bbc_env_reset_n64.npz")
    grp = {}
    for k in z.files:
        g, key = k.split(".", 1)
        grp.setdefault(g, {})[key] = torch.from_numpy(z[k])
    N, seed = int(grp["meta"]["num_envs"]), int(grp["meta"]["seed"])
    cfg = BbcEnvConfig(num_envs=N)
    static = dict(grp["static"])
    static["height_samples"] = synthetic.make_static(cfg, seed=seed, terrain_cells=1600)["height_samples"]
    snap, ref = grp["snap"], grp["ref"]
    env = make_env(cfg, static, snap, None, int(grp["meta"]["counter_before"]))
    env.global_counter = int(grp["meta"]["global_counter"])
    env._delay_schedule = []
    env.reset_buf.fill_(True)
    env._episode_rew_means.fill_(123.0)

    float_keys = ["commands", "latent_eps", "latent_c", "root_states", "dof_state", "obs_buf", "privileged_obs_buf", "obs_disc_buf",
                  "obs_history_buf", "last_actions", "last_dof_vel", "last_root_vel", "last_torques_org", "action_history_buf",
                  "feet_air_time", "base_lin_vel", "base_ang_vel", "projected_gravity", "feet_forces", "rew_buf", "torques_org"]
    exact_keys = ["episode_length_buf", "contact_filt", "last_contacts", "reset_buf", "time_out_buf"]

    def check(tag):
        torch.cuda.synchronize()
        for k in exact_keys:
            assert_close(f"{tag}.{k}", getattr(env, k), ref[f"{tag}.{k}"])
        for k in float_keys:
            assert_close(f"{tag}.{k}", getattr(env, k), ref[f"{tag}.{k}"])
        assert_close(f"{tag}.episode_sums", torch.stack([env.episode_sums[k] for k in C.REWARD_NAMES]), ref[f"{tag}.episode_sums"])
        assert_close(f"{tag}.extras.time_outs", env.extras["time_outs"], ref[f"{tag}.extras_time_outs"])
        got = torch.stack([env.extras["episode"]["rew_" + k] for k in C.REWARD_NAMES])
        assert_close(f"{tag}.extras.episode", got, ref[f"{tag}.extras_episode"], atol=1e-7)
        assert_close(f"{tag}.contact_buf", env.contact_buf, ref[f"{tag}.contact_buf"])
        assert_close(f"{tag}.contact_force_buf", env.contact_force_buf, ref[f"{tag}.contact_force_buf"])

    draws = [{k: v for k, v in grp[f"draws{t}"].items() if k != "mocap_clip_u"} for t in range(3)]
    env.set_parity_draws(draws[0])
    obs, priv = env.reset()
    assert obs is env.obs_buf and priv is env.privileged_obs_buf
    check("reset")
    assert env.common_step_counter == int(grp["meta"]["counter_before"]) + 1
    for t in (1, 2):
        env.set_parity_draws(draws[t])
        out = env.step(ref[f"step{t}.actions_in"].to(DEV))
        check(f"step{t}")
        assert torch.equal(out[5].cpu(), ref[f"step{t}.reset_env_ids"])
        assert_close(f"step{t}.terminal", out[6], ref[f"step{t}.terminal_disc_states"])

"""TSC env step: oracle vs the reference-generated golden fixture (CPU), kernels K16/K17 through
`LeggedRobotTSC` vs the fixture and vs the oracle at 4096 envs (GPU).  Masks / indices / counters bit-exact, floats
within 1e-5 relative (helpers.RTOL)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import tsc_env as OE  # noqa: E402
from helpers import GOLD, assert_close  # noqa: E402
from qa_b200 import synthetic  # noqa: E402

KEYS = ("obs_buf", "obs_bbc_buf", "obs_disc_buf", "rew_buf", "reset_buf", "time_out_buf", "root_states", "dof_state",
        "obst_dof_state", "episode_length_buf", "cur_goal_idx", "cur_goals", "next_goals", "reach_goal_timer",
        "measured_heights", "obs_history_buf", "contact_buf", "action_history_buf", "last_actions", "last_dof_vel",
        "last_torques_org", "last_root_vel", "last_contacts", "contact_filt", "feet_air_time", "delta_yaw", "delta_next_yaw",
        "base_lin_vel", "base_ang_vel", "projected_gravity", "roll", "pitch", "yaw", "target_yaw", "next_target_yaw",
        "cur_obstacle_types", "feet_at_edge", "episode_sums")


def oracle_step(N, seed):
    st = synthetic.make_tsc_static(N, seed)
    sn = synthetic.make_tsc_snapshot(N, st, seed)
    dr = synthetic.make_tsc_draws(N, seed)
    cfg = OE.TscCfg(num_envs=N)
    pre = OE.post_physics_pre(cfg, st, sn, dr)
    return st, sn, dr, OE.post_physics_post(cfg, st, pre, sn["rigid_body_state_post"])


def test_oracle_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "tsc_env_n64.npz"))
    st, sn, dr, out = oracle_step(int(z["num_envs"]), int(z["seed"]))
    n = 0
    for k in z.files:
        if not k.startswith("ref."):
            continue
        want, got = torch.from_numpy(z[k]), out[k[4:]]
        if want.dtype in (torch.bool, torch.int64):
            assert torch.equal(got.to(torch.int64), want.to(torch.int64)), k
        else:
            assert torch.equal(got, want), k
        n += 1
    assert n >= 40 and int(out["reset_buf"].sum()) >= 6


def run_kernels(N, seed, dev="cuda:0", parity=True, hl=True):
    from qa_b200.legged_robot_tsc import LeggedRobotTSC, RecordedPhysicsTSC, TscEnvConfig
    st = synthetic.make_tsc_static(N, seed)
    sn = synthetic.make_tsc_snapshot(N, st, seed)
    dr = synthetic.make_tsc_draws(N, seed)
    snap_dev = {k: v.to(dev).contiguous() for k, v in sn.items() if isinstance(v, torch.Tensor)}
    env = LeggedRobotTSC(TscEnvConfig(num_envs=N), RecordedPhysicsTSC([snap_dev]), st, device=dev, seed=seed)
    env.load_state(sn)
    env.physics.cursor = -1
    env.action_hl_history_buf = snap_dev["action_hl_history_buf"] if hl else None
    if parity:
        env.set_parity_draws(dr)
    ids, terminal = env.post_physics_step()
    return env, ids, terminal, sn


def collect(env):
    ph = env.physics
    out = {k: getattr(env, k) for k in KEYS if k not in ("root_states", "dof_state", "obst_dof_state", "episode_sums",
                                                         "roll", "pitch", "yaw")}
    out.update(root_states=ph.root_states, dof_state=ph.dof_state, obst_dof_state=ph.obst_dof_state,
               episode_sums=env.episode_sums_buf, roll=env.roll, pitch=env.pitch, yaw=env.yaw)
    return out


@pytest.mark.gpu
def test_kernels_match_reference_golden_n64():
    z = np.load(os.path.join(GOLD, "tsc_env_n64.npz"))
    env, ids, terminal, sn = run_kernels(int(z["num_envs"]), int(z["seed"]))
    got = collect(env)
    for k in KEYS:
        assert_close(k, got[k], torch.from_numpy(z["ref." + k]))
    assert torch.equal(ids.cpu(), torch.from_numpy(z["ref.reset_env_ids"]))
    assert_close("terminal_disc_states", terminal, torch.from_numpy(z["ref.terminal_disc_states"]))
    assert_close("time_outs", env.extras["time_outs"], torch.from_numpy(z["ref.time_outs_latched"]))
    assert_close("reach_goal", env.extras["reach_goal"], torch.from_numpy(z["ref.reach_goal"]))
    means = torch.stack([env.extras["episode"]["rew_" + k] for k in env.cfg.reward_names])
    assert_close("episode means", means, torch.from_numpy(z["ref.episode_rew_means"]), rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("N,seed,hl", [(4096, 7, True), (100, 3, False), (33, 5, True)])
def test_kernels_match_oracle(N, seed, hl):
    st, sn, dr, want = oracle_step(N, seed) if hl else (None, None, None, None)
    if not hl:                                                       # action_hl_history_buf = None: the two rate terms are 0
        st = synthetic.make_tsc_static(N, seed)
        sn = synthetic.make_tsc_snapshot(N, st, seed)
        dr = synthetic.make_tsc_draws(N, seed)
        sn2 = dict(sn)
        sn2["action_hl_history_buf"] = None
        cfg = OE.TscCfg(num_envs=N)
        want = OE.post_physics_post(cfg, st, OE.post_physics_pre(cfg, st, sn2, dr), sn["rigid_body_state_post"])
    env, ids, terminal, _ = run_kernels(N, seed, hl=hl)
    got = collect(env)
    for k in KEYS:
        assert_close(k, got[k], want[k])
    assert torch.equal(ids.cpu(), want["reset_env_ids"])
    assert_close("terminal", terminal, want["terminal_disc_states"])


@pytest.mark.gpu
def test_philox_reset_mode_is_deterministic_and_in_range():
    a, ids_a, _, sn = run_kernels(512, 9, parity=False)
    b, ids_b, _, _ = run_kernels(512, 9, parity=False)
    assert torch.equal(ids_a, ids_b) and len(ids_a) > 10
    assert torch.equal(a.physics.root_states, b.physics.root_states)
    ra = a.physics.root_states[ids_a].cpu()
    g0 = a.env_goals[ids_a, 0, :2].cpu()
    assert float((ra[:, 0] - g0[:, 0]).max()) <= 1e-6 and float((ra[:, 0] - g0[:, 0]).min()) >= -0.2 - 1e-6
    assert float((ra[:, 1] - g0[:, 1]).abs().max()) <= 0.1 + 1e-6
    yaw = 2 * torch.atan2(ra[:, 5], ra[:, 6])
    assert float((yaw - np.pi / 2).abs().max()) <= 0.2 + 1e-5
    assert float(ra[:, 0].std()) > 0                                  # envs draw different values


def test_abi_rejects_bad_tsc_arguments():
    import ctypes
    from qa_b200 import _abi
    lib = _abi.load()
    assert lib.qa_post_physics_tsc_pre(None, None, None) == -1
    c, a = _abi.QaTscConst(), _abi.QaTscStepArgs()
    c.num_bodies = 40
    assert lib.qa_post_physics_tsc_pre(ctypes.byref(c), ctypes.byref(a), None) == -2
    assert lib.qa_post_physics_tsc_post(ctypes.byref(c), ctypes.byref(a), None) == -2


@pytest.mark.gpu
def test_tsc_runner_teacher_iteration_runs_and_rewards_match_torch_path():
    """OnPolicyRunnerTSC (tsc on_policy_runner.py:164-300, teacher): a 4-step rollout over recorded state + GAE + PPO update.
    The style reward of the first step is checked against the torch restatement (`Discriminator.predict_disc_reward`)."""
    from qa_b200.config import tsc_train_cfg
    from qa_b200.legged_robot_tsc import LeggedRobotTSC, RecordedPhysicsTSC, TscEnvConfig
    from qa_b200.rsl_rl.tsc_runner import OnPolicyRunnerTSC
    torch.backends.cuda.matmul.allow_tf32 = False
    dev, N, T = "cuda:0", 256, 4
    st = synthetic.make_tsc_static(N, 4)
    snaps = [synthetic.make_tsc_snapshot(N, st, 4, step=t) for t in range(T + 1)]
    dsn = [{k: v.to(dev).contiguous() for k, v in s.items() if isinstance(v, torch.Tensor)} for s in snaps]
    env = LeggedRobotTSC(TscEnvConfig(num_envs=N), RecordedPhysicsTSC(dsn), st, device=dev, seed=4)
    env.load_state(snaps[0])
    env.post_physics_step()                                       # reference ctor: one post_physics_step fills the obs (:106)
    cfg = tsc_train_cfg()
    cfg["runner"]["num_steps_per_env"] = T
    cfg["runner"].update(reward_i_coef=1.0, reward_us_coef=0.01, reward_ss_coef=0.2, reward_t_coef=0.2)
    cfg["algorithm"].update(num_learning_epochs=2, num_mini_batches=2)
    torch.manual_seed(0)
    r = OnPolicyRunnerTSC(env, cfg, device=dev)
    obs, obs_bbc = env.get_observations(), env.get_observations_bbc().clone()
    r._disc_hist = torch.stack([env.get_observations_disc()] * 2, dim=1)
    hist0, disc0 = r._disc_hist.clone(), env.get_observations_disc().clone()
    obs, obs_bbc2, critic, infos = r.rollout_step(obs, obs_bbc, obs, {})
    # torch restatement of the reward of that step
    dones = env.reset_buf
    with_term = torch.where(dones.unsqueeze(1), disc0, env.get_observations_disc())
    hist = torch.stack([hist0[:, 1], with_term], dim=1)
    want = r.discriminator.predict_disc_reward(env.rew_buf.unsqueeze(1), obs_bbc, hist, normalizer=r.disc_normalizer)[0]
    want = want + r.alg.gamma * (r.alg.storage.values[0].squeeze(1) * infos["time_outs"]) if "time_outs" in infos else want
    assert_close("style reward", r.alg.storage.rewards[0].squeeze(1), want.float(), rtol=1e-4, atol=2e-5)
    assert r.alg.storage.actions[0].shape == (N, 19) and float(r.alg.storage.actions[0][:, 0].max()) <= 2
    for _ in range(T - 1):
        obs, obs_bbc2, critic, infos = r.rollout_step(obs, obs_bbc2, critic, infos)
    assert r.alg.storage.step == T
    r.alg.compute_returns(critic)
    out = r.alg.update()
    assert len(out) == 7 and all(np.isfinite(v) for v in out)
    assert torch.isfinite(env.obs_buf).all() and torch.isfinite(env.obs_bbc_buf).all()

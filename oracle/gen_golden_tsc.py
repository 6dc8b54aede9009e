"""Pin the TSC oracle pieces against the UNMODIFIED reference (`/root/reference/tsc`) and write their golden vectors
(build container only).  Run in its own process: bbc/ and tsc/ fork the same package names.

  python oracle/gen_golden_tsc.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

import tsc_depth as OD  # noqa: E402
from ref_harness import import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class _FakeGym:
    """Only what update_depth_buffer touches (:181-202)."""

    def __init__(self, images):
        self.images = images

    def step_graphics(self, sim): pass
    def render_all_camera_sensors(self, sim): pass
    def start_access_image_tensors(self, sim): pass
    def end_access_image_tensors(self, sim): pass

    def get_camera_image_gpu_tensor(self, sim, env, cam, kind):
        return self.images[env]


def depth_case(ref, N, seed):
    g = torch.Generator().manual_seed(seed)
    images = -(0.1 + 6.0 * torch.rand(N, 60, 106, generator=g))          # camera depth: negative metres, some beyond far
    images[0, :5] = -float("inf")                                        # sky pixels
    ep = torch.tensor([0, 1, 2, 17, 1, 250][:N] + [5] * max(0, N - 6), dtype=torch.int64)
    buf0 = torch.randn(N, 2, 58, 87, generator=g) * 0.2
    env = ref.LeggedRobot.__new__(ref.LeggedRobot)
    depth = types.SimpleNamespace(use_camera=True, update_interval=1, near_clip=0.3, far_clip=4, depth_noise=0.05, buffer_len=2)
    env.cfg = types.SimpleNamespace(depth=depth)
    env.global_counter, env.num_envs, env.device = 5, N, "cpu"
    env.sim, env.envs, env.cam_handles = None, list(range(N)), list(range(N))
    env.gym = _FakeGym(images)
    env.episode_length_buf = ep
    env.depth_buffer = buf0.clone()
    # the reference consumes the default CPU generator per env: rand(1), rand(1), rand_like(58x87) -- replay it
    torch.manual_seed(seed + 1)
    u1, u2, up = [], [], []
    for _ in range(N):
        u1.append(torch.rand(1)[0])
        u2.append(torch.rand(1)[0])
        up.append(torch.rand(58, 87))
    u1, u2, up = torch.stack(u1), torch.stack(u2), torch.stack(up)
    torch.manual_seed(seed + 1)
    env.update_depth_buffer()                                             # the reference's own code
    want = env.depth_buffer
    got = OD.update_depth_buffer(buf0, images, ep, 0.3, 4, 0.05, u1, u2, up)
    assert torch.equal(got, want), f"depth buffer: oracle != reference (max err {(got - want).abs().max()})"
    print(f"  depth: N={N}: oracle == reference update_depth_buffer (bit-exact), init envs={int((ep <= 1).sum())}")
    return dict(images=images, ep=ep, buf0=buf0, u1=u1, u2=u2, up=up, want=want)


class _FakeSimGym:
    """The gym calls post_physics_step / reset_idx make (:231-234, :381-384, :835-900).  `simulate` is a no-op; the
    rigid-body refresh that follows it inside reset_idx swaps in the recorded post-reset snapshot."""

    def __init__(self, env, rigid_post):
        self.env, self.rigid_post, self.rb_refreshes, self.set_calls = env, rigid_post, 0, []

    def refresh_actor_root_state_tensor(self, sim): pass
    def refresh_net_contact_force_tensor(self, sim): pass
    def refresh_force_sensor_tensor(self, sim): pass
    def simulate(self, sim): pass
    def fetch_results(self, sim, flag): pass

    def refresh_rigid_body_state_tensor(self, sim):
        self.rb_refreshes += 1
        if self.rb_refreshes == 2:
            self.env.rigid_body_states.copy_(self.rigid_post)

    def set_dof_state_tensor_indexed(self, sim, t, ids, n):
        self.set_calls.append(("dof", ids.clone()))

    def set_actor_root_state_tensor_indexed(self, sim, t, ids, n):
        self.set_calls.append(("root", ids.clone()))


def env_case(ref, N, seed):
    """One full TSC post_physics_step of the UNMODIFIED reference on injected synthetic state vs oracle/tsc_env.py."""
    import importlib
    import tsc_env as OE
    from qa_b200 import synthetic
    cfgmod = importlib.import_module("legged_gym.envs.go2.go2_agility_config")
    rcfg = cfgmod.Go2AgilityCfg()
    st = synthetic.make_tsc_static(N, seed)
    sn = synthetic.make_tsc_snapshot(N, st, seed)
    dr = synthetic.make_tsc_draws(N, seed)
    cfg = OE.TscCfg(num_envs=N)
    LR = ref.LeggedRobot
    env = LR.__new__(LR)
    env.cfg, env.sim_params, env.device, env.num_envs = rcfg, types.SimpleNamespace(dt=0.005), "cpu", N
    env._parse_cfg(rcfg)
    assert abs(env.dt - cfg.dt) < 1e-12 and float(env.max_episode_length) == cfg.max_episode_length
    env._prepare_reward_function()
    assert env.reward_names == cfg.reward_names, (env.reward_names, cfg.reward_names)
    for k in cfg.reward_scales:
        assert abs(env.reward_scales[k] - cfg.reward_scales[k] * cfg.dt) < 1e-12, k
    c = lambda t: t.clone()                                                   # noqa: E731
    env.sim, env.viewer, env.enable_viewer_sync, env.debug_viz, env.extras = None, None, False, False, {}
    env.obstacle = types.SimpleNamespace(cfg=rcfg.obstacle, last_goal_repeat=2, num_goals=4, proportions=[0.2, 0.15, 0.2, 0.15, 0.2, 0.1],
                                         frame_ang=[cfg.frame_ang0] * 6, seesaw_dof_pos=cfg.seesaw_dof_pos, num_envs=N)
    env.terrain, env.custom_origins = None, True
    env.root_states, env.dof_state = c(sn["root_states"]), c(sn["dof_state"])
    env.dof_pos = env.dof_state.view(N, 12, 2)[..., 0]
    env.dof_vel = env.dof_state.view(N, 12, 2)[..., 1]
    env.num_dof = 12
    env.base_quat = env.root_states[:, 3:7]
    env.rigid_body_states = c(sn["rigid_body_state"])
    env.rigid_body_pos = env.rigid_body_states[..., 0:3]
    env.contact_forces = c(sn["contact_forces"])
    env.obst_dof_state = c(sn["obst_dof_state"])
    env.obst_dof_pos, env.obst_dof_vel = env.obst_dof_state[:, 0:1], env.obst_dof_state[:, 1:2]
    env.obst_root_states, env.border_root_states = torch.zeros(6 * N, 13), torch.zeros(N, 13)
    for k in ("feet_indices", "key_body_ids", "penalised_contact_indices", "termination_contact_indices", "height_samples",
              "x_edge_mask", "height_points", "env_goals", "obstacle_types", "gravity_vec", "motor_strength"):
        setattr(env, k, c(st[k]))
    env.num_height_points = st["height_points"].shape[1]
    env.default_dof_pos, env.default_dof_pos_all = c(st["default_dof_pos"]), st["default_dof_pos"].repeat(N, 1)
    env.mass_params_tensor, env.friction_coeffs_tensor = c(st["mass_params"]), c(st["friction_coeffs"])
    env.base_init_state = torch.tensor(cfg.base_init_state)
    for k in ("episode_length_buf", "last_root_vel", "last_contacts", "reach_goal_timer", "cur_goal_idx", "cur_goals",
              "next_goals", "actions", "last_actions", "torques_org", "last_torques_org", "last_dof_vel", "measured_heights",
              "delta_yaw", "delta_next_yaw", "commands", "latent_eps", "latent_c", "obs_history_buf", "contact_buf",
              "action_history_buf", "action_hl_history_buf", "feet_air_time", "obs_disc_buf"):
        setattr(env, k, c(sn[k]))
    env.common_step_counter, env.global_counter = sn["common_step_counter"], sn["global_counter"]
    env.episode_sums = {name: c(sn["episode_sums"][:, i]) for i, name in enumerate(cfg.reward_names + ["termination"])}
    env.base_lin_vel, env.base_ang_vel, env.projected_gravity = torch.zeros(N, 3), torch.zeros(N, 3), torch.zeros(N, 3)
    env.rew_buf, env.reset_buf = torch.zeros(N), torch.ones(N, dtype=torch.long)
    env.time_out_buf = torch.zeros(N, dtype=torch.bool)
    env.gym = _FakeSimGym(env, sn["rigid_body_state_post"])
    # random sources of _reset_root_states (:860-880): three torch_rand_float calls, in order yaw, x, y
    lrmod = ref.legged_robot
    calls = []
    real = lrmod.torch_rand_float

    def fake_rand_float(lower, upper, shape, device):
        key = ("yaw_u", "x_u", "y_u")[len(calls)]
        calls.append(key)
        ids = env.reset_buf.nonzero(as_tuple=False).flatten()
        assert shape == (len(ids), 1)
        return ((upper - lower) * dr[key][ids] + lower).view(-1, 1)

    lrmod.torch_rand_float = fake_rand_float
    try:
        env_ids, term = env.post_physics_step()                               # the reference's own code
    finally:
        lrmod.torch_rand_float = real
    assert calls == ["yaw_u", "x_u", "y_u"], calls
    assert env.gym.rb_refreshes == 2 and [k for k, _ in env.gym.set_calls] == ["dof", "root"]
    pre = OE.post_physics_pre(cfg, st, sn, dr)
    out = OE.post_physics_post(cfg, st, pre, sn["rigid_body_state_post"])
    ref_out = dict(obs_buf=env.obs_buf, obs_bbc_buf=env.obs_bbc_buf, obs_disc_buf=env.obs_disc_buf, rew_buf=env.rew_buf,
                   reset_buf=env.reset_buf, time_out_buf=env.time_out_buf, reset_env_ids=env_ids, terminal_disc_states=term,
                   root_states=env.root_states, dof_state=env.dof_state, obst_dof_state=env.obst_dof_state,
                   episode_length_buf=env.episode_length_buf, cur_goal_idx=env.cur_goal_idx, cur_goals=env.cur_goals,
                   next_goals=env.next_goals, reach_goal_timer=env.reach_goal_timer, measured_heights=env.measured_heights,
                   obs_history_buf=env.obs_history_buf, contact_buf=env.contact_buf, action_history_buf=env.action_history_buf,
                   last_actions=env.last_actions, last_dof_vel=env.last_dof_vel, last_torques_org=env.last_torques_org,
                   last_root_vel=env.last_root_vel, last_contacts=env.last_contacts, contact_filt=env.contact_filt,
                   feet_air_time=env.feet_air_time, delta_yaw=env.delta_yaw, delta_next_yaw=env.delta_next_yaw,
                   base_lin_vel=env.base_lin_vel, base_ang_vel=env.base_ang_vel, projected_gravity=env.projected_gravity,
                   roll=env.roll, pitch=env.pitch, yaw=env.yaw, target_yaw=env.target_yaw, next_target_yaw=env.next_target_yaw,
                   cur_obstacle_types=env.cur_obstacle_types, reach_goal=env.extras["reach_goal"],
                   feet_at_edge=env.feet_at_edge, time_outs_latched=env.extras["time_outs"],
                   episode_sums=torch.stack([env.episode_sums[k] for k in cfg.reward_names + ["termination"]], dim=1),
                   episode_rew_means=torch.stack([env.extras["episode"]["rew_" + k] for k in cfg.reward_names + ["termination"]]))
    worst = 0.0
    for k, rv in ref_out.items():
        ov = out[k]
        assert ov.shape == rv.shape, (k, ov.shape, rv.shape)
        if rv.dtype in (torch.bool, torch.int64, torch.int32, torch.uint8, torch.int16):
            assert torch.equal(ov.to(torch.int64), rv.to(torch.int64)), f"{k}: oracle != reference"
        else:
            assert torch.equal(ov, rv), f"{k}: oracle != reference (max err {(ov - rv).abs().max()})"
    n_reset = len(env_ids)
    print(f"  env: N={N}: oracle == reference post_physics_step on {len(ref_out)} buffers (bit-exact); resets={n_reset}, "
          f"time_outs={int(env.time_out_buf.sum())}, reached={int(env.reached_goal_ids.sum())}, leave={int(env.leave_goal_ids.sum())}")
    assert n_reset >= 6
    return st, sn, dr, ref_out


STRIDE = 97


def sample_params(sd):
    return torch.cat([v.reshape(-1) for v in sd.values()])[::STRIDE].clone()


def trainer_case(ref, seed=3, N=64):
    """ActorCriticTSC forward + one full PPO.update() over a 1 x N storage (one minibatch, one epoch) with the
    reference's own classes, against oracle/tsc_trainer.py."""
    import tsc_trainer as OT
    from qa_b200 import synthetic
    torch.set_num_threads(1)
    w = synthetic.make_tsc_weights(seed)
    policy = dict(scan_encoder_dims=[128, 64, 32], actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128],
                  priv_encoder_dims=[64], activation="elu", init_noise_std=1.0, tanh_encoder_output=False)
    ac = ref.actor_critic.ActorCriticTSC(65, 8, 132, 800, 29, 4, 10, 3, 6, device="cpu", **policy)
    ac.load_state_dict(w["ac"])
    est = ref.estimator.Estimator(input_dim=57, output_dim=4, hidden_dims=[128, 64])
    est.load_state_dict(w["est"])
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(N, 800, generator=g) * 0.5
    obs[:, 65:197] = torch.clip(obs[:, 65:197], -1, 1)
    draw, mode_u = torch.randn(N, 18, generator=g), torch.rand(N, generator=g)
    P = ref.ppo.PPO
    alg = P.__new__(P)
    alg.device, alg.actor_critic, alg.estimator = "cpu", ac, est
    alg.train_with_estimated_states, alg.num_prop, alg.num_auxiliary, alg.num_scan, alg.priv_states_dim = True, 57, 8, 132, 4
    alg.transition = ref.RolloutStorage.Transition()
    alg.optimizer = torch.optim.Adam(ac.parameters(), lr=5e-4)
    alg.estimator_optimizer = torch.optim.Adam(est.parameters(), lr=1e-4)
    alg.learning_rate, alg.desired_kl, alg.schedule = 5e-4, 0.01, "adaptive"
    alg.clip_param, alg.use_clipped_value_loss, alg.value_loss_coef, alg.entropy_coef = 0.2, True, 1.0, 0.01
    alg.max_grad_norm, alg.num_learning_epochs, alg.num_mini_batches = 1.0, 1, 1
    alg.priv_reg_coef_schedual, alg.counter = [0, 0.1, 500, 1000], 800
    alg.gamma, alg.lam = 0.99, 0.95
    alg.update_counter = lambda: None
    # ---- act (ppo.py:101-125) with the draws injected into Categorical.sample / Normal.sample -------------------
    Normal, Categorical = torch.distributions.Normal, torch.distributions.Categorical
    n_orig, c_orig = Normal.sample, Categorical.sample
    Normal.sample = lambda self, sample_shape=torch.Size(): (self.loc + self.scale * draw).detach()
    Categorical.sample = lambda self, sample_shape=torch.Size(): OT.sample_mode(self.probs, mode_u)
    out = {}
    for he in (False, True):
        with torch.no_grad():
            a = alg.act(obs.clone(), obs.clone(), None, hist_encoding=he)
        tr = alg.transition
        o = OT.act(w["ac"], w["est"], obs, obs, draw, mode_u, hist_encoding=he)
        for k, rv in (("actions", a), ("values", tr.values), ("actions_log_prob_d", tr.actions_log_prob_d),
                      ("actions_log_prob_c", tr.actions_log_prob_c), ("action_mean", tr.action_mean),
                      ("action_sigma", tr.action_sigma)):
            assert torch.allclose(o[k], rv, rtol=1e-6, atol=1e-6), ("act", he, k, float((o[k] - rv).abs().max()))
            out[f"act{int(he)}.{k}"] = rv.clone()
    # ---- one PPO.update() (ppo.py:159-262): T=1 storage, one minibatch, one epoch -----------------------------
    o0 = OT.act(w["ac"], w["est"], obs, obs, draw, mode_u, hist_encoding=False)
    batch = dict(obs=obs, critic_obs=obs, actions=o0["actions"], target_values=o0["values"],
                 advantages=torch.randn(N, 1, generator=g), returns=o0["values"] + 0.3 * torch.randn(N, 1, generator=g),
                 old_actions_log_prob_d=(o0["actions_log_prob_d"] + 0.05 * torch.randn(N, generator=g)).unsqueeze(1),
                 old_actions_log_prob_c=(o0["actions_log_prob_c"] + 0.05 * torch.randn(N, generator=g)).unsqueeze(1),
                 old_mu=o0["action_mean"] + 0.05 * torch.randn(N, 18, generator=g), old_sigma=o0["action_sigma"] * 1.05)
    st = ref.RolloutStorage(N, 1, [800], [None], [19], device="cpu")
    st.observations[0], st.actions[0], st.values[0] = batch["obs"], batch["actions"], batch["target_values"]
    st.advantages[0], st.returns[0] = batch["advantages"], batch["returns"]
    st.actions_log_prob_d[0], st.actions_log_prob_c[0] = batch["old_actions_log_prob_d"], batch["old_actions_log_prob_c"]
    st.mu[0], st.sigma[0] = batch["old_mu"], batch["old_sigma"]
    alg.storage = st
    ret = alg.update()                          # (value, surrogate, estimator, 0, 0, priv_reg, priv_reg_coef)
    Normal.sample, Categorical.sample = n_orig, c_orig
    coef = OT.tsc_priv_reg_coef(800)
    assert abs(coef - ret[6]) < 1e-12
    sd_ac = {k: v.clone().requires_grad_(True) for k, v in w["ac"].items()}
    sd_est = {k: v.clone().requires_grad_(True) for k, v in w["est"].items()}
    L = OT.ppo_losses(sd_ac, sd_est, batch, priv_reg_coef=coef)
    opt_e = torch.optim.Adam(list(sd_est.values()), lr=1e-4)
    L["estimator_loss"].backward()
    torch.nn.utils.clip_grad_norm_(list(sd_est.values()), 1.0)
    opt_e.step()
    lr_new = OT.adaptive_lr(5e-4, float(L["kl_mean"]))
    opt_a = torch.optim.Adam(list(sd_ac.values()), lr=lr_new)
    L["ppo_loss"].backward()
    torch.nn.utils.clip_grad_norm_(list(sd_ac.values()), 1.0)
    opt_a.step()
    for k, rv in (("value_loss", ret[0]), ("surrogate_loss", ret[1]), ("estimator_loss", ret[2]), ("priv_reg_loss", ret[5])):
        assert abs(float(L[k]) - rv) <= 1e-5 * abs(rv) + 1e-7, (k, float(L[k]), rv)
        out[f"ppo.{k}"] = torch.tensor(rv, dtype=torch.float32)
    assert abs(alg.learning_rate - lr_new) < 1e-12, (alg.learning_rate, lr_new)
    ref_ac = {k: v.detach() for k, v in ac.state_dict().items()}
    ref_est = {k: v.detach() for k, v in est.state_dict().items()}
    for k in ref_ac:
        d = (sd_ac[k].detach() - ref_ac[k]).abs()
        # Adam's first step moves every weight by lr * g / (|g| + eps): entries whose gradient is ~1e-8 (dead ELU
        # paths) amplify summation-order noise, so a handful of entries may differ by a fraction of lr
        assert float(d.max()) <= 2.5 * lr_new and float((d > 1e-6).float().mean()) < 2e-3, ("post-step", k, float(d.max()), float((d > 1e-6).float().mean()))
    for k in ref_est:
        assert torch.allclose(sd_est[k].detach(), ref_est[k], rtol=1e-5, atol=1e-7), ("post-step est", k)
    out["ppo.kl_mean"], out["ppo.entropy"] = L["kl_mean"].detach().clone(), L["entropy"].detach().clone()
    out["ppo.lr_new"] = torch.tensor(lr_new, dtype=torch.float64)
    out["ppo.ac_params_sampled"], out["ppo.est_params_sampled"] = sample_params(ref_ac), sample_params(ref_est)
    print(f"  trainer: ActorCriticTSC act / PPO.update: oracle == reference (kl={float(L['kl_mean']):.4f}, "
          f"lr 5e-4 -> {lr_new:.6f}, modes drawn {torch.bincount(o0['actions'][:, 0].long(), minlength=3).tolist()})")
    save = {k: v.numpy() for k, v in out.items()}
    save.update({"in.obs": obs.numpy(), "in.draw": draw.numpy(), "in.mode_u": mode_u.numpy(), "in.weights_seed": np.array(seed),
                 "in.counter": np.array(800), "in.param_stride": np.array(STRIDE)})
    for k in ("advantages", "returns", "old_actions_log_prob_d", "old_actions_log_prob_c", "old_mu", "old_sigma",
              "target_values", "actions"):
        save["in.batch." + k] = batch[k].numpy()
    return save


def main():
    ref = import_reference("tsc")
    print("reference tsc imported from", ref.root)
    d = depth_case(ref, 6, seed=11)
    np.savez_compressed(os.path.join(GOLD, "tsc_depth_n6.npz"), **{k: v.numpy() for k, v in d.items()})
    print("wrote tests/golden/tsc_depth_n6.npz")
    np.savez_compressed(os.path.join(GOLD, "tsc_trainer_seed3.npz"), **trainer_case(ref))
    print("wrote tests/golden/tsc_trainer_seed3.npz")
    env_case(ref, 512, seed=2)                                           # wider coverage, not stored
    st, sn, dr, ref_out = env_case(ref, 64, seed=1)
    # static / snapshot / draws are regenerated from the seed by the tests (qa_b200.synthetic); only the reference's
    # outputs are stored
    np.savez_compressed(os.path.join(GOLD, "tsc_env_n64.npz"), seed=np.array(1), num_envs=np.array(64),
                        **{"ref." + k: v.numpy() for k, v in ref_out.items()})
    print("wrote tests/golden/tsc_env_n64.npz")


if __name__ == "__main__":
    main()

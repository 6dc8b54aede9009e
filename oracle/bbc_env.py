"""CPU ORACLE (test infrastructure, not product code) for the BBC `LeggedRobot` hot path.

A torch restatement of the reference's per-step pipeline, written functionally over a
dict of tensors and with *dense, pre-drawn* randoms (one value per env; only the entries of
envs that resample / reset are consumed).  Each function cites the reference lines it
follows (all under /root/reference/bbc/legged_gym/envs/base/legged_robot.py unless stated).

Status of the pin: the reference ships no tests/golden vectors for this path ("parity
unpinned").  The oracle is pinned instead against OUTPUTS OF THE
REFERENCE ITSELF run in the build container: `oracle/gen_golden.py` imports the unmodified
reference classes, drives them on the same seeded state with the same draws, asserts
bit-equality with this file and commits the vectors to `tests/golden/`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference
legs may import this module.  It works on any torch device (cpu for the baseline).
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------
# isaacgym.torch_utils restated (Isaac Gym Preview 4 public definitions)
# ------------------------------------------------------------------------------------------


def normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


def quat_apply(a, b):
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, :3]
    t = xyz.cross(b, dim=-1) * 2
    return (b + a[:, 3:] * t + xyz.cross(t, dim=-1)).view(shape)


def _quat_rotate_core(q, v, sign: float):
    shape = q.shape
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a + b + c if sign > 0 else a - b + c


def quat_rotate(q, v):
    return _quat_rotate_core(q, v, 1.0)


def quat_rotate_inverse(q, v):
    return _quat_rotate_core(q, v, -1.0)


def quat_from_angle_axis(angle, axis):
    theta = (angle / 2).unsqueeze(-1)
    xyz = normalize(axis) * theta.sin()
    w = theta.cos()
    return normalize(torch.cat([xyz, w], dim=-1))


# torch_jit_utils.py:24-35, 64-75, 117-122, 169-192 -------------------------------------------

def calc_heading(q):
    ref_dir = torch.zeros_like(q[..., 0:3])
    ref_dir[..., 0] = 1
    rot_dir = quat_rotate(q, ref_dir)
    return torch.atan2(rot_dir[..., 1], rot_dir[..., 0])


def calc_heading_quat_inv(q):
    heading = calc_heading(q)
    axis = torch.zeros_like(q[..., 0:3])
    axis[..., 2] = 1
    return quat_from_angle_axis(-heading, axis)


def quat_apply_yaw(quat, vec):
    quat_yaw = quat.clone().view(-1, 4)
    quat_yaw[:, :2] = 0.
    quat_yaw = normalize(quat_yaw)
    return quat_apply(quat_yaw, vec)


def euler_from_quaternion(q):
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    t0 = +2.0 * (w * x + y * z)
    t1 = +1.0 - 2.0 * (x * x + y * y)
    roll_x = torch.atan2(t0, t1)
    t2 = +2.0 * (w * y - z * x)
    t2 = torch.clip(t2, -1, 1)
    pitch_y = torch.asin(t2)
    t3 = +2.0 * (w * z + x * y)
    t4 = +1.0 - 2.0 * (y * y + z * z)
    yaw_z = torch.atan2(t3, t4)
    return roll_x, pitch_y, yaw_z


def compute_flat_key_pos(root_states, key_body_pos):
    """legged_robot.py:1377-1396."""
    root_pos = root_states[:, 0:3]
    root_rot = root_states[:, 3:7]
    heading_rot = calc_heading_quat_inv(root_rot)
    local = key_body_pos - root_pos.unsqueeze(-2)
    n, k = local.shape[0], local.shape[1]
    hr = heading_rot.unsqueeze(-2).repeat((1, k, 1)).view(n * k, 4)
    out = quat_rotate(hr, local.reshape(n * k, 3))
    return out.view(n, k * 3)


# ------------------------------------------------------------------------------------------
# a2  _compute_torques  (:547-579)
# ------------------------------------------------------------------------------------------

def compute_torques(cfg, st: Dict[str, torch.Tensor], actions: torch.Tensor):
    """Returns (clipped torques, torques_org).  `st` needs dof_state, motor_strength, p/d gains."""
    N = actions.shape[0]
    dof = st["dof_state"].view(N, 12, 2)
    dof_pos, dof_vel = dof[..., 0], dof[..., 1]
    actions_scaled = actions * cfg.action_scale
    actions_scaled[:, [0, 3, 6, 9]] *= cfg.hip_scale_reduction
    ms = st["motor_strength"]
    torques = ms[0] * st["p_gains"] * (actions_scaled + st["default_dof_pos"] - dof_pos) \
        - ms[1] * st["d_gains"] * dof_vel
    return torch.clip(torques, -st["torque_limits"], st["torque_limits"]), torques


# ------------------------------------------------------------------------------------------
# a1  step() front half (:84-98): history push, delayed-action select, clip
# ------------------------------------------------------------------------------------------

def action_push(cfg, action_history_buf: torch.Tensor, actions: torch.Tensor, delay: int):
    hist = torch.cat([action_history_buf[:, 1:].clone(), actions[:, None, :].clone()], dim=1)
    if cfg.action_delay:
        actions = hist[:, -delay - 1]
    clip_actions = cfg.clip_actions / cfg.action_scale
    return hist, torch.clip(actions, -clip_actions, clip_actions)


# ------------------------------------------------------------------------------------------
# a5  _get_heights (:1190-1228)
# ------------------------------------------------------------------------------------------

def get_heights(cfg, st, root_states):
    N = root_states.shape[0]
    P = st["height_points"].shape[0]
    height_points = st["height_points"].unsqueeze(0).expand(N, P, 3)
    base_quat = root_states[:, 3:7]
    points = quat_apply_yaw(base_quat.repeat(1, P), height_points) + (root_states[:, :3]).unsqueeze(1)
    points = points + cfg.border_size
    points = (points / cfg.horizontal_scale).long()
    px = points[:, :, 0].view(-1)
    py = points[:, :, 1].view(-1)
    hs = st["height_samples"]
    px = torch.clip(px, 0, hs.shape[0] - 2)
    py = torch.clip(py, 0, hs.shape[1] - 2)
    h1 = hs[px, py]
    h2 = hs[px + 1, py]
    h3 = hs[px, py + 1]
    heights = torch.min(torch.min(h1, h2), h3)
    return heights.view(N, -1) * cfg.vertical_scale


# ------------------------------------------------------------------------------------------
# a6  _resample_latent_eps / _resample_latent_c / _resample_commands (:474-540), dense form
# ------------------------------------------------------------------------------------------

def resample_dense(cfg, commands, latent_eps, latent_c, mask, eps_u, c_idx, cmd_u):
    """Applies the three resamplers to rows where `mask` is set, using per-env draws.

    eps_u   (N,) f64  -> np.random.rand()*2-1 evaluated in float64, cast to f32 (:533)
    c_idx   (N,) int  -> the multinomial outcome (:539)
    cmd_u   (N,5) f32 -> torch.rand draws for vx, vy, yaw, jump_h, loco_h (:504-525)
    """
    dev = commands.device
    eps_new = (eps_u * 2. - 1.).to(torch.float32).unsqueeze(1)
    c_new = F.one_hot(c_idx.long(), num_classes=latent_c.shape[1]).to(latent_c.dtype)
    latent_eps = torch.where(mask[:, None], eps_new, latent_eps)
    latent_c = torch.where(mask[:, None], c_new, latent_c)

    idx = torch.argmax(latent_c, dim=-1)
    lx = torch.tensor(cfg.lin_vel_x, device=dev)
    ly = torch.tensor(cfg.lin_vel_y, device=dev)
    lw = torch.tensor(cfg.ang_vel_yaw, device=dev)

    def draw(rng, u):                                # torch_rand_floats, torch_jit_utils.py:111-114
        lo, hi = rng[idx, 0], rng[idx, 1]
        return (hi - lo) * u + lo

    new = torch.zeros_like(commands)
    new[:, 0] = draw(lx, cmd_u[:, 0])
    new[:, 1] = draw(ly, cmd_u[:, 1])
    new[:, 2] = draw(lw, cmd_u[:, 2])
    jump = (idx == (latent_c.shape[1] - 1))
    jh, lh = cfg.jump_height, cfg.locomotion_height
    # torch_rand_float(lo, hi, ...) with python-float bounds: (hi-lo) is evaluated in double, then
    # multiplies an f32 tensor (isaacgym.torch_utils.torch_rand_float)
    new[:, 3] = ((jh[1] - jh[0]) * cmd_u[:, 3] + jh[0]) * (jump.float())
    new[:, 4] = ((lh[1] - lh[0]) * cmd_u[:, 4] + lh[0]) * ((~jump).float())
    new[:, 0] *= (torch.abs(new[:, 0]) > cfg.lin_vel_x_clip)
    new[:, 1] *= (torch.abs(new[:, 1]) > cfg.lin_vel_y_clip)
    new[:, 2] *= (torch.abs(new[:, 2]) > cfg.ang_vel_yaw_clip)
    commands = torch.where(mask[:, None], new, commands)
    return commands, latent_eps, latent_c


# ------------------------------------------------------------------------------------------
# a10  MotionLoader.get_full_frame_at_time_batch (+ traj_time_sample_batch) and quaternion_slerp
#      bbc/rsl_rl/datasets/motion_loader.py:333-341, 410-447 ; bbc/rsl_rl/utils/utils.py:126-159
# ------------------------------------------------------------------------------------------
_EPS = torch.finfo(torch.float64).eps * 4.0


def quaternion_slerp(q0, q1, fraction):
    """utils.py:126-159 with spin=0, shortestpath=True.  NB: scales by 1/angle, not 1/sin(angle)
    (:154) -- reproduced as is.  Does not mutate its arguments (the reference does)."""
    q0 = q0.clone()
    q1 = q1.clone()
    out = torch.zeros_like(q0)
    zero_mask = torch.isclose(fraction, torch.zeros_like(fraction)).squeeze(-1)
    ones_mask = torch.isclose(fraction, torch.ones_like(fraction)).squeeze(-1)
    out[zero_mask] = q0[zero_mask]
    out[ones_mask] = q1[ones_mask]
    d = torch.sum(q0 * q1, dim=-1, keepdim=True)
    dist_mask = (torch.abs(torch.abs(d) - 1.0) < _EPS).squeeze(-1)
    out[dist_mask] = q0[dist_mask]
    d_old = torch.clone(d)
    d = torch.where(d_old < 0, -d, d)
    q1 = torch.where(d_old < 0, -q1, q1)
    d = torch.clip(d, -1, 1)
    angle = torch.acos(d)
    angle_mask = (torch.abs(angle) < _EPS).squeeze(-1)
    out[angle_mask] = q0[angle_mask]
    final_mask = ~(zero_mask | ones_mask | dist_mask | angle_mask)
    isin = 1.0 / angle
    q0 = q0 * (torch.sin((1.0 - fraction) * angle) * isin)
    q1 = q1 * (torch.sin(fraction * angle) * isin)
    q0 = q0 + q1
    out[final_mask] = q0[final_mask]
    return out


def mocap_frames_dense(table, clip_idx, time_u, time_between_frames: float, disc_obs_len: int = 2):
    """Blended 49-float frame for EVERY env (callers mask).  Index math in float64 like numpy.

    clip_idx (N,) int  : sampled clip id (the np.random.choice outcome, motion_loader.py:314-320)
    time_u   (N,) f64  : np.random.uniform() draw (:336-337)
    """
    dev = table.frames.device
    ci = clip_idx.long()
    lens = table.clip_len_s[ci]
    subst = time_between_frames * disc_obs_len + table.clip_frame_dur[ci]
    t = (lens - subst) * time_u
    t = torch.maximum(torch.zeros_like(t) + 1e-7, t)
    p = t / lens
    n = table.clip_nframes[ci]
    pn = p * n
    lo = torch.floor(pn).long()
    hi = torch.ceil(pn).long()
    start = table.clip_start[ci].long()
    f0 = table.frames[start + lo]
    f1 = table.frames[start + hi]
    blend = (pn - lo.double()).to(torch.float32).unsqueeze(-1)
    pos = (1.0 - blend) * f0[:, 0:3] + blend * f1[:, 0:3]
    rot = quaternion_slerp(f0[:, 3:7], f1[:, 3:7], blend)
    traj = (1.0 - blend) * f0[:, 7:49] + blend * f1[:, 7:49]
    return torch.cat([pos, rot, traj], dim=-1).to(dev)


# ------------------------------------------------------------------------------------------
# a3/a4/a7/a8/a9/a11  post_physics_step (:124-166) and everything it calls
# ------------------------------------------------------------------------------------------

def reward_terms(cfg, st, s, base_lin_vel, base_ang_vel, root_h_pre):
    """The 14 active `_reward_*` terms in dir() order (:1231-1374).  Returns list of (N,) tensors."""
    N = s["actions"].shape[0]
    dof = s["dof_state"].view(N, 12, 2)
    dof_pos, dof_vel = dof[..., 0], dof[..., 1]
    cf = s["contact_forces"]
    cmd = s["commands"]
    dpl = st["dof_pos_limits"]
    r = {}
    r["action_rate"] = torch.sum(torch.square(s["last_actions"] - s["actions"]), dim=1)
    r["collision"] = torch.sum(1. * (torch.norm(cf[:, cfg.penalised_contact_indices, :], dim=-1) > 0.1), dim=1)
    r["delta_torques"] = torch.sum(torch.square(s["torques_org"] - s["last_torques_org"]), dim=1)
    r["dof_acc"] = torch.sum(torch.square((s["last_dof_vel"] - dof_vel) / cfg.dt), dim=1)
    r["dof_error"] = torch.sum(torch.square(dof_pos - st["default_dof_pos"]), dim=1)
    out = -(dof_pos - dpl[:, 0]).clip(max=0.)
    out = out + (dof_pos - dpl[:, 1]).clip(min=0.)
    r["dof_pos_limits"] = torch.sum(out, dim=1)
    r["dof_vel_limits"] = torch.sum(
        (torch.abs(dof_vel) - st["dof_vel_limits"] * cfg.soft_dof_vel_limit).clip(min=0., max=1.), dim=1)
    hip = cfg.hip_indices
    r["hip_pos"] = torch.sum(torch.square(dof_pos[:, hip] - st["default_dof_pos"][:, hip]), dim=1)
    # jump_up_height :1312-1322
    err_j = torch.sqrt(torch.square(cmd[:, 3] - root_h_pre))
    jump_sig = cmd[:, 3] >= cfg.jump_height[0]
    jr = torch.zeros_like(cmd[:, 3])
    jr[(err_j < 0.05) & jump_sig] += cfg.jump_goal
    r["jump_up_height"] = jr
    # locomotion_height :1324-1335
    err_l = torch.sqrt(torch.square(cmd[:, 4] - root_h_pre))
    rl = torch.exp(-10.0 * torch.square(err_l) / cfg.tracking_sigma)
    loco_sig = ~(cmd[:, 3] > cfg.jump_height[0])
    lr = torch.zeros_like(cmd[:, 4])
    lr[loco_sig] += rl[loco_sig]
    r["locomotion_height"] = lr
    r["torque_limits"] = torch.sum(
        (torch.abs(s["torques_org"]) - st["torque_limits"] * cfg.soft_torque_limit).clip(min=0.), dim=1)
    r["torques"] = torch.sum(torch.square(s["torques_org"]), dim=1)
    r["tracking_ang_vel"] = torch.exp(-torch.square(cmd[:, 2] - base_ang_vel[:, 2]) / cfg.tracking_sigma)
    r["tracking_lin_vel"] = torch.exp(
        -torch.sum(torch.square(cmd[:, :2] - base_lin_vel[:, :2]), dim=1) / cfg.tracking_sigma)
    return r


def post_physics_step(cfg, st, s, draws, table, common_step_counter: int):
    """One full `post_physics_step` on snapshot `s` (dict, NOT mutated).  Returns a dict with the
    new carried buffers and the step outputs.  `common_step_counter` is the value AFTER the
    increment at :134."""
    from qa_b200 import config as C   # constants only
    N = s["root_states"].shape[0]
    dev = s["root_states"].device
    o = {}
    root = s["root_states"].clone()
    dof_state = s["dof_state"].clone()
    rb_pos = s["rigid_body_state"].view(N, -1, 13)[..., 0:3]
    cf = s["contact_forces"]

    ep = s["episode_length_buf"] + 1                                                    # :133
    base_quat = root[:, 3:7]
    gravity_vec = torch.tensor([0., 0., -1.], device=dev).repeat(N, 1)
    base_lin_vel = quat_rotate_inverse(base_quat, root[:, 7:10])                        # :138
    base_ang_vel = quat_rotate_inverse(base_quat, root[:, 10:13])
    projected_gravity = quat_rotate_inverse(base_quat, gravity_vec)
    roll, pitch, yaw = euler_from_quaternion(base_quat)                                 # :141
    feet_forces = torch.norm(cf[:, cfg.feet_indices], dim=-1)                           # :143
    contact = feet_forces > 2.
    contact_filt = torch.logical_or(contact, s["last_contacts"])
    o["last_contacts"] = contact

    # ---- callback (:449-472) ------------------------------------------------------------
    rs_mask = (ep % cfg.resample_period == 0)
    commands, latent_eps, latent_c = resample_dense(
        cfg, s["commands"], s["latent_eps"], s["latent_c"], rs_mask,
        draws["rs_eps_u"], draws["rs_c_idx"], draws["rs_cmd_u"])
    measured_heights = get_heights(cfg, st, root)                                       # :470
    do_push = cfg.push_robots and (common_step_counter % cfg.push_interval == 0)
    if do_push:                                                                         # :682-687
        mv = cfg.max_push_vel_xy
        root[:, 7:9] = (mv - (-mv)) * draws["push_u"] + (-mv)

    # ---- termination (:168-176) ---------------------------------------------------------
    reset_buf = torch.any(torch.norm(cf[:, cfg.termination_contact_indices, :], dim=-1) > 1., dim=1)
    time_out_buf = ep > cfg.max_episode_length
    time_out_buf = time_out_buf | (root[:, 2] < -6.0)
    reset_buf = reset_buf | time_out_buf

    # ---- reward (:242-259) --------------------------------------------------------------
    ci = cfg.center_height_index
    root_h_pre = root[:, 2] - measured_heights[:, ci]
    sv = dict(s)
    sv["commands"] = commands
    sv["dof_state"] = dof_state
    terms = reward_terms(cfg, st, sv, base_lin_vel, base_ang_vel, root_h_pre)
    scales = cfg.reward_scales_dt()
    rew = torch.zeros(N, device=dev)
    episode_sums = s["episode_sums"].clone()
    for k, name in enumerate(C.REWARD_NAMES):
        t = terms[name] * scales[k]
        rew = rew + t
        episode_sums[k] = episode_sums[k] + t
    if cfg.only_positive_rewards:
        rew = torch.clip(rew, min=0.)

    # ---- reset_idx (:178-240), dense -----------------------------------------------------
    env_ids = reset_buf.nonzero(as_tuple=False).flatten()
    o["reset_env_ids"] = env_ids
    o["terminal_disc_states"] = s["obs_disc_buf"][env_ids]                              # stale, :153-154
    rmask = reset_buf
    commands, latent_eps, latent_c = resample_dense(
        cfg, commands, latent_eps, latent_c, rmask,
        draws["rt_eps_u"], draws["rt_c_idx"], draws["rt_cmd_u"])
    frames = mocap_frames_dense(table, draws["mocap_clip_idx"], draws["mocap_time_u"], cfg.dt)
    dof = dof_state.view(N, 12, 2)
    dof[..., 0] = torch.where(rmask[:, None], frames[:, 7:19], dof[..., 0])             # :607-608
    dof[..., 1] = torch.where(rmask[:, None], frames[:, 37:49], dof[..., 1])
    root_pos = frames[:, 0:3] + st["env_origins"]                                       # :668-670
    root_orn = frames[:, 3:7]
    new_root = torch.cat([root_pos, root_orn, quat_rotate(root_orn, frames[:, 31:34]),
                          quat_rotate(root_orn, frames[:, 34:37])], dim=-1)
    root = torch.where(rmask[:, None], new_root, root)
    ep = torch.where(rmask, torch.zeros_like(ep), ep)                                   # :225
    action_history_buf = torch.where(rmask[:, None, None], torch.zeros_like(s["action_history_buf"]),
                                     s["action_history_buf"])                           # :227
    obs_history_buf = torch.where(rmask[:, None, None], torch.zeros_like(s["obs_history_buf"]),
                                  s["obs_history_buf"])                                 # :228
    feet_air_time = torch.where(rmask[:, None], torch.zeros_like(s["feet_air_time"]), s["feet_air_time"])
    if env_ids.numel() > 0:                                                             # :230-234
        o["episode_rew_means"] = torch.stack(
            [torch.mean(episode_sums[k][env_ids]) / cfg.episode_length_s for k in range(len(C.REWARD_NAMES))])
        episode_sums = torch.where(rmask[None, :], torch.zeros_like(episode_sums), episode_sums)
    else:
        o["episode_rew_means"] = None

    # ---- compute_observations (:261-331) -------------------------------------------------
    dof_pos, dof_vel = dof[..., 0], dof[..., 1]
    root_h = (root[:, 2] - measured_heights[:, ci]).view(-1, 1)
    imu_obs = torch.stack((roll, pitch), dim=1)
    key_body_pos = rb_pos[:, cfg.feet_indices, :]
    flat_local_key_pos = compute_flat_key_pos(root, key_body_pos)
    dq = (dof_pos - st["default_dof_pos"]) * cfg.s_dof_pos
    obs_disc_buf = torch.cat([imu_obs, root_h, base_lin_vel * cfg.s_lin_vel_dist,
                              base_ang_vel * cfg.s_ang_vel_dist, dq, dof_vel * cfg.s_dof_vel,
                              flat_local_key_pos * cfg.s_key_pos,
                              contact_filt.float() * cfg.s_foot_contact], dim=-1)
    obs57 = torch.cat([imu_obs, base_ang_vel * cfg.s_ang_vel, dq, dof_vel * cfg.s_dof_vel,
                       action_history_buf[:, -1], contact_filt.float() - 0.5, flat_local_key_pos * 0], dim=-1)
    root_h_obs = root_h if cfg.root_height_obs else torch.zeros_like(root_h)
    priv_explicit = torch.cat([root_h_obs, base_lin_vel * cfg.s_lin_vel], dim=-1)
    ms = st["motor_strength"]
    priv_latent = torch.cat((st["mass_params_tensor"], st["friction_coeffs_tensor"], ms[0] - 1, ms[1] - 1), dim=-1)
    obs_history_buf = torch.where((ep <= 1)[:, None, None],
                                  torch.stack([obs57] * C.HISTORY_LEN, dim=1),
                                  torch.cat([obs_history_buf[:, 1:], obs57.unsqueeze(1)], dim=1))
    priv = torch.cat([obs57, priv_explicit, priv_latent, obs_history_buf.view(N, -1),
                      commands, latent_eps, latent_c], dim=-1)
    if cfg.add_noise:
        priv = priv + (2 * draws["noise_u"] - 1) * st["noise_scale_vec"]
    clip_obs = cfg.clip_observations
    o["obs_buf"] = torch.clip(torch.clone(priv), -clip_obs, clip_obs)
    o["privileged_obs_buf"] = torch.clip(priv, -clip_obs, clip_obs)
    o["obs_history_buf"] = torch.clip(obs_history_buf, -clip_obs, clip_obs)
    o["obs_disc_buf"] = obs_disc_buf

    # ---- carried buffers (:158-161) -------------------------------------------------------
    o["last_actions"] = s["actions"].clone()
    o["last_dof_vel"] = dof_vel.clone()
    o["last_root_vel"] = root[:, 7:13].clone()
    o["last_torques_org"] = s["torques_org"].clone()
    o.update(root_states=root, dof_state=dof_state, episode_length_buf=ep, commands=commands,
             latent_eps=latent_eps, latent_c=latent_c, action_history_buf=action_history_buf,
             episode_sums=episode_sums, feet_air_time=feet_air_time, rew_buf=rew, reset_buf=reset_buf,
             time_out_buf=time_out_buf, base_lin_vel=base_lin_vel, base_ang_vel=base_ang_vel,
             projected_gravity=projected_gravity, roll=roll, pitch=pitch, yaw=yaw,
             feet_forces=feet_forces, contact_filt=contact_filt, measured_heights=measured_heights,
             do_push=do_push)
    return o

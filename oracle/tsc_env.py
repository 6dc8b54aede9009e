"""CPU ORACLE (test infrastructure, not product code) for the TSC `LeggedRobot` per-step pipeline.

A torch restatement of tsc/legged_gym/envs/base/legged_robot.py `post_physics_step` (:226-298) written functionally over
a dict of tensors, split where the reference steps physics inside `reset_idx` (:381-384):

  post_physics_pre   :233-270 up to and including the simulator-state writes of reset_idx (_update_goals, heights,
                     termination, the active reward terms, _reset_dofs / _reset_root_states)
  -- the backend re-simulates the reset envs and refreshes the rigid-body tensor --
  post_physics_post  :386-285: buffer resets, episode statistics, goal gathers, compute_observations, last_* copies

Random draws are dense per-env inputs (`yaw_u`, `x_u`, `y_u` in [0,1), consumed only by envs that reset), the same
interposition as the BBC oracle's.  Configuration = the shipped go2 agility teacher config
(legged_robot_config.py + go2_agility_config.py): obstacle terrain, measure_heights, no noise, no push, no camera,
randomize_start False, action_delay, 7 reward terms + termination.

Pinned against the UNMODIFIED reference by oracle/gen_golden_tsc.py (fixture tests/golden/tsc_env_n64.npz).
Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List

import torch
import torch.nn.functional as F

from bbc_env import (compute_flat_key_pos, euler_from_quaternion, quat_apply_yaw, quat_rotate_inverse)  # noqa: F401


@dataclass
class TscCfg:
    """The scalars of LeggedRobotCfg / Go2AgilityCfg the hot path reads."""
    num_envs: int = 64
    dt: float = 0.02                                 # decimation 4 x sim dt 0.005
    max_episode_length: float = 2000.0               # ceil(40 / 0.02)
    episode_length_s: float = 40.0
    history_len: int = 10
    contact_buf_len: int = 100
    next_goal_threshold: float = 0.4
    reach_goal_delay: float = 0.02
    leave_goal_threshold: float = 4.0
    root_height_obs: bool = True
    num_goals: int = 4                               # per obstacle
    last_goal_repeat: int = 2
    num_obstacle_types: int = 6
    update_interval: int = 1                         # cfg.depth.update_interval
    use_camera: bool = False
    border_size: float = 5.0                         # obstacle.cfg.*
    horizontal_scale: float = 0.05
    vertical_scale: float = 0.005
    target_lin_vel: float = 0.4
    only_positive_rewards: bool = True
    clip_obs: float = 100.0
    # obs scales (go2 agility: the dist / key_pos / foot_contact scales are 0.0)
    s_lin_vel: float = 0.5
    s_ang_vel: float = 0.25
    s_dof_pos: float = 1.0
    s_dof_vel: float = 0.05
    s_key_pos: float = 0.0
    s_foot_contact: float = 0.0
    s_lin_vel_dist: float = 0.0
    s_ang_vel_dist: float = 0.0
    # reset randomisation (legged_robot_config.py:44-52)
    rand_yaw_range: float = 0.2
    rand_x_range: float = 0.2
    rand_y_range: float = 0.1
    frame_ang0: float = math.pi / 2                  # obstacle.frame_ang[0]
    base_init_state: List[float] = field(default_factory=lambda: [0., 0., 0.42, 0., 0., 0., 1., 0., 0., 0., 0., 0., 0.])
    seesaw_dof_pos: float = 0.3
    # reward scales BEFORE the dt multiplication (legged_robot_config.py:308-332, zero scales dropped), in the order
    # class_to_dict yields them: dir() = alphabetical (helpers.py:11-25)
    reward_scales: Dict[str, float] = field(default_factory=lambda: dict(
        action_hl_rate=-0.2, collision=-20.0, feet_edge=-1.0, latent_c_rate=-1.0, reach_goal=5.0, termination=-50.0,
        tracking_goal_vel=0.4, tracking_yaw=2.0))

    @property
    def reward_names(self):
        """compute_reward order: alphabetical, `termination` handled after the clip (:1108-1122, :423-430).  The
        episode-sum columns of this oracle / the kernels are [these 7 | termination]."""
        return [k for k in self.reward_scales if k != "termination"]


def update_goals(cfg: TscCfg, s):
    """:204-224.  Mutates reach_goal_timer / cur_goal_idx; returns the derived per-step quantities."""
    next_flag = s["reach_goal_timer"] > cfg.reach_goal_delay / cfg.dt
    s["cur_goal_idx"] = s["cur_goal_idx"] + next_flag.long()
    s["reach_goal_timer"] = torch.where(next_flag, torch.zeros_like(s["reach_goal_timer"]), s["reach_goal_timer"])
    d = torch.norm(s["root_states"][:, :2] - s["cur_goals"][:, :2], dim=1)
    reached = d < cfg.next_goal_threshold
    leave = d > cfg.leave_goal_threshold
    s["reach_goal_timer"] = s["reach_goal_timer"] + reached.float()
    target_pos_rel = s["cur_goals"][:, :2] - s["root_states"][:, :2]
    next_target_pos_rel = s["next_goals"][:, :2] - s["root_states"][:, :2]
    norm = torch.norm(target_pos_rel, dim=-1, keepdim=True)
    tv = target_pos_rel / (norm + 1e-5)
    target_yaw = torch.atan2(tv[:, 1], tv[:, 0])
    norm = torch.norm(next_target_pos_rel, dim=-1, keepdim=True)
    tv = next_target_pos_rel / (norm + 1e-5)
    next_target_yaw = torch.atan2(tv[:, 1], tv[:, 0])
    return dict(reached_goal_ids=reached, leave_goal_ids=leave, target_pos_rel=target_pos_rel, target_yaw=target_yaw,
                next_target_yaw=next_target_yaw)


def get_heights(cfg: TscCfg, static, root_states):
    """:1708-1755 (mesh_type 'obstacle')."""
    N = root_states.shape[0]
    base_quat = root_states[:, 3:7]
    P = static["height_points"].shape[1]
    points = quat_apply_yaw(base_quat.repeat(1, P), static["height_points"]) + root_states[:, :3].unsqueeze(1)
    points = points + cfg.border_size
    points = (points / cfg.horizontal_scale).long()
    hs = static["height_samples"]
    px = torch.clip(points[:, :, 0].view(-1), 0, hs.shape[0] - 2)
    py = torch.clip(points[:, :, 1].view(-1), 0, hs.shape[1] - 2)
    h = torch.min(torch.min(hs[px, py], hs[px + 1, py]), hs[px, py + 1])
    return h.view(N, -1) * cfg.vertical_scale


def post_physics_pre(cfg: TscCfg, static, state, draws):
    """Returns a dict with every buffer the reference has mutated / produced when it reaches gym.simulate in reset_idx.
    `state` is not modified."""
    s = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in state.items()}
    N = cfg.num_envs
    dof = s["dof_state"].view(N, 12, 2)
    s["episode_length_buf"] = s["episode_length_buf"] + 1                                       # :236
    s["common_step_counter"] = s["common_step_counter"] + 1
    q = s["root_states"][:, 3:7]
    s["base_lin_vel"] = quat_rotate_inverse(q, s["root_states"][:, 7:10])                        # :241-243
    s["base_ang_vel"] = quat_rotate_inverse(q, s["root_states"][:, 10:13])
    s["projected_gravity"] = quat_rotate_inverse(q, static["gravity_vec"])
    s["base_lin_acc"] = (s["root_states"][:, 7:10] - s["last_root_vel"][:, :3]) / cfg.dt         # :244
    s["roll"], s["pitch"], s["yaw"] = euler_from_quaternion(q)
    feet = static["feet_indices"]
    contact = torch.norm(s["contact_forces"][:, feet], dim=-1) > 2.                             # :247-249
    s["contact_filt"] = torch.logical_or(contact, s["last_contacts"])
    s["last_contacts"] = contact
    g = update_goals(cfg, s)                                                                    # :252
    s.update(g)
    if s["global_counter"] % cfg.update_interval == 0:                                          # :635-637
        s["measured_heights"] = get_heights(cfg, static, s["root_states"])
    G = static["env_goals"].shape[1]
    idx = torch.clamp(s["cur_goal_idx"], 0, G - cfg.last_goal_repeat - 1)                        # :255-258
    s["cur_obstacle_types"] = static["obstacle_types"].gather(1, (idx // cfg.num_goals).unsqueeze(1)).squeeze(1)
    # check_termination :322-346
    reset = torch.any(torch.norm(s["contact_forces"][:, static["termination_contact_indices"], :], dim=-1) > 1., dim=1)
    roll_cut, pitch_cut = torch.abs(s["roll"]) > 1.5, torch.abs(s["pitch"]) > 1.5
    reach_goal_cutoff = s["cur_goal_idx"] >= (G - cfg.last_goal_repeat)
    height_cut = s["root_states"][:, 2] < -0.25
    reach_last_goal = torch.norm(s["root_states"][:, :2] - static["env_goals"][:, -cfg.last_goal_repeat, :2],
                                 dim=1) < cfg.next_goal_threshold
    time_out = (s["episode_length_buf"] > cfg.max_episode_length) | reach_goal_cutoff
    reset = reset | time_out | roll_cut | pitch_cut | height_cut | g["leave_goal_ids"]
    if cfg.use_camera:
        reset = reset | reach_last_goal
    s["reset_buf"], s["time_out_buf"], s["reach_goal"] = reset, time_out, reach_goal_cutoff
    # compute_reward :412-430
    terms = reward_terms(cfg, static, s)
    rew = torch.zeros(N)
    s["episode_sums"] = s["episode_sums"].clone()
    names = cfg.reward_names
    for k, name in enumerate(names):
        r = terms[name] * (cfg.reward_scales[name] * cfg.dt)
        rew = rew + r
        s["episode_sums"][:, k] = s["episode_sums"][:, k] + r
    if cfg.only_positive_rewards:
        rew = torch.clip(rew, min=0.)
    if "termination" in cfg.reward_scales:
        r = (reset * ~time_out) * (cfg.reward_scales["termination"] * cfg.dt)
        rew = rew + r
        s["episode_sums"][:, len(names)] = s["episode_sums"][:, len(names)] + r
    s["rew_buf"] = rew
    env_ids = reset.nonzero(as_tuple=False).flatten()
    s["reset_env_ids"] = env_ids
    s["terminal_disc_states"] = s["obs_disc_buf"][env_ids]                                      # :264 (stale buffer)
    if len(env_ids) > 0:                                                                        # reset_idx :348-384
        s["cur_goal_idx"][env_ids] = 0                                                          # randomize_start False
        dof[env_ids, :, 0] = static["default_dof_pos"]                                          # _reset_dofs :798-804
        dof[env_ids, :, 1] = 0.
        s["dof_state"] = dof.reshape(N * 12, 2)
        s["obst_dof_state"] = s["obst_dof_state"].clone()
        s["obst_dof_state"][static["seesaw_dof_index"][env_ids], 0] = cfg.seesaw_dof_pos         # :825-829
        s["obst_dof_state"][:, 1] = 0.0
        rs = s["root_states"]                                                                   # _reset_root_states :840-884
        rs[env_ids] = torch.tensor(cfg.base_init_state)
        rs[env_ids, :2] = static["env_goals"][env_ids, 0, :2].clone()
        u = lambda key, lo, hi: ((hi - lo) * draws[key][env_ids] + lo)                          # noqa: E731  torch_rand_float
        rand_yaw = cfg.rand_yaw_range * u("yaw_u", -1., 1.)
        rand_pitch = torch.zeros(len(env_ids))
        root_yaw = rand_yaw + cfg.frame_ang0
        rs[env_ids, 3:7] = quat_from_euler_xyz(0 * root_yaw, rand_pitch, root_yaw)
        rs[env_ids, 0] += cfg.rand_x_range * u("x_u", -1., 0.)
        rs[env_ids, 1] += cfg.rand_y_range * u("y_u", -1., 1.)
    return s


def quat_from_euler_xyz(roll, pitch, yaw):
    """isaacgym.torch_utils.quat_from_euler_xyz (Preview 4 public definition), xyzw."""
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qx, qy, qz, qw], dim=-1)


def reward_terms(cfg: TscCfg, static, s):
    """The active `_reward_*` functions (:1779-1930)."""
    out = {}
    out["reach_goal"] = s["reached_goal_ids"].float()                                           # :1921
    norm = torch.norm(s["target_pos_rel"], dim=-1, keepdim=True)                                 # :1779-1791
    tvn = s["target_pos_rel"] / (norm + 1e-5)
    cur_vel = s["root_states"][:, 7:9]
    proj = torch.sum(tvn * cur_vel, dim=-1)
    z = s["commands"][:, 0] * 0
    rew = torch.minimum(proj, z + cfg.target_lin_vel) / (z + cfg.target_lin_vel + 1e-5)
    fast = torch.minimum(proj, z + 2.5) / (z + 2.5 + 1e-5)
    t = s["cur_obstacle_types"]
    out["tracking_goal_vel"] = torch.where((t == 0) | (t == 4), fast, rew)
    dy = ((s["target_yaw"] - s["yaw"]) + torch.pi) % (2 * torch.pi) - torch.pi                   # :1793-1797
    out["tracking_yaw"] = torch.exp(-torch.abs(dy))
    out["collision"] = torch.sum(1. * (torch.norm(s["contact_forces"][:, static["penalised_contact_indices"], :], dim=-1) > 0.1), dim=1)
    hl = s.get("action_hl_history_buf")
    if hl is None:                                                                              # :1847-1859
        out["action_hl_rate"] = torch.zeros(cfg.num_envs)
        out["latent_c_rate"] = torch.zeros(cfg.num_envs)
    else:
        out["action_hl_rate"] = torch.norm(hl[:, -2, :] - hl[:, -1, :], dim=1)
        out["latent_c_rate"] = 0.5 * (torch.abs(hl[:, -3, 0] - hl[:, -1, 0]) + torch.abs(hl[:, -2, 0] - hl[:, -1, 0]))
    feet = static["feet_indices"]                                                               # feet_edge :1899-1915
    xy = ((s["rigid_body_state"][:, feet, :2] + cfg.border_size) / cfg.horizontal_scale).round().long()
    m = static["x_edge_mask"]
    fx = torch.clip(xy[..., 0], 0, m.shape[0] - 1)
    fy = torch.clip(xy[..., 1], 0, m.shape[1] - 1)
    s["feet_at_edge"] = s["contact_filt"] & m[fx, fy]
    out["feet_edge"] = torch.sum(s["feet_at_edge"], dim=-1).float()
    return out


def post_physics_post(cfg: TscCfg, static, s_pre, rigid_body_state_post):
    """Everything after the physics step inside reset_idx (:386-285).  `rigid_body_state_post` = the refreshed tensor."""
    s = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in s_pre.items()}
    N = cfg.num_envs
    env_ids = s["reset_env_ids"]
    s["rigid_body_state"] = rigid_body_state_post
    s["episode_rew_means"] = None
    if len(env_ids) > 0:                                                                        # :386-410
        s["last_actions"][env_ids] = 0.
        s["last_dof_vel"][env_ids] = 0.
        s["last_torques_org"][env_ids] = 0.
        s["last_root_vel"][:] = 0.
        s["feet_air_time"][env_ids] = 0.
        s["obs_history_buf"][env_ids] = 0.
        s["contact_buf"][env_ids] = 0.
        s["action_history_buf"][env_ids] = 0.
        s["reach_goal_timer"][env_ids] = 0
        s["episode_rew_means"] = torch.stack([torch.mean(s["episode_sums"][:, k][env_ids]) / cfg.episode_length_s
                                              for k in range(s["episode_sums"].shape[1])])                 # :401-404
        s["episode_sums"][env_ids] = 0.
        s["episode_length_buf"][env_ids] = 0
        s["time_outs_latched"] = s["time_out_buf"].clone()
    goals = static["env_goals"]                                                                 # :271-272
    gi = s["cur_goal_idx"][:, None, None]
    s["cur_goals"] = goals.gather(1, gi.expand(-1, -1, goals.shape[-1])).squeeze(1)
    s["next_goals"] = goals.gather(1, (gi + 1).expand(-1, -1, goals.shape[-1])).squeeze(1)
    compute_observations(cfg, static, s)                                                        # :276
    dof = s["dof_state"].view(N, 12, 2)
    s["last_actions"] = s["actions"].clone()                                                    # :278-281
    s["last_dof_vel"] = dof[:, :, 1].clone()
    s["last_torques_org"] = s["torques_org"].clone()
    s["last_root_vel"] = s["root_states"][:, 7:13].clone()
    return s


def compute_observations(cfg: TscCfg, static, s):
    """:432-515."""
    N = cfg.num_envs
    dof = s["dof_state"].view(N, 12, 2)
    dof_pos, dof_vel = dof[:, :, 0], dof[:, :, 1]
    mh = s["measured_heights"]
    root_h = (s["root_states"][:, 2] - mh[:, mh.shape[1] // 2 + 1]).view(-1, 1)
    root_h_obs = root_h if cfg.root_height_obs else torch.zeros_like(root_h)
    imu = torch.stack((s["roll"], s["pitch"]), dim=1)
    if s["global_counter"] % cfg.update_interval == 0:
        dy = s["target_yaw"] - s["yaw"]
        dn = s["next_target_yaw"] - s["yaw"]
        s["delta_yaw"] = (dy + torch.pi) % (2 * torch.pi) - torch.pi
        s["delta_next_yaw"] = (dn + torch.pi) % (2 * torch.pi) - torch.pi
    delta_yaws = torch.cat([s["delta_yaw"][:, None], s["delta_next_yaw"][:, None]], dim=-1)
    key = s["rigid_body_state"][:, static["key_body_ids"], 0:3]
    flat_key = compute_flat_key_pos(s["root_states"], key)
    cf = s["contact_filt"].float()
    dd = static["default_dof_pos"]
    s["obs_disc_buf"] = torch.cat([imu, root_h, s["base_lin_vel"] * cfg.s_lin_vel_dist, s["base_ang_vel"] * cfg.s_ang_vel_dist,
                                   (dof_pos - dd) * cfg.s_dof_pos, dof_vel * cfg.s_dof_vel, flat_key * cfg.s_key_pos,
                                   cf * cfg.s_foot_contact], dim=-1)
    obs57 = torch.cat([imu, s["base_ang_vel"] * cfg.s_ang_vel, (dof_pos - dd) * cfg.s_dof_pos, dof_vel * cfg.s_dof_vel,
                       s["action_history_buf"][:, -1], cf - 0.5, flat_key * 0], dim=-1)
    priv_explicit = torch.cat([root_h_obs, s["base_lin_vel"] * cfg.s_lin_vel], dim=-1)
    priv_latent = torch.cat((static["mass_params"], static["friction_coeffs"], static["motor_strength"][0] - 1,
                             static["motor_strength"][1] - 1), dim=-1)
    types = F.one_hot(s["cur_obstacle_types"], num_classes=cfg.num_obstacle_types)
    heights = torch.clip(s["root_states"][:, 2].unsqueeze(1) - 0.3 - mh, -1, 1.)
    hist = s["obs_history_buf"].view(N, -1)
    obs = torch.cat([obs57, delta_yaws, types, heights, priv_explicit, priv_latent, hist], dim=-1)
    obs_bbc = torch.cat([obs57, priv_explicit, priv_latent], dim=-1)
    obs_bbc = torch.cat((obs_bbc, hist), dim=-1)
    obs_bbc = torch.cat([obs_bbc, s["commands"], s["latent_eps"], s["latent_c"]], dim=-1)
    fill = (s["episode_length_buf"] <= 1)[:, None, None]
    s["obs_history_buf"] = torch.where(fill, torch.stack([obs57] * cfg.history_len, dim=1),
                                       torch.cat([s["obs_history_buf"][:, 1:], obs57.unsqueeze(1)], dim=1))
    s["contact_buf"] = torch.cat([s["contact_buf"][:, 1:], cf.unsqueeze(1)], dim=1)
    c = cfg.clip_obs
    s["obs_buf"] = torch.clip(obs, -c, c)
    s["obs_bbc_buf"] = torch.clip(obs_bbc, -c, c)
    s["obs_history_buf"] = torch.clip(s["obs_history_buf"], -c, c)
    s["contact_buf"] = torch.clip(s["contact_buf"], -c, c)


def action_push(cfg: TscCfg, action_history_buf, actions, delay_step=1, clip_actions=100.0, action_scale=0.25):
    """LeggedRobot.step front half (:108-127): history shift, delayed action, clip."""
    hist = torch.cat([action_history_buf[:, 1:].clone(), actions[:, None, :].clone()], dim=1)
    delayed = hist[:, -delay_step - 1]
    c = clip_actions / action_scale
    return hist, torch.clip(delayed, -c, c)


def compute_torques(static, dof_state, actions, action_scale=0.25, hip_scale_reduction=0.5):
    """_compute_torques (:762-796), control_type 'P', randomize_motor False."""
    N = actions.shape[0]
    dof = dof_state.view(N, 12, 2)
    a = actions * action_scale
    a[:, [0, 3, 6, 9]] *= hip_scale_reduction
    tq = static["p_gains"] * (a + static["default_dof_pos"] - dof[:, :, 0]) - static["d_gains"] * dof[:, :, 1]
    return torch.clip(tq, -static["torque_limits"], static["torque_limits"]), tq

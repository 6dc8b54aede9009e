"""TSC (agility task-level controller) environment behind the reference's class API:
`LeggedRobotTSC.step(actions, action_hl_history_buf)`, `post_physics_step`, `set_commands`, `get_observations*`
mirror tsc/legged_gym/envs/base/legged_robot.py:108-149, :226-298, :699-760, with IsaacGym behind `PhysicsBackend`.

Per env step the device work is: K0 (action history push + 1-step action delay), decimation x K1 (PD torques; the TSC
config has randomize_motor False, i.e. unit motor strengths), K16 `qa_post_physics_tsc_pre`, the backend's reset step
(`gym.set_*_tensor_indexed` + `simulate` + rigid-body refresh, :381-384), K17 `qa_post_physics_tsc_post`.
`set_commands` (the high-level action -> BBC command mapping) stays a handful of torch ops.
"""
import math
import types
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import _abi, ops
from .legged_robot import PhysicsBackend


@dataclass
class TscEnvConfig:
    """Scalars of LeggedRobotCfg + Go2AgilityCfg that the hot path reads (legged_robot_config.py, go2_agility_config.py)."""
    num_envs: int = 4096
    num_bodies: int = 17
    dt: float = 0.02
    decimation: int = 4
    episode_length_s: float = 40.0
    history_len: int = 10
    contact_buf_len: int = 100
    action_buf_len: int = 8
    action_delay_step: int = 1
    clip_actions: float = 100.0
    action_scale: float = 0.25
    hip_scale_reduction: float = 0.5
    next_goal_threshold: float = 0.4
    reach_goal_delay: float = 0.02
    leave_goal_threshold: float = 4.0
    root_height_obs: bool = True
    num_goals: int = 4
    last_goal_repeat: int = 2
    num_obstacle_types: int = 6
    update_interval: int = 1
    use_camera: bool = False
    border_size: float = 5.0
    horizontal_scale: float = 0.05
    vertical_scale: float = 0.005
    target_lin_vel: float = 0.4
    only_positive_rewards: bool = True
    clip_obs: float = 100.0
    s_lin_vel: float = 0.5
    s_ang_vel: float = 0.25
    s_dof_pos: float = 1.0
    s_dof_vel: float = 0.05
    s_key_pos: float = 0.0
    s_foot_contact: float = 0.0
    s_lin_vel_dist: float = 0.0
    s_ang_vel_dist: float = 0.0
    rand_yaw_range: float = 0.2
    rand_x_range: float = 0.2
    rand_y_range: float = 0.1
    frame_ang0: float = math.pi / 2
    base_init_state: List[float] = field(default_factory=lambda: [0., 0., 0.42, 0., 0., 0., 1., 0., 0., 0., 0., 0., 0.])
    seesaw_dof_pos: float = 0.3
    randomize_action: bool = True
    action_noise: tuple = (0.8, 1.2)
    resampling_time: float = 0.02
    mocap_category: tuple = ("trot", "canter", "jump")
    mocap_category_all: tuple = ("walk", "pace", "trot", "canter", "jump")
    num_actions_c: int = 6
    command_ranges: Dict[str, list] = field(default_factory=lambda: dict(
        lin_vel_x=[[0.0, 0.6], [0.5, 1.5], [0.5, 1.5], [0.8, 2.5], [0.8, 2.0]],
        lin_vel_y=[[-0.15, 0.15], [-0.3, 0.3], [-0.3, 0.3], [-0.5, 0.5], [-0.3, 0.3]],
        ang_vel_yaw=[[-1.0, 1.0], [-1.57, 1.57], [-1.57, 1.57], [-0.5, 0.5], [-0.5, 0.5]],
        jump_height=[0.45, 0.58], locomotion_height=[0.25, 0.34]))
    # class_to_dict order (dir() = alphabetical), zero scales dropped; `termination` is applied after the >= 0 clip
    reward_scales: Dict[str, float] = field(default_factory=lambda: dict(
        action_hl_rate=-0.2, collision=-20.0, feet_edge=-1.0, latent_c_rate=-1.0, reach_goal=5.0, termination=-50.0,
        tracking_goal_vel=0.4, tracking_yaw=2.0))

    @property
    def max_episode_length(self) -> float:
        return float(math.ceil(self.episode_length_s / self.dt))

    # `cfg.env.* / cfg.domain_rand.* / ...` as the reference's TSC runner reads them (tsc/rsl_rl/runners/on_policy_runner.py:33-57,
    # 281-296): read-only views; change scalars through `LeggedRobotTSC.reconfigure(...)`, which rebuilds the kernel constants
    @property
    def env(self):
        return types.SimpleNamespace(
            num_envs=self.num_envs, n_proprio=65, n_delta_yaw=2, n_obst_type=self.num_obstacle_types, n_auxiliary=2 + self.num_obstacle_types,
            n_scan=132, n_priv=4, n_priv_latent=29, history_len=self.history_len, num_command=self.num_actions_c + len(self.mocap_category_all),
            num_observations_bbc=671 - 570, num_actions_bbc=12, num_obs_disc=49, disc_obs_len=2, next_goal_threshold=self.next_goal_threshold,
            draw_height_maps=False)

    @property
    def domain_rand(self):
        return types.SimpleNamespace(action_buf_len=self.action_buf_len, randomize_action=self.randomize_action)

    @property
    def noise(self):
        return types.SimpleNamespace(add_noise=False)

    @property
    def obstacle(self):
        return types.SimpleNamespace(curriculum=False)

    @property
    def reward_names(self):
        return [k for k in self.reward_scales if k != "termination"] + ["termination"]


def _mask(indices) -> int:
    m = 0
    for i in indices:
        m |= 1 << int(i)
    return m


class LeggedRobotTSC:
    num_obs, num_obs_bbc, num_obs_disc, num_actions, dim_c = 800, 671, 49, 12, 5
    num_privileged_obs = None

    def __init__(self, cfg: TscEnvConfig, physics: PhysicsBackend, static: Dict[str, torch.Tensor], device="cuda", seed=0):
        self.cfg, self.physics, self.device = cfg, physics, torch.device(device)
        dev, N, B = self.device, cfg.num_envs, cfg.num_bodies
        self.num_envs, self.dt = N, cfg.dt
        self.max_episode_length = cfg.max_episode_length
        self.depth = None
        f = lambda *s: torch.zeros(*s, device=dev)                             # noqa: E731
        u8 = lambda *s: torch.zeros(*s, device=dev, dtype=torch.uint8)         # noqa: E731
        self.static = {k: v.to(dev).contiguous() for k, v in static.items()}
        st = self.static
        self.static["x_edge_mask_u8"] = st["x_edge_mask"].to(torch.uint8).contiguous()
        self.env_goals, self.obstacle_types = st["env_goals"], st["obstacle_types"]
        self.num_height_points = st["height_points"].shape[1]
        self.default_dof_pos = st["default_dof_pos"]
        self.p_gains, self.d_gains, self.torque_limits = st["p_gains"], st["d_gains"], st["torque_limits"]
        self._unit_strength = torch.ones(2, N, 12, device=dev)
        # carried buffers (names = the reference's attributes)
        self.episode_length_buf = torch.zeros(N, device=dev, dtype=torch.int64)
        self.last_contacts = u8(N, 4)
        self.reach_goal_timer = f(N)
        self.cur_goal_idx = torch.zeros(N, device=dev, dtype=torch.int64)
        self.cur_goals, self.next_goals = f(N, 3), f(N, 3)
        self.actions, self.torques, self.torques_org = f(N, 12), f(N, 12), f(N, 12)
        self.last_actions, self.last_dof_vel, self.last_torques_org, self.last_root_vel = f(N, 12), f(N, 12), f(N, 12), f(N, 6)
        self.commands, self.latent_eps, self.latent_c = f(N, 5), f(N, 1), f(N, 5)
        self.episode_sums_buf = f(N, _abi.TSC_NUM_REWARDS)
        self.feet_air_time = f(N, 4)
        self.obs_history_buf = f(N, cfg.history_len, 57)
        self.action_history_buf = f(N, cfg.action_buf_len, 12)
        self._contact_ring = f(N, cfg.contact_buf_len, 4)
        self.measured_heights = f(N, self.num_height_points)
        self.delta_yaw, self.delta_next_yaw = f(N), f(N)
        self.base_lin_vel, self.base_ang_vel, self.projected_gravity, self.base_lin_acc, self._rpy = (f(N, 3) for _ in range(5))
        self.contact_filt = u8(N, 4)
        self.target_yaw, self.next_target_yaw = f(N), f(N)
        self.cur_obstacle_types = torch.zeros(N, device=dev, dtype=torch.int64)
        self.reached_goal_ids, self._reach_goal_cutoff, self.feet_at_edge = u8(N).bool(), u8(N).bool(), u8(N, 4).bool()
        self.reset_buf, self.time_out_buf, self._time_outs_latched = u8(N).bool(), u8(N).bool(), u8(N).bool()
        self.rew_buf = f(N)
        self._episode_rew_means = f(_abi.TSC_NUM_REWARDS)
        self._num_resets = torch.zeros(1, device=dev, dtype=torch.int32)
        self._workspace = torch.zeros(16, device=dev, dtype=torch.float64)
        self.obs_buf, self.obs_bbc_buf, self.obs_disc_buf = f(N, self.num_obs), f(N, self.num_obs_bbc), f(N, self.num_obs_disc)
        self.privileged_obs_buf = None
        self._reset_ids = torch.zeros(N, device=dev, dtype=torch.int64)
        self._reset_ids_i32 = torch.zeros(N, device=dev, dtype=torch.int32)
        self._terminal_disc = f(N, self.num_obs_disc)
        self._reset_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.action_hl_history_buf: Optional[torch.Tensor] = None
        self.extras: Dict = {}
        self.global_counter = self.total_env_steps_counter = self.common_step_counter = 0
        self.seed = seed
        self._draws: Optional[Dict[str, torch.Tensor]] = None
        all_cats = list(cfg.mocap_category_all)
        self.mocap_indices = torch.tensor([all_cats.index(c) for c in cfg.mocap_category], device=dev)
        self._const = self._build_const()

    # ---- reference-named views -----------------------------------------------------------------------------
    @property
    def root_states(self):
        return self.physics.root_states

    @property
    def dof_state(self):
        return self.physics.dof_state

    @property
    def roll(self):
        return self._rpy[:, 0]

    @property
    def pitch(self):
        return self._rpy[:, 1]

    @property
    def yaw(self):
        return self._rpy[:, 2]

    @property
    def episode_sums(self):
        return {k: self.episode_sums_buf[:, i] for i, k in enumerate(self.cfg.reward_names)}

    @property
    def contact_buf(self):
        """(N,L,4) in the reference's order (oldest first); the kernel keeps a ring with slot `counter % L` newest."""
        L = self.cfg.contact_buf_len
        head = (self.common_step_counter - 1) % L
        return torch.roll(self._contact_ring, shifts=L - 1 - head, dims=1)

    def _build_const(self) -> _abi.QaTscConst:
        cfg, st, c = self.cfg, self.static, _abi.QaTscConst()
        c.num_bodies = cfg.num_bodies
        for j in range(4):
            c.feet_indices[j] = int(st["feet_indices"][j])
        c.termination_body_mask = _mask(st["termination_contact_indices"].tolist())
        c.penalised_body_mask = _mask(st["penalised_contact_indices"].tolist())
        c.dt, c.max_episode_length, c.episode_length_s = cfg.dt, cfg.max_episode_length, cfg.episode_length_s
        c.next_goal_threshold, c.leave_goal_threshold = cfg.next_goal_threshold, cfg.leave_goal_threshold
        c.reach_goal_delay_steps = cfg.reach_goal_delay / cfg.dt
        c.num_goals_total, c.num_goals_per_obstacle = self.env_goals.shape[1], cfg.num_goals
        c.last_goal_repeat, c.num_obstacle_types = cfg.last_goal_repeat, cfg.num_obstacle_types
        c.update_interval, c.use_camera = cfg.update_interval, int(cfg.use_camera)
        c.root_height_obs, c.only_positive_rewards = int(cfg.root_height_obs), int(cfg.only_positive_rewards)
        c.target_lin_vel = cfg.target_lin_vel
        for i, k in enumerate(cfg.reward_names):
            c.reward_scale[i] = cfg.reward_scales.get(k, 0.0) * cfg.dt           # scale *= dt in double (:1112)
        for i in range(12):
            c.default_dof_pos[i] = float(st["default_dof_pos"].flatten()[i])
        for i in range(13):
            c.base_init_state[i] = cfg.base_init_state[i]
        c.rand_yaw_range, c.rand_x_range, c.rand_y_range = cfg.rand_yaw_range, cfg.rand_x_range, cfg.rand_y_range
        c.frame_ang0, c.seesaw_dof_pos = cfg.frame_ang0, cfg.seesaw_dof_pos
        for k in ("s_lin_vel", "s_ang_vel", "s_dof_pos", "s_dof_vel", "s_key_pos", "s_foot_contact", "s_lin_vel_dist", "s_ang_vel_dist"):
            setattr(c, k, getattr(cfg, k))
        c.clip_obs, c.num_height_points, c.contact_ring_len = cfg.clip_obs, self.num_height_points, cfg.contact_buf_len
        return c

    def set_parity_draws(self, draws: Optional[Dict[str, torch.Tensor]]) -> None:
        """dense uniforms {yaw_u, x_u, y_u} (N,) replacing the in-kernel Philox stream of the reset randomisation."""
        self._draws = None if draws is None else {k: v.to(self.device).float().contiguous() for k, v in draws.items()}

    def load_state(self, snap: Dict[str, torch.Tensor]) -> None:
        """Overwrite the carried buffers from a recorded snapshot (tests / bench)."""
        dev = self.device
        for k in ("episode_length_buf", "reach_goal_timer", "cur_goal_idx", "cur_goals", "next_goals", "actions",
                  "torques_org", "last_actions", "last_dof_vel", "last_torques_org", "last_root_vel", "commands", "latent_eps",
                  "latent_c", "feet_air_time", "obs_history_buf", "action_history_buf", "measured_heights", "delta_yaw",
                  "delta_next_yaw", "obs_disc_buf"):
            if k in snap:
                getattr(self, k).copy_(snap[k].to(dev))
        if "last_contacts" in snap:
            self.last_contacts.copy_(snap["last_contacts"].to(dev).to(torch.uint8))
        if "episode_sums" in snap:
            self.episode_sums_buf.copy_(snap["episode_sums"].to(dev))
        if "contact_buf" in snap:      # reference order (oldest first) -> ring whose newest slot is (counter - 1) % L
            L = self.cfg.contact_buf_len
            head = (int(snap.get("common_step_counter", 0)) - 1) % L
            self._contact_ring.copy_(torch.roll(snap["contact_buf"].to(dev), shifts=-(L - 1 - head), dims=1))
        self.common_step_counter = int(snap.get("common_step_counter", self.common_step_counter))
        self.global_counter = int(snap.get("global_counter", self.global_counter))

    # ---- args -------------------------------------------------------------------------------------------------
    def _args(self) -> _abi.QaTscStepArgs:
        ph, st, a = self.physics, self.static, _abi.QaTscStepArgs()
        p = lambda t: None if t is None else t.data_ptr()                      # noqa: E731
        a.num_envs, a.global_counter = self.num_envs, self.global_counter
        a.root_states, a.dof_state = p(ph.root_states), p(ph.dof_state)
        a.rigid_body_state, a.contact_forces = p(ph.rigid_body_state), p(ph.contact_forces)
        obst = getattr(ph, "obst_dof_state", None)
        a.obst_dof_state, a.seesaw_dof_index = p(obst), p(st.get("seesaw_dof_index"))
        a.num_obst_dofs = 0 if obst is None else obst.shape[0]
        a.terrain = ops.terrain_struct(st["height_samples"], self.cfg.border_size, self.cfg.horizontal_scale, self.cfg.vertical_scale)
        a.x_edge_mask, a.height_points, a.env_goals = p(st["x_edge_mask_u8"]), p(st["height_points"]), p(st["env_goals"])
        a.obstacle_types, a.mass_params = p(st["obstacle_types"]), p(st["mass_params"])
        a.friction_coeffs, a.motor_strength = p(st["friction_coeffs"]), p(st["motor_strength"])
        a.episode_length_buf, a.last_root_vel_in, a.last_contacts = p(self.episode_length_buf), p(self.last_root_vel), p(self.last_contacts)
        a.reach_goal_timer, a.cur_goal_idx, a.cur_goals, a.next_goals = p(self.reach_goal_timer), p(self.cur_goal_idx), p(self.cur_goals), p(self.next_goals)
        a.actions, a.torques_org = p(self.actions), p(self.torques_org)
        a.last_actions, a.last_dof_vel, a.last_torques_org, a.last_root_vel = p(self.last_actions), p(self.last_dof_vel), p(self.last_torques_org), p(self.last_root_vel)
        a.commands, a.latent_eps, a.latent_c = p(self.commands), p(self.latent_eps), p(self.latent_c)
        hl = self.action_hl_history_buf
        a.action_hl_history_buf = p(hl)
        self._const.hl_hist_len, self._const.hl_action_dim = (0, 0) if hl is None else (hl.shape[1], hl.shape[2])
        a.episode_sums, a.feet_air_time = p(self.episode_sums_buf), p(self.feet_air_time)
        a.obs_history_buf, a.action_history_buf, a.contact_buf = p(self.obs_history_buf), p(self.action_history_buf), p(self._contact_ring)
        a.contact_ring_head = (self.common_step_counter - 1) % self.cfg.contact_buf_len
        a.measured_heights, a.delta_yaw, a.delta_next_yaw = p(self.measured_heights), p(self.delta_yaw), p(self.delta_next_yaw)
        a.base_lin_vel, a.base_ang_vel, a.projected_gravity = p(self.base_lin_vel), p(self.base_ang_vel), p(self.projected_gravity)
        a.base_lin_acc, a.rpy, a.contact_filt = p(self.base_lin_acc), p(self._rpy), p(self.contact_filt)
        a.target_yaw, a.next_target_yaw, a.cur_obstacle_types = p(self.target_yaw), p(self.next_target_yaw), p(self.cur_obstacle_types)
        a.reached_goal, a.reach_goal_cutoff, a.feet_at_edge = p(self.reached_goal_ids), p(self._reach_goal_cutoff), p(self.feet_at_edge)
        a.reset_buf, a.time_out_buf, a.time_outs_latched = p(self.reset_buf), p(self.time_out_buf), p(self._time_outs_latched)
        a.rew_buf, a.episode_rew_means, a.num_resets, a.workspace = p(self.rew_buf), p(self._episode_rew_means), p(self._num_resets), p(self._workspace)
        a.obs_buf, a.obs_bbc_buf, a.obs_disc_buf = p(self.obs_buf), p(self.obs_bbc_buf), p(self.obs_disc_buf)
        d = self._draws
        a.yaw_u, a.x_u, a.y_u = (None, None, None) if d is None else (p(d["yaw_u"]), p(d["x_u"]), p(d["y_u"]))
        a.rng_seed, a.rng_step = self.seed, self.common_step_counter
        return a

    # ---- the step ------------------------------------------------------------------------------------------------
    def post_physics_step(self):
        """:226-298.  Returns (reset_env_ids, terminal_disc_states) like the reference (one 4-byte D2H for the count)."""
        ph = self.physics
        ph.refresh()
        self.common_step_counter += 1
        a = self._args()
        ops.post_physics_tsc(self._const, a, "pre")
        # reset_buf.nonzero() + stale disc-obs rows of the reset envs (:263-264), device side
        ops.compact_resets(self.reset_buf, self.obs_disc_buf, self._reset_ids, self._reset_ids_i32, self._terminal_disc, self._reset_count)
        count = int(self._reset_count.item())
        if count > 0:                                                        # :381-384
            ph.set_states_indexed(self._reset_ids_i32, count)
            ph.simulate()
            ph.refresh()
            a.rigid_body_state = ph.rigid_body_state.data_ptr()
            self.extras["episode"] = {"rew_" + k: self._episode_rew_means[i] for i, k in enumerate(self.cfg.reward_names)}
            self.extras["time_outs"] = self._time_outs_latched
        self.extras["reach_goal"] = self._reach_goal_cutoff
        terminal = self._terminal_disc[:count].clone()
        ops.post_physics_tsc(self._const, a, "post")
        if self.depth is not None:                                           # :275, after reset_idx: fresh episodes refill
            self.depth.update_depth_buffer(self.episode_length_buf, self.global_counter)   # their whole history (K14)
        return self._reset_ids[:count], terminal

    def reconfigure(self, **changes) -> None:
        """Change config scalars after construction (the student loop raises `next_goal_threshold`, tsc
        on_policy_runner.py:286) and rebuild the kernels' constant block."""
        for k, v in changes.items():
            if not hasattr(self.cfg, k):
                raise AttributeError(k)
            setattr(self.cfg, k, v)
        self._const = self._build_const()
        self._cmd_range_tensors = None                                       # set_commands' cached range table

    def attach_depth(self, depth) -> None:
        """Student path (`--use_camera`): `depth` is a `qa_b200.depth.DepthBuffer` bound to the simulator's camera tensors;
        `post_physics_step` then refreshes it every `update_interval` steps and `step` hands out `extras["depth"]` (:145-148)."""
        self.depth = depth

    @property
    def depth_buffer(self):
        return None if self.depth is None else self.depth.depth_buffer

    def _pre_physics(self, actions, action_hl_history_buf=None):
        cfg = self.cfg
        self.action_hl_history_buf = action_hl_history_buf
        self.global_counter += 1
        self.total_env_steps_counter += 1
        ops.action_push(actions.contiguous(), self.action_history_buf, self.actions, cfg.action_delay_step,
                        cfg.clip_actions / cfg.action_scale)
        for _ in range(cfg.decimation):                                        # :128-138
            ops.pd_torques(self.actions, self.physics.dof_state, self._unit_strength, self.p_gains, self.d_gains,
                           self.default_dof_pos.flatten(), self.torque_limits, self.torques, self.torques_org,
                           cfg.action_scale, cfg.hip_scale_reduction)
            self.physics.set_dof_actuation_force(self.torques)
            self.physics.simulate()

    def step(self, actions, action_hl_history_buf=None):
        """:108-149; 7-tuple (obs, privileged_obs, rew, reset, extras, reset_env_ids, terminal_disc_states)."""
        self._pre_physics(actions, action_hl_history_buf)
        ids, terminal = self.post_physics_step()
        self.extras["delta_yaw_ok"] = torch.abs(self.delta_yaw) < 0.6
        d = self.depth
        if d is not None and d.cfg.use_camera and self.global_counter % d.cfg.update_interval == 0:
            self.extras["depth"] = d.depth_buffer[:, -2]                        # "have already selected last one" (:146)
        else:
            self.extras["depth"] = None
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras, ids, terminal

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def get_observations_bbc(self):
        return self.obs_bbc_buf

    def get_observations_disc(self):
        return self.obs_disc_buf

    def get_history_observations(self):
        return self.obs_history_buf

    @torch.no_grad()
    def set_commands(self, actions, action_noise_u: Optional[torch.Tensor] = None):
        """High-level action (mode index + 3 x 6 continuous) -> BBC command vector (:699-760).  The reference writes the
        envs whose episode step is a multiple of the resampling period through `nonzero()` index lists (a host sync per
        step); here the same values are computed for every env and selected with the mask, in place -- bit-equal results
        (`oracle/check_interop.py tsc`), no sync."""
        cfg, dev = self.cfg, self.device
        period = int(cfg.resampling_time / cfg.dt)
        due = (self.episode_length_buf % period == 0)
        actions_d = actions[:, 0].to(torch.long)
        m = self.mocap_indices[actions_d]
        cols = actions_d[:, None] * cfg.num_actions_c + torch.arange(cfg.num_actions_c, device=dev) + 1
        cmd = torch.clip(torch.gather(actions, 1, cols), -1, 1)
        rt = getattr(self, "_cmd_range_tensors", None)
        if rt is None:                                                       # (3, dim_c, 2): lin_vel_x / lin_vel_y / ang_vel_yaw
            r = cfg.command_ranges
            rt = self._cmd_range_tensors = torch.tensor([r["lin_vel_x"], r["lin_vel_y"], r["ang_vel_yaw"]], device=dev)
        self.latent_c.copy_(torch.where(due[:, None], torch.nn.functional.one_hot(m, self.latent_c.shape[1]).to(self.latent_c.dtype),
                                        self.latent_c))
        self.latent_eps[:, 0].copy_(torch.where(due, cmd[:, -1], self.latent_eps[:, 0]))
        cmd01 = (cmd + 1) / 2
        lo, hi = rt[:, m, 0].t(), rt[:, m, 1].t()                              # (N, 3)
        new = torch.zeros_like(self.commands)
        new[:, 0:3] = lo + (hi - lo) * cmd01[:, 0:3]
        jump = m == (self.dim_c - 1)
        jl, jh = cfg.command_ranges["jump_height"]
        ll, lh = cfg.command_ranges["locomotion_height"]
        new[:, 3] = (jl + (jh - jl) * cmd01[:, 3]) * jump.float()
        new[:, 4] = (ll + (lh - ll) * cmd01[:, 4]) * (~jump).float()
        self.commands.copy_(torch.where(due[:, None], new, self.commands))
        if cfg.randomize_action:
            lo_n, hi_n = cfg.action_noise
            u = torch.rand(self.commands.shape, device=dev) if action_noise_u is None else action_noise_u
            self.commands *= (hi_n - lo_n) * u + lo_n
        return torch.cat([self.commands, self.latent_eps, self.latent_c], dim=-1)


class RecordedPhysicsTSC(PhysicsBackend):
    """Recorded / synthetic simulator state for the TSC env: each snapshot holds the tensors IsaacGym would expose after
    the decimation loop, plus `rigid_body_state_post` = what the rigid-body refresh returns after the physics step that
    reset_idx takes (:381-384)."""

    def __init__(self, snapshots):
        self.snapshots, self.cursor, self._after_reset = snapshots, -1, False
        self._bind(0)

    def _bind(self, i):
        s = self.snapshots[i]
        self.root_states, self.dof_state, self.contact_forces = s["root_states"], s["dof_state"], s["contact_forces"]
        self.rigid_body_state, self.obst_dof_state = s["rigid_body_state"], s.get("obst_dof_state")

    def set_dof_actuation_force(self, torques):
        pass

    def simulate(self):
        pass

    def set_states_indexed(self, env_ids_i32, count):
        self._after_reset = True

    def refresh(self):
        if self._after_reset:
            self._after_reset = False
            self.rigid_body_state = self.snapshots[self.cursor]["rigid_body_state_post"]
            return
        self.cursor = (self.cursor + 1) % len(self.snapshots)
        self._bind(self.cursor)


from .rsl_rl.vec_env import VecEnv  # noqa: E402  (bottom of the module: rsl_rl imports nothing from here)

VecEnv.register(LeggedRobotTSC)

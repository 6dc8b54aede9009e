"""`LeggedRobot` -- the BBC vectorised env over the fused sm_100a kernels.

Host-side mirror of `bbc/legged_gym/envs/base/legged_robot.py` (reference) for the per-step hot
path: same attribute names, same `step / reset / get_observations / get_privileged_observations /
get_disc_observations` API and the same 7-tuple from `step` (:78-115), so `bbc/rsl_rl`'s runner and
this repo's runner drive it unchanged.  Physics stays outside: a *physics backend* object owns the
simulator tensors (`root_states`, `dof_state`, `rigid_body_state`, `contact_forces`) -- IsaacGym in
production, `RecordedPhysics` for tests / benchmarks where IsaacGym is not
installable.

Per env step the host issues exactly:  K0 action push, `decimation` x (K1 torques + backend
simulate), K2 fused post-physics, reset compaction -- no per-term Python, no host sync unless the
caller asks for the variable-length `reset_env_ids` (`step()` does, `step_device()` does not).
"""
import ctypes as C
import os
import types
from typing import Dict, Optional

import torch

from . import _abi, ops
from . import config as K
from .config import BbcEnvConfig
from .mocap import MocapTable


class PhysicsBackend:
    """What the env needs from a simulator (IsaacGym's tensor API, gymtorch.wrap_tensor views).

    Attributes (CUDA tensors owned by the simulator, updated in place by it and by the env):
      root_states (N,13), dof_state (N*12,2), rigid_body_state (N*B,13), contact_forces (N,B,3)
    """
    root_states: torch.Tensor
    dof_state: torch.Tensor
    rigid_body_state: torch.Tensor
    contact_forces: torch.Tensor

    def set_dof_actuation_force(self, torques: torch.Tensor) -> None: ...      # gym.set_dof_actuation_force_tensor
    def simulate(self) -> None: ...                                            # gym.simulate + fetch_results + refresh_dof_state
    def refresh(self) -> None: ...                                             # refresh_{actor_root_state,net_contact_force,rigid_body_state}
    def set_states_indexed(self, env_ids_i32: torch.Tensor, count) -> None: ...  # set_{dof,actor_root}_state_tensor_indexed
    def set_root_states_all(self) -> None: ...                                 # set_actor_root_state_tensor (push)


class RecordedPhysics(PhysicsBackend):
    """Plays back recorded / synthetic post-physics state snapshots: `refresh()`
    makes the next snapshot current.  The env's in-place resets land in the current snapshot, as they
    would in simulator memory."""

    def __init__(self, snapshots):
        self.snapshots = snapshots
        self.cursor = -1
        self._bind(0)

    def _bind(self, i):
        s = self.snapshots[i]
        self.root_states, self.dof_state = s["root_states"], s["dof_state"]
        self.rigid_body_state, self.contact_forces = s["rigid_body_state"], s["contact_forces"]

    def set_dof_actuation_force(self, torques):
        pass

    def simulate(self):
        pass

    def refresh(self):
        self.cursor = (self.cursor + 1) % len(self.snapshots)
        self._bind(self.cursor)

    def set_states_indexed(self, env_ids_i32, count):
        pass

    def set_root_states_all(self):
        pass


class LeggedRobot:
    """BBC go2_locomotion env.  `static` carries the per-env constants the reference creates in
    `_create_envs` / `_init_buffers` (motor_strength, mass_params_tensor, friction_coeffs_tensor,
    env_origins, height_samples, p/d gains ...), see `qa_b200.synthetic.make_static`."""

    def __init__(self, cfg: BbcEnvConfig, physics: PhysicsBackend, static: Dict[str, torch.Tensor],
                 mocap: MocapTable, device="cuda:0", seed: int = 1, bulk_store: bool = True,
                 keep_contact_rings: bool = True, tiled: bool = True):
        _abi.load()                                        # fail loudly if the CUDA library is missing
        self.cfg, self.physics, self.device = cfg, physics, torch.device(device)
        dev, N = self.device, cfg.num_envs
        self.num_envs, self.num_dof, self.num_actions = N, K.NUM_DOF, K.NUM_ACTIONS
        self.num_obs = self.num_privileged_obs = K.NUM_OBS
        self.num_obs_disc = K.NUM_OBS_DISC
        self.num_bodies = cfg.num_bodies
        self.dt = cfg.dt
        self.max_episode_length_s = cfg.episode_length_s
        self.max_episode_length = cfg.max_episode_length
        self.dim_c = K.DIM_C
        self.mocap_category = list(K.MOCAP_CATEGORY)
        self.mocap_category_all = list(K.MOCAP_CATEGORY)
        self.reward_names = list(K.REWARD_NAMES)
        self.obs_scales = types.SimpleNamespace(                               # cfg.normalization.obs_scales (:139-148); the
            lin_vel=cfg.s_lin_vel, ang_vel=cfg.s_ang_vel, dof_pos=cfg.s_dof_pos, dof_vel=cfg.s_dof_vel, key_pos=cfg.s_key_pos,
            foot_contact=cfg.s_foot_contact, lin_vel_dist=cfg.s_lin_vel_dist, ang_vel_dist=cfg.s_ang_vel_dist)   # MotionLoader reads it
        self.reward_scales = dict(zip(K.REWARD_NAMES, cfg.reward_scales_dt()))
        self.task_obs_weight_decay = cfg.task_obs_weight_decay
        self.task_obs_weight_decay_steps = cfg.task_obs_weight_decay_steps
        self.task_obs_weight = 1.0
        self.common_step_counter = 0
        self.global_counter = 0
        self.delay = 0
        self._delay_schedule = list(cfg.action_curr_step)
        self.seed = seed
        self.extras = {}

        f32 = dict(device=dev, dtype=torch.float32)
        to = lambda t: t.to(dev).contiguous()                                  # noqa: E731
        self.motor_strength = to(static["motor_strength"])
        self.mass_params_tensor = to(static["mass_params_tensor"])
        self.friction_coeffs_tensor = to(static["friction_coeffs_tensor"])
        self.env_origins = to(static["env_origins"])
        self.height_samples = to(static["height_samples"])
        self.height_points = to(static["height_points"])
        self.default_dof_pos = to(static["default_dof_pos"])
        self.p_gains, self.d_gains = to(static["p_gains"]), to(static["d_gains"])
        self.torque_limits = to(static["torque_limits"])
        self.dof_vel_limits = to(static["dof_vel_limits"])
        self.dof_pos_limits = to(static["dof_pos_limits"])
        self.noise_scale_vec = to(static["noise_scale_vec"])
        self.prior_parameters = to(static["prior_parameters"])
        self.prior_prob = self.prior_parameters.clone()
        self._prior_cdf = torch.zeros(K.DIM_C, device=dev, dtype=torch.float32)   # live CDF the fused step samples modes from
        self.mocap = mocap.to(dev)

        z = lambda *s: torch.zeros(*s, **f32)                                  # noqa: E731
        self.actions, self.last_actions = z(N, 12), z(N, 12)
        self.torques, self.torques_org, self.last_torques_org = z(N, 12), z(N, 12), z(N, 12)
        self.last_dof_vel, self.last_root_vel = z(N, 12), z(N, 6)
        self.action_history_buf = z(N, K.ACTION_BUF_LEN, 12)
        self.obs_history_buf = z(N, K.HISTORY_LEN, K.NUM_PROP)
        self.commands, self.latent_eps = z(N, K.NUM_COMMANDS), z(N, 1)
        self.latent_c = z(N, K.DIM_C)
        self.latent_c[:, 0] = 1.0
        self._episode_sums = z(N, K.EPISODE_SUMS_PITCH)                        # env-major, see DESIGN.md
        self.feet_air_time = z(N, 4)
        self.episode_length_buf = torch.zeros(N, device=dev, dtype=torch.long)
        self.last_contacts = torch.zeros(N, 4, device=dev, dtype=torch.bool)
        self.contact_filt = torch.zeros(N, 4, device=dev, dtype=torch.bool)
        self.feet_forces = z(N, 4)
        self.base_lin_vel, self.base_ang_vel, self.projected_gravity = z(N, 3), z(N, 3), z(N, 3)
        self._rpy, self._root_h = z(N, 3), z(N)
        self.rew_buf = z(N)
        self.reset_buf = torch.ones(N, device=dev, dtype=torch.bool)
        self.time_out_buf = torch.zeros(N, device=dev, dtype=torch.bool)
        self._time_outs_latched = torch.zeros(N, device=dev, dtype=torch.bool)
        self._episode_rew_means = z(K.NUM_REWARDS)
        self._num_resets = torch.zeros(1, device=dev, dtype=torch.int32)
        self._workspace = torch.zeros(16, device=dev, dtype=torch.float64)
        self._ring_len = K.CONTACT_BUF_LEN if keep_contact_rings else 0
        self._ring_head = -1
        self._contact_ring = z(N, K.CONTACT_BUF_LEN, 4) if keep_contact_rings else None
        self._contact_force_ring = z(N, K.CONTACT_BUF_LEN, 4) if keep_contact_rings else None
        # ping-pong outputs: the reference REBINDS fresh tensors every step (:321), so a caller may
        # still hold last step's obs; obs_disc must survive one step for terminal states (:153-154)
        self._obs = [z(N, K.OBS_WIDTH), z(N, K.OBS_WIDTH)]
        self._priv = [z(N, K.OBS_WIDTH), z(N, K.OBS_WIDTH)]
        self._disc = [z(N, K.NUM_OBS_DISC), z(N, K.NUM_OBS_DISC)]
        self._pp = 0
        self.obs_buf, self.privileged_obs_buf, self.obs_disc_buf = self._obs[0], self._priv[0], self._disc[0]
        # reset compaction outputs (padded to N; `count` valid)
        self._reset_ids = torch.zeros(N, device=dev, dtype=torch.int64)
        self._reset_ids_i32 = torch.zeros(N, device=dev, dtype=torch.int32)
        self._terminal_disc = z(N, K.NUM_OBS_DISC)
        self._reset_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self._count_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._measured_heights = z(N, cfg.num_height_points)

        self._const = ops.bbc_const(cfg, self.prior_parameters.tolist())
        self.refresh_prior()
        # kernel variant request; the library falls back to the warp-per-env kernel when a tile constraint fails
        self._flags = (_abi.QA_K2_BULK_STORE if (bulk_store and N % 4 == 0) else 0) | (_abi.QA_K2_TILED if tiled else 0)
        if os.environ.get("QA_K2_PDL", "1") == "1":         # programmatic dependent launch of the fused step (include/qa_b200.h)
            self._flags |= _abi.QA_K2_PDL
        self._draws = None
        # device-resident step counter: lets a captured CUDA graph of post_physics_step be replayed (the Philox
        # counter, the push schedule and the contact-ring head advance on the device)
        self._step_state = torch.zeros(2, device=dev, dtype=torch.int64)
        self.device_step_counter = False
        self.k2_events = None
        self._args = self._build_args()

    # ---- reference-named views ------------------------------------------------------------------
    @property
    def root_states(self):
        return self.physics.root_states

    @property
    def dof_state(self):
        return self.physics.dof_state

    @property
    def dof_pos(self):
        return self.physics.dof_state.view(self.num_envs, self.num_dof, 2)[..., 0]

    @property
    def dof_vel(self):
        return self.physics.dof_state.view(self.num_envs, self.num_dof, 2)[..., 1]

    @property
    def base_quat(self):
        return self.physics.root_states[:, 3:7]

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
        """dict name -> (N,) view, like the reference's `self.episode_sums` (:938-940)."""
        return {k: self._episode_sums[:, i] for i, k in enumerate(K.REWARD_NAMES)}

    @property
    def measured_heights(self):
        """All 187 scan points (K3).  The fused step only needs the centre point and samples it itself."""
        ops.height_scan(self.physics.root_states, self.height_points, self.height_samples, self.cfg.border_size,
                        self.cfg.horizontal_scale, self.cfg.vertical_scale, self._measured_heights)
        return self._measured_heights

    def _ring_view(self, ring):
        h = self._ring_head
        return torch.cat([ring[:, h + 1:], ring[:, :h + 1]], dim=1) if h >= 0 else ring

    @property
    def contact_buf(self):
        return self._ring_view(self._contact_ring)

    @property
    def contact_force_buf(self):
        return self._ring_view(self._contact_force_ring)

    # ---- argument struct --------------------------------------------------------------------------
    def _build_args(self) -> _abi.QaBbcStepArgs:
        a = _abi.QaBbcStepArgs()
        cfg, p = self.cfg, ops._p
        f, i64, i32, f64 = torch.float32, torch.int64, torch.int32, torch.float64
        a.num_envs, a.obs_pitch = self.num_envs, K.OBS_WIDTH
        a.contact_ring_len = self._ring_len
        a.flags = self._flags
        a.rng_seed = self.seed & 0xFFFFFFFFFFFFFFFF
        a.motor_strength = p(self.motor_strength, f)
        a.mass_params = p(self.mass_params_tensor, f)
        a.friction_coeffs = p(self.friction_coeffs_tensor, f)
        a.env_origins = p(self.env_origins, f)
        a.noise_scale_vec = p(self.noise_scale_vec, f)
        a.terrain = ops.terrain_struct(self.height_samples, cfg.border_size, cfg.horizontal_scale, cfg.vertical_scale)
        a.mocap = ops.mocap_struct(self.mocap)
        a.episode_length_buf = p(self.episode_length_buf, i64)
        a.last_contacts = self.last_contacts.data_ptr()
        a.commands, a.latent_eps, a.latent_c = p(self.commands, f), p(self.latent_eps, f), p(self.latent_c, f)
        a.actions, a.last_actions = p(self.actions, f), p(self.last_actions, f)
        a.torques_org, a.last_torques_org = p(self.torques_org, f), p(self.last_torques_org, f)
        a.last_dof_vel, a.last_root_vel = p(self.last_dof_vel, f), p(self.last_root_vel, f)
        a.action_history_buf, a.obs_history_buf = p(self.action_history_buf, f), p(self.obs_history_buf, f)
        a.episode_sums, a.feet_air_time = p(self._episode_sums, f), p(self.feet_air_time, f)
        a.contact_buf = p(self._contact_ring, f)
        a.contact_force_buf = p(self._contact_force_ring, f)
        a.rew_buf = p(self.rew_buf, f)
        a.reset_buf, a.time_out_buf = self.reset_buf.data_ptr(), self.time_out_buf.data_ptr()
        a.base_lin_vel, a.base_ang_vel = p(self.base_lin_vel, f), p(self.base_ang_vel, f)
        a.projected_gravity, a.rpy = p(self.projected_gravity, f), p(self._rpy, f)
        a.feet_forces, a.contact_filt = p(self.feet_forces, f), self.contact_filt.data_ptr()
        a.root_h = p(self._root_h, f)
        a.episode_rew_means = p(self._episode_rew_means, f)
        a.time_outs_latched = self._time_outs_latched.data_ptr()
        a.num_resets = p(self._num_resets, i32)
        a.workspace = p(self._workspace, f64)
        a.push_interval = int(cfg.push_interval) if cfg.push_robots else 0
        a.prior_cdf = p(self._prior_cdf, f)
        return a

    def set_parity_draws(self, draws: Optional[Dict[str, torch.Tensor]]) -> None:
        """Parity mode: consume pre-drawn per-env randoms (qa_b200.synthetic.make_rng_draws) instead of
        the in-kernel Philox stream.  `None` switches back to Philox."""
        a, p = self._args, ops._p
        if draws is None:
            self._draws = None
            for k in ("noise_u", "rs_eps_u", "rs_c_idx", "rs_cmd_u", "rt_eps_u", "rt_c_idx", "rt_cmd_u", "push_u",
                      "mocap_clip_idx", "mocap_time_u"):
                setattr(a, k, None)
            return
        d = {k: v.to(self.device).contiguous() for k, v in draws.items()}
        self._draws = d                                     # keep alive
        f, f64, i32 = torch.float32, torch.float64, torch.int32
        a.noise_u = p(d["noise_u"], f)
        a.rs_eps_u, a.rs_c_idx, a.rs_cmd_u = p(d["rs_eps_u"], f64), p(d["rs_c_idx"], i32), p(d["rs_cmd_u"], f)
        a.rt_eps_u, a.rt_c_idx, a.rt_cmd_u = p(d["rt_eps_u"], f64), p(d["rt_c_idx"], i32), p(d["rt_cmd_u"], f)
        a.push_u = p(d["push_u"], f)
        a.mocap_clip_idx, a.mocap_time_u = p(d["mocap_clip_idx"], i32), p(d["mocap_time_u"], f64)

    def load_state(self, snap: Dict[str, torch.Tensor]) -> None:
        """Overwrite the env's carried buffers from a snapshot dict (qa_b200.synthetic.make_snapshot
        naming == the reference's attribute names).  Simulator tensors are NOT touched."""
        dev = self.device
        for k in ("actions", "last_actions", "torques_org", "last_torques_org", "last_dof_vel", "last_root_vel",
                  "action_history_buf", "obs_history_buf", "commands", "latent_eps", "latent_c", "feet_air_time",
                  "episode_length_buf", "last_contacts"):
            if k in snap:
                getattr(self, k).copy_(snap[k].to(dev))
        if "episode_sums" in snap:                                 # (14,N) name-major -> (N,16) env-major
            self._episode_sums.zero_()
            self._episode_sums[:, :K.NUM_REWARDS].copy_(snap["episode_sums"].to(dev).t())
        if "obs_disc_buf" in snap:
            self.obs_disc_buf.copy_(snap["obs_disc_buf"].to(dev))

    def use_device_step_counter(self, on: bool = True) -> None:
        """Switch the per-step scalars (Philox counter, push flag, ring head) to the device-resident counter, synced
        from the host's `common_step_counter`.  Required before capturing `post_physics_step` in a CUDA graph; the
        host counters keep advancing in Python but are only authoritative again after `sync_step_counter()`."""
        self.device_step_counter = on
        if on:
            self._step_state[0] = self.common_step_counter

    def sync_step_counter(self) -> None:
        """Host counters := device counter (one 8-byte D2H)."""
        self.common_step_counter = int(self._step_state[0].item())
        if self._ring_len:
            self._ring_head = (self.common_step_counter - 1) % self._ring_len

    def refresh_prior(self) -> None:
        """`_resample_latent_c` re-derives `prior_prob = softmax(prior_parameters / T)` on every call (:536-538).  Here the
        mode draw happens inside the fused step, which reads the inclusive CDF from a device tensor at a fixed address:
        three tiny torch kernels, no host sync, visible to captured graphs.  The trainer calls this after every
        discriminator update (the only writer of `prior_parameters`, gail.py:462-464)."""
        with torch.no_grad():        # float64 like ops.bbc_const / oracle/philox.prior_cdf, rounded to fp32 once at the end
            prob = torch.softmax(self.prior_parameters.to(torch.float64) / self.cfg.latent_c_temperature, dim=-1)
            self.prior_prob.copy_(prob)
            self._prior_cdf.copy_(torch.cumsum(prob, dim=0))

    def set_prior_parameters(self, prior: torch.Tensor) -> None:
        """Overwrite `env.prior_parameters` IN PLACE (the trainer and captured graphs hold the tensor) and refresh the CDF."""
        self.prior_parameters.copy_(prior.to(self.device))
        self.refresh_prior()

    # ---- VecEnv API ---------------------------------------------------------------------------------
    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def get_disc_observations(self):
        return self.obs_disc_buf

    def _pre_physics(self, actions: torch.Tensor) -> None:
        cfg = self.cfg
        if cfg.action_delay and self.global_counter % cfg.delay_update_global_steps == 0 and self._delay_schedule:
            self.delay = int(self._delay_schedule.pop(0))                      # :88-91
        self.global_counter += 1
        ops.action_push(actions.contiguous(), self.action_history_buf, self.actions,
                        self.delay if cfg.action_delay else 0, cfg.clip_actions / cfg.action_scale)
        for _ in range(cfg.decimation):                                         # :101-106
            ops.pd_torques(self.actions, self.physics.dof_state, self.motor_strength, self.p_gains, self.d_gains,
                           self.default_dof_pos, self.torque_limits, self.torques, self.torques_org,
                           cfg.action_scale, cfg.hip_scale_reduction)
            self.physics.set_dof_actuation_force(self.torques)
            self.physics.simulate()

    def post_physics_step(self, refresh: bool = True) -> None:
        """K2 + reset compaction, all on the current stream, no host sync."""
        ph, a, cfg = self.physics, self._args, self.cfg
        if refresh:
            ph.refresh()
        self.common_step_counter += 1
        do_push = bool(cfg.push_robots and (self.common_step_counter % cfg.push_interval == 0))
        prev_disc = self._disc[self._pp]
        self._pp ^= 1
        self.obs_buf, self.privileged_obs_buf = self._obs[self._pp], self._priv[self._pp]
        self.obs_disc_buf = self._disc[self._pp]
        if self._ring_len:
            self._ring_head = (self._ring_head + 1) % self._ring_len
        a.do_push = int(do_push)
        a.contact_ring_head = max(self._ring_head, 0)
        a.rng_step = self.common_step_counter
        a.step_state = self._step_state.data_ptr() if self.device_step_counter else None
        a.root_states, a.dof_state = ph.root_states.data_ptr(), ph.dof_state.data_ptr()
        a.rigid_body_state, a.contact_forces = ph.rigid_body_state.data_ptr(), ph.contact_forces.data_ptr()
        a.obs_buf, a.privileged_obs_buf = self.obs_buf.data_ptr(), self.privileged_obs_buf.data_ptr()
        a.obs_disc_buf = self.obs_disc_buf.data_ptr()
        if self.k2_events is not None:                       # bench.py: per-launch CUDA events on this stream
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.post_physics_bbc(self._const, a)
            e1.record()
            self.k2_events.append((e0, e1))
        else:
            ops.post_physics_bbc(self._const, a)
        ops.compact_resets(self.reset_buf, prev_disc, self._reset_ids, self._reset_ids_i32, self._terminal_disc,
                           self._reset_count)
        ph.set_states_indexed(self._reset_ids_i32, self._reset_count)
        if do_push:
            ph.set_root_states_all()

    def _k2_only_step(self) -> None:
        """The fused post-physics launch alone (bench roofline leg)."""
        ph, a = self.physics, self._args
        ph.refresh()
        self._pp ^= 1
        self.obs_buf, self.privileged_obs_buf = self._obs[self._pp], self._priv[self._pp]
        self.obs_disc_buf = self._disc[self._pp]
        a.step_state = self._step_state.data_ptr() if self.device_step_counter else None
        a.root_states, a.dof_state = ph.root_states.data_ptr(), ph.dof_state.data_ptr()
        a.rigid_body_state, a.contact_forces = ph.rigid_body_state.data_ptr(), ph.contact_forces.data_ptr()
        a.obs_buf, a.privileged_obs_buf = self.obs_buf.data_ptr(), self.privileged_obs_buf.data_ptr()
        a.obs_disc_buf = self.obs_disc_buf.data_ptr()
        ops.post_physics_bbc(self._const, a)

    def step_device(self, actions: torch.Tensor):
        """Sync-free step.  Returns (obs, priv_obs, rew, reset, reset_ids_padded, count, terminal_padded);
        the first `count` entries / rows of the padded tensors are valid."""
        self._pre_physics(actions)
        self.post_physics_step()
        return (self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self._reset_ids,
                self._reset_count, self._terminal_disc)

    def step(self, actions: torch.Tensor):
        """Reference signature (:78-115): 7-tuple with variable-length reset ids (costs one 4-byte D2H)."""
        self._pre_physics(actions)
        self.post_physics_step()
        self._count_host.copy_(self._reset_count, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        k = int(self._count_host[0])
        if k > 0:                                                                # :188-189, :230-240
            means = self._episode_rew_means
            self.extras["episode"] = {"rew_" + n: means[i] for i, n in enumerate(K.REWARD_NAMES)}
            if self.cfg.send_timeouts:
                self.extras["time_outs"] = self._time_outs_latched
        return (self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras,
                self._reset_ids[:k], self._terminal_disc[:k])

    def reset(self):
        """Reset all robots (:67-76): `reset_idx(all envs)` followed by one step with zero actions.

        `reset_idx` is the reset branch of the fused step, so the first half runs K2 once with every env past its episode
        limit and then puts back what `reset_idx` itself does not touch but a full step does (:178-240 vs :124-166): the step
        counters and push schedule, `last_contacts` and the contact rings, the latched `extras["time_outs"]` (the time-out mask
        from BEFORE the reset, :239-240), `extras["episode"]` (means of the episode sums as they were, without the phantom
        step's reward, :230-234) and the zeroed `last_*` buffers (:217-221).  Parity: tests/test_env_gpu.py against the
        reference's own `reset()` (tests/golden/bbc_env_reset_n64.npz)."""
        dev_counter = self.device_step_counter
        if dev_counter:
            self.sync_step_counter()
            self.device_step_counter = False
        counter, ring_head = self.common_step_counter, self._ring_head
        keep = [t.clone() for t in (self.last_contacts, self.time_out_buf, self.obs_disc_buf)]
        rings = [r.clone() for r in (self._contact_ring, self._contact_force_ring)] if self._ring_len else []
        means = self._episode_sums[:, :K.NUM_REWARDS].mean(dim=0) / self.max_episode_length_s
        # a counter value that neither pushes nor resamples in this pass and keys a Philox stream no real step uses
        fake = counter + (1 << 40)
        if self.cfg.push_robots and (fake + 1) % self.cfg.push_interval == 0:
            fake += 1
        self.common_step_counter = fake
        over = int(self.max_episode_length) + 1
        if (over + 1) % self.cfg.resample_period == 0:
            over += 1
        self.episode_length_buf.fill_(over)
        self.post_physics_step(refresh=False)
        self.common_step_counter, self._ring_head = counter, ring_head
        self.last_contacts.copy_(keep[0])
        self._time_outs_latched.copy_(keep[1])
        self.obs_disc_buf.copy_(keep[2])                    # the next step's "previous disc obs" (terminal states, :153-154)
        for ring, saved in zip((self._contact_ring, self._contact_force_ring), rings):
            ring.copy_(saved)
        self._episode_rew_means.copy_(means)
        for t in (self.last_actions, self.last_dof_vel, self.last_root_vel, self.last_torques_org):
            t.zero_()
        self.extras["episode"] = {"rew_" + n: self._episode_rew_means[i] for i, n in enumerate(K.REWARD_NAMES)}
        if self.cfg.send_timeouts:
            self.extras["time_outs"] = self._time_outs_latched
        # ... then the reference steps once with zero actions and returns that observation
        obs, priv, *_ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device))
        if dev_counter:
            self.use_device_step_counter(True)
        return obs, priv


from .rsl_rl.vec_env import VecEnv  # noqa: E402  (bottom of the module: rsl_rl imports nothing from here)

VecEnv.register(LeggedRobot)

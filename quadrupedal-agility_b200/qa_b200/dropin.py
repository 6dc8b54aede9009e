"""Drop-in construction: the reference builds its env as `task_class(cfg, sim_params, physics_engine, sim_device, headless)`
(bbc/legged_gym/utils/task_registry.py:36-73).  Everything that constructor does is simulator side -- IsaacGym sim / terrain /
asset / actor creation (`_create_envs`, bbc/legged_gym/envs/base/legged_robot.py:743-859) and the buffers `_init_buffers`
derives from the asset (:1053-1074) -- and stays the reference's.  This module takes the env the reference has built and hands
the hot path to this package:

  config_from_reference(ref_env)   the flat `BbcEnvConfig` out of the reference's nested config classes + the asset's index lists
  static_from_reference(ref_env)   the per-env constants (`static` of `qa_b200.legged_robot.LeggedRobot`) under the reference's
                                   attribute names
  mocap_from_reference(ref_env)    the `MocapTable` out of the reference `MotionLoader`'s trajectories
  from_reference_env(ref_env)      -> `qa_b200.legged_robot.LeggedRobot` over `IsaacGymPhysics(ref_env.gym, ref_env.sim)`
  make_task_class(RefLeggedRobot)  a class with the REFERENCE's constructor signature, registrable with the reference's
                                   `task_registry` in place of `LeggedRobot`: train.py runs unchanged

    from legged_gym.envs.base.legged_robot import LeggedRobot as RefLeggedRobot
    from qa_b200.dropin import make_task_class
    task_registry.register("go2_locomotion", make_task_class(RefLeggedRobot), Go2LocomotionCfg(), Go2LocomotionCfgAlgo())

`oracle/check_dropin.py` (build container) runs the extraction over the reference env the parity harness builds and
checks the round trip against the values that were injected.
"""
from typing import Dict, Optional

import torch

from . import config as K
from .config import BbcEnvConfig

STATIC_KEYS = ("motor_strength", "mass_params_tensor", "friction_coeffs_tensor", "env_origins", "height_samples", "height_points",
               "default_dof_pos", "p_gains", "d_gains", "torque_limits", "dof_vel_limits", "dof_pos_limits", "noise_scale_vec",
               "prior_parameters")


def _lst(t):
    return [int(v) for v in (t.tolist() if torch.is_tensor(t) else t)]


def config_from_reference(ref_env) -> BbcEnvConfig:
    """`ref_env.cfg` is the reference's `Go2LocomotionCfg` instance (go2_locomotion_config.py:8-181 over
    legged_robot_config.py); index lists come from the asset lookups of `_create_envs` (:1024-1051)."""
    c = ref_env.cfg
    ctl, dr, rw, cm, nz, nm, tr = c.control, c.domain_rand, c.rewards, c.commands, c.noise, c.normalization, c.terrain
    stiff = ctl.stiffness["joint"] if isinstance(ctl.stiffness, dict) else float(ctl.stiffness)
    damp = ctl.damping["joint"] if isinstance(ctl.damping, dict) else float(ctl.damping)
    kw = dict(
        num_envs=int(ref_env.num_envs), num_bodies=int(ref_env.num_bodies),
        feet_indices=_lst(ref_env.feet_indices), penalised_contact_indices=_lst(ref_env.penalised_contact_indices),
        termination_contact_indices=_lst(ref_env.termination_contact_indices), hip_indices=_lst(ref_env.hip_indices),
        sim_dt=float(ref_env.sim_params.dt), decimation=int(ctl.decimation), action_scale=float(ctl.action_scale),
        hip_scale_reduction=float(ctl.hip_scale_reduction), stiffness=float(stiff), damping=float(damp),
        clip_actions=float(nm.clip_actions), clip_observations=float(nm.clip_observations),
        default_dof_pos=[float(v) for v in ref_env.default_dof_pos.reshape(-1).tolist()],
        dof_vel_limits=[float(v) for v in ref_env.dof_vel_limits.tolist()],
        torque_limits=[float(v) for v in ref_env.torque_limits.tolist()],
        soft_dof_pos_limit=float(rw.soft_dof_pos_limit), soft_dof_vel_limit=float(rw.soft_dof_vel_limit),
        soft_torque_limit=float(rw.soft_torque_limit),
        episode_length_s=float(c.env.episode_length_s), resampling_time=float(cm.resampling_time),
        push_interval_s=float(dr.push_interval_s), max_push_vel_xy=float(dr.max_push_vel_xy), push_robots=bool(dr.push_robots),
        tracking_sigma=float(rw.tracking_sigma), jump_goal=float(rw.jump_goal), only_positive_rewards=bool(rw.only_positive_rewards),
        lin_vel_x=[list(map(float, r)) for r in cm.ranges.lin_vel_x], lin_vel_y=[list(map(float, r)) for r in cm.ranges.lin_vel_y],
        ang_vel_yaw=[list(map(float, r)) for r in cm.ranges.ang_vel_yaw], jump_height=list(map(float, cm.ranges.jump_height)),
        locomotion_height=list(map(float, cm.ranges.locomotion_height)),
        lin_vel_x_clip=float(cm.lin_vel_x_clip), lin_vel_y_clip=float(cm.lin_vel_y_clip), ang_vel_yaw_clip=float(cm.ang_vel_yaw_clip),
        s_lin_vel=float(nm.obs_scales.lin_vel), s_ang_vel=float(nm.obs_scales.ang_vel), s_dof_pos=float(nm.obs_scales.dof_pos),
        s_dof_vel=float(nm.obs_scales.dof_vel), s_key_pos=float(nm.obs_scales.key_pos), s_foot_contact=float(nm.obs_scales.foot_contact),
        s_lin_vel_dist=float(nm.obs_scales.lin_vel_dist), s_ang_vel_dist=float(nm.obs_scales.ang_vel_dist),
        add_noise=bool(nz.add_noise), noise_level=float(nz.noise_level), n_roll_pitch=float(nz.noise_scales.roll_pitch),
        n_dof_pos=float(nz.noise_scales.dof_pos), n_dof_vel=float(nz.noise_scales.dof_vel), n_lin_vel=float(nz.noise_scales.lin_vel),
        n_ang_vel=float(nz.noise_scales.ang_vel), root_height_obs=bool(c.env.root_height_obs),
        measure_heights=bool(tr.measure_heights), border_size=float(tr.border_size), horizontal_scale=float(tr.horizontal_scale),
        vertical_scale=float(tr.vertical_scale), measured_points_x=list(map(float, tr.measured_points_x)),
        measured_points_y=list(map(float, tr.measured_points_y)),
        action_delay=bool(dr.action_delay), delay_update_global_steps=int(dr.delay_update_global_steps),
        action_curr_step=list(map(int, dr.action_curr_step)),
        task_obs_weight_decay=bool(nm.task_obs_weight_decay), task_obs_weight_decay_steps=int(nm.task_obs_weight_decay_steps),
        recovery_init_prob=float(c.env.recovery_init_prob), send_timeouts=bool(getattr(c.env, "send_timeouts", True)))
    cfg = BbcEnvConfig(**kw)
    # what the kernels hard-wire must be what the reference env was built with: refuse a config this library cannot serve
    # instead of running a different task
    want = dict(num_dof=K.NUM_DOF, num_obs=K.NUM_OBS, num_obs_disc=K.NUM_OBS_DISC, history_len=K.HISTORY_LEN, dim_c=K.DIM_C)
    got = dict(num_dof=int(ref_env.num_dof), num_obs=int(c.env.num_obs), num_obs_disc=int(c.env.num_obs_disc),
               history_len=int(c.env.history_len), dim_c=len(c.env.mocap_category_all))
    if want != got:
        raise ValueError(f"reference env layout {got} differs from the layout libqa_b200 is built for {want}")
    ref_scales = getattr(ref_env, "reward_scales", None)
    if ref_scales is not None:                     # `_prepare_reward_function` (:917-933) has multiplied them by dt and dropped the zeros
        mine = dict(zip(K.REWARD_NAMES, cfg.reward_scales_dt()))
        for name, v in ref_scales.items():
            if name == "termination":
                continue
            if name not in mine or abs(mine[name] - float(v)) > 1e-12 * max(1.0, abs(float(v))):
                raise ValueError(f"reward scale '{name}' = {v} of the reference env is not the one libqa_b200 computes ({mine.get(name)})")
    return cfg


def static_from_reference(ref_env) -> Dict[str, torch.Tensor]:
    """Per-env constants under the names `_create_envs` / `_init_buffers` give them (:796-859, :1051-1076)."""
    st = {}
    for k in STATIC_KEYS:
        v = getattr(ref_env, k)
        st[k] = v.detach().clone() if torch.is_tensor(v) else torch.as_tensor(v)
    hp = st["height_points"]
    if hp.dim() == 3:                                # (N, P, 3), identical rows (:1176-1188): the kernels take one copy
        st["height_points"] = hp[0].clone()
    st["height_samples"] = st["height_samples"].to(torch.int16)
    return st


def mocap_from_reference(ref_env):
    """The labelled clips of the reference `MotionLoader` (bbc/rsl_rl/datasets/motion_loader.py:152-249) as one device table."""
    from .mocap import MocapTable
    ml = ref_env.motion_loader
    files = list(getattr(ml, "motion_files_lb", None) or ref_env.cfg.env.motion_files_lb)
    return MocapTable.from_json_files(files)          # the loader's own clip order (motion_loader.py:111): clip ids stay the same


def from_reference_env(ref_env, device: Optional[str] = None, physics=None, mocap=None, seed: Optional[int] = None, **kw):
    """The reference env, once built (IsaacGym sim created, actors spawned, buffers initialised), replaced on the hot path:
    returns a `qa_b200.legged_robot.LeggedRobot` that steps the SAME simulator through `IsaacGymPhysics`."""
    from .isaacgym_backend import IsaacGymPhysics
    from .legged_robot import LeggedRobot
    cfg = config_from_reference(ref_env)
    static = static_from_reference(ref_env)
    device = device or str(ref_env.device)
    if physics is None:
        physics = IsaacGymPhysics(ref_env.gym, ref_env.sim, cfg.num_envs)
    if mocap is None:
        mocap = mocap_from_reference(ref_env)
    env = LeggedRobot(cfg, physics, static, mocap, device=device, seed=1 if seed is None else seed, **kw)
    env.reference_env = ref_env                      # viewer, terrain, gym handles: still the reference's
    return env


def make_task_class(reference_class):
    """A class with the reference's constructor signature (task_registry.py:66-70) whose instances ARE this package's env."""

    class LeggedRobotB200:
        REFERENCE_CLASS = reference_class

        def __new__(cls, cfg, sim_params, physics_engine, sim_device, headless):
            ref_env = cls.REFERENCE_CLASS(cfg=cfg, sim_params=sim_params, physics_engine=physics_engine, sim_device=sim_device,
                                          headless=headless)
            return from_reference_env(ref_env, device=sim_device)

    LeggedRobotB200.__name__ = f"{reference_class.__name__}B200"
    return LeggedRobotB200

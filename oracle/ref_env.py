"""Drive the UNMODIFIED reference `LeggedRobot` (bbc) on injected synthetic state.

TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  `LeggedRobot.__new__`
+ attribute injection replaces `__init__` (which needs IsaacGym); every arithmetic method
that runs is the reference's own.  The only interposition is on the *random sources*: the
reference draws on variable-length index sets from numpy / torch / multinomial, here those
calls return the matching entries of the dense per-env draw arrays that the oracle and the
CUDA kernels consume ("RNG parity").
"""
import copy
import glob
import os
import types

import numpy as np
import torch

from ref_harness import import_reference, REFERENCE_ROOT


class _Ctx:
    env_ids = None
    site = "rs"
    k = 0
    draws = None
    choice_calls = 0
    latent_c_idx = None


CTX = _Ctx()


class _NpRandomProxy:
    """numpy.random replacement seen by legged_robot.py / motion_loader.py."""

    def rand(self, *shape):                               # legged_robot.py:533
        u = CTX.draws[f"{CTX.site}_eps_u"][CTX.env_ids].numpy().astype(np.float64)
        return u.reshape(shape)

    def random(self):                                     # legged_robot.py:208 (recovery_init)
        return 0.5

    def choice(self, a, size=None, p=None, replace=True):   # motion_loader.py:314-320
        mode = CTX.choice_calls
        CTX.choice_calls += 1
        sel = (CTX.latent_c_idx == mode).cpu().numpy()
        ids = CTX.env_ids.cpu().numpy()[sel]
        out = CTX.draws["mocap_clip_idx"][ids].numpy().astype(np.int64)
        assert len(out) == size
        assert all(int(x) in set(np.asarray(a).tolist()) for x in out)
        return out

    def uniform(self, low=0.0, high=1.0, size=None):        # motion_loader.py:336-337
        return CTX.draws["mocap_time_u"][CTX.env_ids].numpy().astype(np.float64)


class _ModProxy:
    def __init__(self, real, **over):
        self._real = real
        self._over = over

    def __getattr__(self, name):
        if name in self._over:
            return self._over[name]
        return getattr(self._real, name)


class _PriorProb(torch.Tensor):
    def multinomial(self, n, replacement=True):            # legged_robot.py:539
        return CTX.draws[f"{CTX.site}_c_idx"][CTX.env_ids].long()


def _softmax_proxy(real_F):
    def softmax(x, dim=-1):
        return real_F.softmax(x, dim=dim).as_subclass(_PriorProb)
    return softmax


def build_reference_env(cfg, static, snap, files_lb):
    """cfg: qa_b200.config.BbcEnvConfig (only used for sizes / index lists)."""
    ref = import_reference("bbc")
    LR = ref.legged_robot
    real_torch, real_np, real_F = torch, np, LR.F if not isinstance(LR.F, _ModProxy) else LR.F._real

    def rand_floats(lower, upper, shape, device):          # torch_jit_utils.py:111-114, rand injected
        u = CTX.draws[f"{CTX.site}_cmd_u"][CTX.env_ids, CTX.k:CTX.k + 1]
        CTX.k += 1
        return (upper - lower) * u + lower

    def rand_float(lower, upper, shape, device):           # isaacgym.torch_utils.torch_rand_float
        if CTX.site == "push":
            u = CTX.draws["push_u"]
        else:
            u = CTX.draws[f"{CTX.site}_cmd_u"][CTX.env_ids, CTX.k:CTX.k + 1]
            CTX.k += 1
        return (upper - lower) * u + lower

    LR.torch_rand_floats = rand_floats
    LR.torch_rand_float = rand_float
    LR.np = _ModProxy(real_np, random=_NpRandomProxy())
    LR.F = _ModProxy(real_F, softmax=_softmax_proxy(real_F))
    LR.torch = _ModProxy(real_torch, rand_like=lambda t: CTX.draws["noise_u"].clone())
    ML = ref.motion_loader
    ML.np = _ModProxy(real_np, random=_NpRandomProxy())

    class RefEnv(ref.LeggedRobot):
        """Adds bookkeeping of WHICH envs a random call serves; no arithmetic of its own."""

        def _post_physics_step_callback(self):
            CTX.site = "rs"
            super()._post_physics_step_callback()

        def _resample_latent_eps(self, env_ids):
            CTX.env_ids = env_ids.cpu()
            super()._resample_latent_eps(env_ids)

        def _resample_latent_c(self, env_ids, temperature=0.25):
            CTX.env_ids = env_ids.cpu()
            super()._resample_latent_c(env_ids, temperature)

        def _resample_commands(self, env_ids):
            CTX.env_ids = env_ids.cpu()
            CTX.k = 0
            super()._resample_commands(env_ids)

        def _push_robots(self):
            CTX.site = "push"
            super()._push_robots()

        def reset_idx(self, env_ids):
            CTX.site = "rt"
            CTX.env_ids = env_ids.cpu()
            super().reset_idx(env_ids)

    class RefLoader(ML.MotionLoader):
        def get_full_frame_batch(self, num_frames, latent_c_idx=None):
            CTX.choice_calls = 0
            CTX.latent_c_idx = latent_c_idx.cpu()
            return super().get_full_frame_batch(num_frames, latent_c_idx)

    N, B = cfg.num_envs, cfg.num_bodies
    env = RefEnv.__new__(RefEnv)
    rcfg = copy.deepcopy(ref.envs.Go2LocomotionCfg())
    env.cfg = rcfg
    env.sim_params = types.SimpleNamespace(dt=cfg.sim_dt)
    env.device = "cpu"
    env.num_envs = N
    env.num_dof = 12
    env.num_bodies = B
    env.num_actions = 12
    env.mocap_category = rcfg.env.mocap_category
    env.mocap_category_all = rcfg.env.mocap_category_all
    env.num_mocap = len(env.mocap_category)
    env.dim_c = len(env.mocap_category_all)
    env._parse_cfg()
    from isaacgym import gymapi
    env.gym = gymapi.acquire_gym()
    env.sim = None
    env.viewer = None
    env.enable_viewer_sync = False
    env.debug_viz = False
    env.headless = True
    env.up_axis_idx = 2
    env.init_done = True
    env.common_step_counter = 0
    env.global_counter = 0
    env.extras = {}

    t = lambda x: x.clone()                                 # noqa: E731
    env.root_states = t(snap["root_states"])
    env.dof_state = t(snap["dof_state"])
    env.rigid_body_state = t(snap["rigid_body_state"])
    env.dof_pos = env.dof_state.view(N, 12, 2)[..., 0]
    env.dof_vel = env.dof_state.view(N, 12, 2)[..., 1]
    env.base_quat = env.root_states[:, 3:7]
    env.rigid_body_pos = env.rigid_body_state.view(N, B, 13)[..., 0:3]
    env.contact_forces = t(snap["contact_forces"])
    env.noise_scale_vec = env._get_noise_scale_vec(rcfg)
    env.gravity_vec = torch.tensor([0., 0., -1.]).repeat(N, 1)
    env.forward_vec = torch.tensor([1., 0., 0.]).repeat(N, 1)
    env.torques = torch.zeros(N, 12)
    env.torques_org = t(snap["torques_org"])
    env.p_gains = t(static["p_gains"])
    env.d_gains = t(static["d_gains"])
    env.actions = t(snap["actions"])
    env.last_actions = t(snap["last_actions"])
    env.last_dof_vel = t(snap["last_dof_vel"])
    env.last_root_vel = t(snap["last_root_vel"])
    env.last_torques_org = t(snap["last_torques_org"])
    env.action_history_buf = t(snap["action_history_buf"])
    env.motor_strength = t(static["motor_strength"])
    env.obs_history_buf = t(snap["obs_history_buf"])
    env.contact_buf = torch.zeros(N, rcfg.env.contact_buf_len, 4)
    env.contact_force_buf = torch.zeros(N, rcfg.env.contact_force_buf_len, 4)
    env.commands = t(snap["commands"])
    env.latent_eps = t(snap["latent_eps"])
    env.latent_c = t(snap["latent_c"])
    env.prior_parameters = t(static["prior_parameters"])
    env.prior_prob = t(static["prior_parameters"])
    env.feet_indices = torch.tensor(cfg.feet_indices, dtype=torch.long)
    env.penalised_contact_indices = torch.tensor(cfg.penalised_contact_indices, dtype=torch.long)
    env.termination_contact_indices = torch.tensor(cfg.termination_contact_indices, dtype=torch.long)
    env.hip_indices = torch.tensor(cfg.hip_indices, dtype=torch.long)
    env.key_body_ids = torch.tensor(cfg.feet_indices, dtype=torch.long)
    env.feet_air_time = t(snap["feet_air_time"])
    env.last_contacts = t(snap["last_contacts"])
    env.jump_goal = torch.zeros(N, dtype=torch.bool)
    env.base_lin_vel = torch.zeros(N, 3)
    env.base_ang_vel = torch.zeros(N, 3)
    env.projected_gravity = torch.zeros(N, 3)
    env.height_points = static["height_points"].unsqueeze(0).repeat(N, 1, 1)
    env.num_height_points = static["height_points"].shape[0]
    env.measured_heights = 0
    env.height_samples = t(static["height_samples"])
    env.terrain = types.SimpleNamespace(cfg=rcfg.terrain)
    env.default_dof_pos = t(static["default_dof_pos"])
    env.dof_pos_limits = t(static["dof_pos_limits"])
    env.dof_vel_limits = t(static["dof_vel_limits"])
    env.torque_limits = t(static["torque_limits"])
    env.mass_params_tensor = t(static["mass_params_tensor"])
    env.friction_coeffs_tensor = t(static["friction_coeffs_tensor"])
    env.env_origins = t(static["env_origins"])
    env.custom_origins = True
    env.obs_buf = torch.zeros(N, rcfg.env.num_obs)
    env.obs_disc_buf = t(snap["obs_disc_buf"])
    env.privileged_obs_buf = torch.zeros(N, rcfg.env.num_obs)
    env.rew_buf = torch.zeros(N)
    env.reset_buf = torch.ones(N, dtype=torch.long)
    env.episode_length_buf = t(snap["episode_length_buf"])
    env.time_out_buf = torch.zeros(N, dtype=torch.bool)
    env.task_obs_weight_decay = rcfg.normalization.task_obs_weight_decay
    env.task_obs_weight = 1.0
    env._prepare_reward_function()
    for k, name in enumerate(env.episode_sums.keys()):
        env.episode_sums[name] = t(snap["episode_sums"][k])
    env.motion_loader = RefLoader(motion_files_lb=files_lb, motion_files_ulb=[],
                                  mocap_category=env.mocap_category, time_between_frames=env.dt,
                                  mocap_state_init=True, device="cpu") if files_lb is not None else None
    return ref, env


def labelled_clip_files():
    return sorted(glob.glob(os.path.join(REFERENCE_ROOT, "bbc", "mocap_data", "mocap_all_lb", "*.json")))

"""Constants of the BBC `go2_locomotion` task that feed the fused kernels.

Every number is the shipped value of the reference configuration; the citation after
each group is the file:line in /root/reference it restates.  Kept as a plain dataclass
(not the reference's nested-class config) because the kernels need a flat POD.
"""
import math
import types
from dataclasses import dataclass, field
from typing import List, Tuple

# ---- dimensions (bbc/legged_gym/envs/go2/go2_locomotion_config.py:9-32) -------------------
NUM_DOF = 12
NUM_ACTIONS = 12
NUM_PROP = 57
NUM_EXPLICIT = 4
NUM_LATENT = 29
NUM_COMMAND = 11            # 5 commands + 1 epsilon + 5 behaviour modes
NUM_OBS = NUM_PROP + NUM_EXPLICIT + NUM_LATENT + NUM_COMMAND        # 101
HISTORY_LEN = 10
NUM_HIST = HISTORY_LEN * NUM_PROP                                    # 570
OBS_WIDTH = NUM_OBS + NUM_HIST                                       # 671 (stored obs row)
NUM_OBS_DISC = 49
DISC_OBS_LEN = 2
DIM_C = 5
NUM_COMMANDS = 5
ACTION_BUF_LEN = 8          # go2_locomotion_config.py:98
CONTACT_BUF_LEN = 100       # go2_locomotion_config.py:31-32
MOCAP_FRAME_WIDTH = 49      # motion_loader.py:133-135 (JOINT_VEL_END_IDX)

# alphabetical (dir()) order of the non-zero reward scales,
# legged_robot.py:922-946 + helpers.py:12-27, scales at go2_locomotion_config.py:137-163
REWARD_NAMES: Tuple[str, ...] = (
    "action_rate", "collision", "delta_torques", "dof_acc", "dof_error", "dof_pos_limits",
    "dof_vel_limits", "hip_pos", "jump_up_height", "locomotion_height", "torque_limits",
    "torques", "tracking_ang_vel", "tracking_lin_vel",
)
REWARD_SCALES_RAW = {
    "action_rate": -0.1, "collision": -10.0, "delta_torques": -1.0e-7, "dof_acc": -2.5e-7,
    "dof_error": -0.1, "dof_pos_limits": -0.1, "dof_vel_limits": -0.1, "hip_pos": -0.5,
    "jump_up_height": 0.2, "locomotion_height": 0.1, "torque_limits": -0.03,
    "torques": -0.00001, "tracking_ang_vel": 1.5, "tracking_lin_vel": 2.0,
}
NUM_REWARDS = len(REWARD_NAMES)    # 14
EPISODE_SUMS_PITCH = 16            # (N,16) env-major, 64 B per env

MOCAP_CATEGORY = ("walk", "pace", "trot", "canter", "jump")

# Go2 bodies after URDF import with collapse_fixed_joints and 6 dont_collapse links
# (bbc/resources/robots/go2/urdf/go2.urdf:89,127,386,645,904,1163)
GO2_BODY_NAMES = (
    "base", "Head_upper", "Head_lower",
    "FL_hip", "FL_thigh", "FL_calf", "FL_foot",
    "FR_hip", "FR_thigh", "FR_calf", "FR_foot",
    "RL_hip", "RL_thigh", "RL_calf", "RL_foot",
    "RR_hip", "RR_thigh", "RR_calf", "RR_foot",
)


def _idx(sub):
    return [i for i, n in enumerate(GO2_BODY_NAMES) if sub in n]


@dataclass
class BbcEnvConfig:
    num_envs: int = 4096
    num_bodies: int = len(GO2_BODY_NAMES)                        # 19 (runtime parameter)
    feet_indices: List[int] = field(default_factory=lambda: _idx("foot"))
    # legged_robot.py:1034-1039 (names matched in cfg order: thigh then calf / base then hip)
    penalised_contact_indices: List[int] = field(default_factory=lambda: _idx("thigh") + _idx("calf"))
    termination_contact_indices: List[int] = field(default_factory=lambda: _idx("base") + _idx("hip"))
    hip_indices: List[int] = field(default_factory=lambda: [0, 3, 6, 9])

    # control (go2_locomotion_config.py:53-61, legged_robot.py:1139)
    sim_dt: float = 0.005
    decimation: int = 4
    action_scale: float = 0.25
    hip_scale_reduction: float = 0.5
    stiffness: float = 40.0
    damping: float = 1.0
    clip_actions: float = 100.0
    clip_observations: float = 100.0
    default_dof_pos: List[float] = field(default_factory=lambda: [0.0, 0.9, -1.8] * 4)
    # URDF limits (go2.urdf <limit> tags): hip, thigh(front/rear), calf
    dof_pos_lower: List[float] = field(default_factory=lambda: [-1.0472, -1.5708, -2.7227] * 2 + [-1.0472, -0.5236, -2.7227] * 2)
    dof_pos_upper: List[float] = field(default_factory=lambda: [1.0472, 3.4907, -0.83776] * 2 + [1.0472, 4.5379, -0.83776] * 2)
    dof_vel_limits: List[float] = field(default_factory=lambda: [30.1, 30.1, 20.07] * 4)
    torque_limits: List[float] = field(default_factory=lambda: [20.0, 20.0, 40.0] * 4)
    soft_dof_pos_limit: float = 0.9
    soft_dof_vel_limit: float = 1.0
    soft_torque_limit: float = 1.0

    # episode / resampling (go2_locomotion_config.py:25,82,169)
    episode_length_s: float = 20.0
    resampling_time: float = 6.0
    push_interval_s: float = 8.0
    max_push_vel_xy: float = 0.5
    push_robots: bool = True

    # rewards (legged_robot_config.py:129-135, go2_locomotion_config.py:132-135)
    tracking_sigma: float = 0.25
    jump_goal: float = 10.0
    only_positive_rewards: bool = True

    # commands (go2_locomotion_config.py:165-181), per mode ['walk','pace','trot','canter','jump']
    lin_vel_x: List[List[float]] = field(default_factory=lambda: [[0.0, 0.6], [0.5, 1.5], [0.5, 1.5], [0.8, 2.5], [0.8, 2.0]])
    lin_vel_y: List[List[float]] = field(default_factory=lambda: [[-0.15, 0.15], [-0.3, 0.3], [-0.3, 0.3], [-0.5, 0.5], [-0.3, 0.3]])
    ang_vel_yaw: List[List[float]] = field(default_factory=lambda: [[-1.0, 1.0], [-1.57, 1.57], [-1.57, 1.57], [-0.5, 0.5], [-0.5, 0.5]])
    jump_height: List[float] = field(default_factory=lambda: [0.45, 0.58])
    locomotion_height: List[float] = field(default_factory=lambda: [0.25, 0.34])
    lin_vel_x_clip: float = 0.1
    lin_vel_y_clip: float = 0.05
    ang_vel_yaw_clip: float = 0.05
    latent_c_temperature: float = 0.25            # legged_robot.py:536

    # observation scales / noise (go2_locomotion_config.py:102-127)
    s_lin_vel: float = 0.5
    s_ang_vel: float = 0.25
    s_dof_pos: float = 1.0
    s_dof_vel: float = 0.05
    s_key_pos: float = 1.0
    s_foot_contact: float = 1.0
    s_lin_vel_dist: float = 0.5
    s_ang_vel_dist: float = 0.25
    add_noise: bool = True
    noise_level: float = 1.0
    n_roll_pitch: float = 0.01
    n_dof_pos: float = 0.01
    n_dof_vel: float = 1.5
    n_lin_vel: float = 0.1
    n_ang_vel: float = 0.2
    root_height_obs: bool = True

    # terrain (legged_robot_config.py:19-32)
    measure_heights: bool = True
    border_size: float = 30.0
    horizontal_scale: float = 0.1
    vertical_scale: float = 0.005
    measured_points_x: List[float] = field(default_factory=lambda: [-0.8, -0.7, -0.6, -0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8])
    measured_points_y: List[float] = field(default_factory=lambda: [-0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5])

    # action delay (go2_locomotion_config.py:97-100)
    action_delay: bool = True
    delay_update_global_steps: int = 24 * 20000
    action_curr_step: List[int] = field(default_factory=lambda: [0, 1])

    # discriminator-obs decay (go2_locomotion_config.py:129-130)
    task_obs_weight_decay: bool = True
    task_obs_weight_decay_steps: int = 50000
    recovery_init_prob: float = 0.0
    send_timeouts: bool = True

    # ---- derived ------------------------------------------------------------------------
    @property
    def dt(self) -> float:                      # legged_robot.py:1139
        return self.decimation * self.sim_dt

    @property
    def max_episode_length(self) -> float:      # legged_robot.py:1148 (np.ceil -> float)
        return float(math.ceil(self.episode_length_s / self.dt))

    @property
    def resample_period(self) -> int:           # legged_robot.py:454
        return int(self.resampling_time / self.dt)

    @property
    def push_interval(self) -> float:           # legged_robot.py:1150
        return float(math.ceil(self.push_interval_s / self.dt))

    @property
    def num_height_points(self) -> int:
        return len(self.measured_points_x) * len(self.measured_points_y)

    @property
    def center_height_index(self) -> int:       # legged_robot.py:266 (shape[1] // 2 + 1)
        return self.num_height_points // 2 + 1

    @property
    def env(self):
        """`cfg.env.*` as the reference's trainer reads it (bbc/rsl_rl/runners/on_policy_runner.py:37-63, gail.py:68-87;
        values of go2_locomotion_config.py:9-32): lets the reference's own runner be constructed over this env."""
        return types.SimpleNamespace(
            num_envs=self.num_envs, num_prop=NUM_PROP, num_explicit=NUM_EXPLICIT, num_latent=NUM_LATENT, num_command=NUM_COMMAND,
            num_obs=NUM_PROP + NUM_EXPLICIT + NUM_LATENT + NUM_COMMAND, num_privileged_obs=NUM_PROP + NUM_EXPLICIT + NUM_LATENT + NUM_COMMAND,
            num_obs_disc=49, history_len=HISTORY_LEN, disc_history_len=2, disc_obs_len=DISC_OBS_LEN, obs_disc_weight_step=0.0,
            frame_duration_scale=1.0, root_height_obs=self.root_height_obs, send_timeouts=self.send_timeouts)

    def reward_scales_dt(self) -> List[float]:
        """scale_k * dt in REWARD_NAMES order, rounded the way python does (legged_robot.py:932)."""
        return [REWARD_SCALES_RAW[k] * self.dt for k in REWARD_NAMES]

    def soft_dof_pos_limits(self):
        """legged_robot.py:426-429, evaluated in fp32 like the reference (torch tensors)."""
        import torch
        lo = torch.tensor(self.dof_pos_lower, dtype=torch.float32)
        hi = torch.tensor(self.dof_pos_upper, dtype=torch.float32)
        m = (lo + hi) / 2
        r = hi - lo
        return torch.stack([m - 0.5 * r * self.soft_dof_pos_limit, m + 0.5 * r * self.soft_dof_pos_limit], dim=1)

    def noise_scale_vec(self):
        """legged_robot.py:721-740."""
        import torch
        v = torch.zeros(OBS_WIDTH, dtype=torch.float32)
        v[:2] = self.n_roll_pitch * self.noise_level
        v[2:5] = self.n_ang_vel * self.noise_level * self.s_ang_vel
        v[5:17] = self.n_dof_pos * self.noise_level * self.s_dof_pos
        v[17:29] = self.n_dof_vel * self.noise_level * self.s_dof_vel
        v[58:61] = self.n_lin_vel * self.noise_level * self.s_lin_vel
        return v


def bbc_train_cfg() -> dict:
    """`class_to_dict(Go2LocomotionCfgAlgo())` of the reference (go2_locomotion_config.py:184-244 over
    legged_robot_config.py:196-233): the dict `OnPolicyRunner(env, train_cfg, ...)` takes."""
    return {
        "seed": 1,
        "runner_class_name": "OnPolicyRunner",
        "policy": dict(init_noise_std=1.0, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128],
                       priv_encoder_dims=[64], activation="elu", train_with_estimated_latent=True),
        "algorithm": dict(value_loss_coef=5.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01,
                          num_learning_epochs=5, num_mini_batches=4, schedule="adaptive", gamma=0.99, lam=0.95,
                          desired_kl=0.01, max_grad_norm=1.0, lr_ac=1e-3, lr_disc=5e-4, lr_q=1e-3,
                          surrogate_loss_coef=2.0, bounds_loss_coef=0.0, disc_coef=1.0, disc_logit_reg=0.05,
                          disc_grad_penalty=0.1, disc_weight_decay=0.0001, disc_replay_buffer_size=1000000,
                          us_coef=1.0, ss_coef=1.0, prior_soft_coef=1e-3, info_max_coef=1.0, begin_rim=200,
                          disc_loss_function="MSELoss", priv_reg_coef_schedual=[0, 0.1, 1000, 2000],
                          priv_reg_coef_schedual_resume=[0, 0.1, 0, 1]),
        "estimator": dict(train_with_estimated_explicit=True, learning_rate=1.0e-4, hidden_dims=[128, 64]),
        "runner": dict(policy_class_name="ActorCritic", algorithm_class_name="SSInfoGAIL", num_steps_per_env=24,
                       max_iterations=500000, save_interval=100, experiment_name="go2_locomotion", run_name="", experiment_idx=0,
                       pre_trained_actor_path=[],
                       dagger_update_freq=20, motion_files_lb=[], motion_files_ulb=[], num_preload_transitions=200000,
                       reward_i_coef=1.0, reward_us_coef=0.01, reward_ss_coef=0.2, reward_t_coef=0.2,
                       disc_hidden_units=[512, 256], min_normalized_std=[0.05, 0.02, 0.05] * 4,
                       resume=False, load_run=-1, checkpoint=-1, resume_path=None),
    }


def tsc_train_cfg(use_camera: bool = False) -> dict:
    """`class_to_dict(Go2AgilityCfgPPO())` of the reference (tsc/legged_gym/envs/base/legged_robot_config.py:365-470 +
    go2_agility_config.py:52-60): the dict the TSC `OnPolicyRunner(env, train_cfg, ...)` takes.  `use_camera=True` is the
    student configuration (`--use_camera`): `depth_encoder.if_depth` follows `depth.use_camera` (:406-414)."""
    return {
        "depth_encoder": dict(if_depth=use_camera, depth_shape=(87, 58), buffer_len=2, hidden_dims=512, learning_rate=1.e-3,
                              learning_rate_byol=3.e-4, learning_rate_min=1.e-5, num_steps_per_env=24),
        "seed": 1,
        "runner_class_name": "OnPolicyRunner",
        "policy": dict(init_noise_std=1.0, continue_from_last_std=True, scan_encoder_dims=[128, 64, 32],
                       actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128], priv_encoder_dims=[64],
                       activation="elu", tanh_encoder_output=False, rnn_type="lstm", rnn_hidden_size=512, rnn_num_layers=1),
        "algorithm": dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01,
                          num_learning_epochs=5, num_mini_batches=4, learning_rate=5.e-4, schedule="adaptive", gamma=0.99,
                          lam=0.95, desired_kl=0.01, max_grad_norm=1.0, dagger_update_freq=20,
                          priv_reg_coef_schedual=[0, 0.1, 500, 1000], priv_reg_coef_schedual_resume=[0, 0.1, 0, 1]),
        "estimator": dict(train_with_estimated_states=True, learning_rate=1.e-4, hidden_dims=[128, 64], load_estimator_bbc=True,
                          priv_states_dim=4, num_prop=57, num_auxiliary=8, num_scan=132),
        "runner": dict(policy_class_name="ActorCritic", algorithm_class_name="PPO", num_steps_per_env=24, max_iterations=50000,
                       save_interval=100, experiment_name="agility", run_name="", disc_loss_function="MSELoss",
                       reward_i_coef=0.05, reward_us_coef=0.0, reward_ss_coef=0.0, reward_t_coef=2.0,
                       disc_hidden_units=[512, 256], bbc_path="weights/bbc/model.pt", resume=False, load_run=-1, checkpoint=-1,
                       resume_path=None),
    }

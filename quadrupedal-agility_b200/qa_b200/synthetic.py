"""Synthetic / recorded-state generator for the BBC go2_locomotion hot path.

IsaacGym (the physics backend) is not installable outside its own python-3.8 binary
distribution, so tests and `bench.py` drive the env with *synthetic state snapshots*
whose distributions follow SURVEY.md section 8(d).  One snapshot = the simulator-owned
tensors (`root_states`, `dof_state`, `rigid_body_state`, `contact_forces`) plus the env's
persistent buffers, named exactly like the reference attributes
(bbc/legged_gym/envs/base/legged_robot.py:743-859).

All draws come from a seeded CPU `torch.Generator`, so the CPU oracle and the CUDA path
see bit-identical inputs.
"""
import math
from typing import Dict, Optional

import torch

from . import config as C
from .config import BbcEnvConfig


def _rand(gen, *shape):
    return torch.rand(*shape, generator=gen, dtype=torch.float32)


def _randn(gen, *shape):
    return torch.randn(*shape, generator=gen, dtype=torch.float32)


def make_static(cfg: BbcEnvConfig, seed: int = 1234, terrain_cells: int = 1600) -> Dict[str, torch.Tensor]:
    """Per-env constants created once at env construction (legged_robot.py:796-859, 1051-1076)."""
    gen = torch.Generator().manual_seed(seed * 7919 + 1)
    N = cfg.num_envs
    st = {}
    # EASI motor strength (legged_robot.py:861-888, go2_locomotion_config.py:90-95)
    easi_mean = [1.270984856442925803e+00, 1.269402596100474012e+00, 8.637638584658215990e-01,
                 8.973783516018792872e-01, 7.804512147922660903e-01, 1.069519100829913416e+00]
    easi_var = [9.087216265313172864e-03, 6.342416661098186637e-03, 1.376369951477590226e-05,
                4.598280851616735464e-05, 5.266858327126125377e-06, 8.413655048485571975e-05]
    mean_p = torch.tensor([easi_mean[0], easi_mean[2], easi_mean[4]] * 4)
    mean_d = torch.tensor([easi_mean[1], easi_mean[3], easi_mean[5]] * 4)
    std_p = torch.tensor([easi_var[0], easi_var[2], easi_var[4]] * 4)
    std_d = torch.tensor([easi_var[1], easi_var[3], easi_var[5]] * 4)
    ms_p = mean_p + std_p * _randn(gen, N, 12)
    ms_d = mean_d + std_d * _randn(gen, N, 12)
    st["motor_strength"] = torch.stack([ms_p, ms_d], dim=0).contiguous()            # (2,N,12)
    mass = torch.empty(N, 4)
    mass[:, 0] = 1.5 * _rand(gen, N)
    mass[:, 1:] = 0.2 * _rand(gen, N, 3) - 0.1
    st["mass_params_tensor"] = mass
    st["friction_coeffs_tensor"] = (0.6 + 1.4 * _rand(gen, N, 1))
    origins = torch.zeros(N, 3)
    origins[:, :2] = torch.floor(_rand(gen, N, 2) * 10.0) * 10.0 + 5.0              # terrain tile centres
    st["env_origins"] = origins
    st["height_samples"] = torch.randint(-10, 11, (terrain_cells, terrain_cells), generator=gen,
                                         dtype=torch.int16)
    gx, gy = torch.meshgrid(torch.tensor(cfg.measured_points_x), torch.tensor(cfg.measured_points_y),
                            indexing="ij")
    pts = torch.zeros(cfg.num_height_points, 3)
    pts[:, 0] = gx.flatten()
    pts[:, 1] = gy.flatten()
    st["height_points"] = pts                                                       # (187,3), same for every env
    st["default_dof_pos"] = torch.tensor(cfg.default_dof_pos, dtype=torch.float32).unsqueeze(0)
    st["p_gains"] = torch.full((12,), cfg.stiffness, dtype=torch.float32)
    st["d_gains"] = torch.full((12,), cfg.damping, dtype=torch.float32)
    st["torque_limits"] = torch.tensor(cfg.torque_limits, dtype=torch.float32)
    st["dof_vel_limits"] = torch.tensor(cfg.dof_vel_limits, dtype=torch.float32)
    st["dof_pos_limits"] = cfg.soft_dof_pos_limits()
    st["noise_scale_vec"] = cfg.noise_scale_vec()
    st["prior_parameters"] = torch.ones(C.DIM_C) * (1.0 / C.DIM_C)
    return st


def _center_cell_margin(cfg, root_states):
    """Distance (in cells, float64) of the centre height-scan point to the nearest cell boundary."""
    q = root_states[:, 3:7].double()
    yaw_q = q.clone()
    yaw_q[:, :2] = 0
    yaw_q = yaw_q / yaw_q.norm(dim=-1, keepdim=True)
    idx = cfg.center_height_index
    ny = len(cfg.measured_points_y)
    px = cfg.measured_points_x[idx // ny]
    py = cfg.measured_points_y[idx % ny]
    ang = 2.0 * torch.atan2(yaw_q[:, 2], yaw_q[:, 3])
    wx = math.cos(0) * 0 + (torch.cos(ang) * px - torch.sin(ang) * py) + root_states[:, 0].double()
    wy = (torch.sin(ang) * px + torch.cos(ang) * py) + root_states[:, 1].double()
    fx = (wx + cfg.border_size) / cfg.horizontal_scale
    fy = (wy + cfg.border_size) / cfg.horizontal_scale
    mx = torch.minimum(fx - torch.floor(fx), torch.ceil(fx) - fx)
    my = torch.minimum(fy - torch.floor(fy), torch.ceil(fy) - fy)
    return torch.minimum(mx, my)


def make_snapshot(cfg: BbcEnvConfig, seed: int = 1234, step: int = 0,
                  reset_frac: float = 0.015, plant_frac: float = 0.003) -> Dict[str, torch.Tensor]:
    """One post-physics state snapshot (what IsaacGym's refresh_* calls would expose) plus
    the env's carried buffers.  Distributions: SURVEY.md section 8(d)."""
    gen = torch.Generator().manual_seed(seed * 1000003 + step * 101 + 17)
    N, B = cfg.num_envs, cfg.num_bodies
    s = {}
    root = torch.zeros(N, 13)
    root[:, 0:2] = -2.0 + 104.0 * _rand(gen, N, 2)
    root[:, 2] = 0.22 + 0.23 * _rand(gen, N)
    yaw = (2 * _rand(gen, N) - 1) * math.pi
    roll = 0.15 * _randn(gen, N)
    pitch = 0.15 * _randn(gen, N)
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    quat = torch.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                        sy * cr * cp - cy * sr * sp, cy * cr * cp + sy * sr * sp], dim=-1)
    root[:, 3:7] = quat / quat.norm(dim=-1, keepdim=True)
    root[:, 7:10] = 0.6 * _randn(gen, N, 3)
    root[:, 10:13] = 0.8 * _randn(gen, N, 3)
    # keep the centre height-scan sample away from knife-edge cell boundaries: torch-CPU and
    # torch-CUDA themselves disagree there by one ulp (-> a different int16 cell).
    for _ in range(8):
        bad = _center_cell_margin(cfg, root) < 2e-3
        if not bool(bad.any()):
            break
        root[bad, 0:2] += 0.013
    s["root_states"] = root

    q0 = torch.tensor(cfg.default_dof_pos)
    dof = torch.zeros(N, 12, 2)
    dof[..., 0] = q0 + 0.25 * _randn(gen, N, 12)
    dof[..., 1] = 3.0 * _randn(gen, N, 12)
    s["dof_state"] = dof.reshape(N * 12, 2).contiguous()

    rb = torch.zeros(N, B, 13)
    rb[:, :, 0:3] = root[:, None, 0:3] + 0.2 * _randn(gen, N, B, 3)
    rb[:, :, 2] = rb[:, :, 2].clamp(min=0.0)
    rb[:, :, 6] = 1.0
    s["rigid_body_state"] = rb.reshape(N * B, 13).contiguous()

    cf = torch.zeros(N, B, 3)
    feet = torch.tensor(cfg.feet_indices)
    fz = torch.relu(40.0 + 40.0 * _randn(gen, N, 4)) * (_rand(gen, N, 4) > 0.5).float()
    fxy = 5.0 * _randn(gen, N, 4, 2) * (fz > 0).float().unsqueeze(-1)
    cf[:, feet, 2] = fz
    cf[:, feet, 0:2] = fxy
    nonfoot = torch.tensor([i for i in range(B) if i not in cfg.feet_indices])
    hit = _rand(gen, N, len(nonfoot)) < (reset_frac / 5.0)
    mag = 1.0 + 49.0 * _rand(gen, N, len(nonfoot))
    dirv = _randn(gen, N, len(nonfoot), 3)
    dirv = dirv / dirv.norm(dim=-1, keepdim=True)
    cf[:, nonfoot] = dirv * (mag * hit.float()).unsqueeze(-1)
    s["contact_forces"] = cf

    s["actions"] = _randn(gen, N, 12)
    s["last_actions"] = s["actions"] + 0.3 * _randn(gen, N, 12)
    s["torques_org"] = 8.0 * _randn(gen, N, 12)
    s["last_torques_org"] = 8.0 * _randn(gen, N, 12)
    s["last_dof_vel"] = dof[..., 1] + 0.5 * _randn(gen, N, 12)
    s["last_root_vel"] = torch.zeros(N, 6)
    ah = _randn(gen, N, C.ACTION_BUF_LEN, 12)
    s["action_history_buf"] = ah
    ep = torch.randint(0, 1000, (N,), generator=gen, dtype=torch.int64)
    planted = _rand(gen, N)
    pf = plant_frac
    ep = torch.where(planted < pf, torch.full_like(ep, 1000), ep)                   # -> time-out after += 1
    ep = torch.where((planted >= pf) & (planted < 2 * pf), torch.full_like(ep, 299), ep)     # -> resample
    ep = torch.where((planted >= 2 * pf) & (planted < 3 * pf), torch.full_like(ep, 0), ep)   # -> ep_len == 1 (history fill)
    root[(planted >= 3 * pf) & (planted < 3.3 * pf), 2] = -7.0                      # fell off the world (:174)
    s["episode_length_buf"] = ep
    s["last_contacts"] = _rand(gen, N, 4) > 0.5

    c_idx = torch.randint(0, C.DIM_C, (N,), generator=gen)
    s["latent_c"] = torch.nn.functional.one_hot(c_idx, C.DIM_C).float()
    s["latent_eps"] = 2 * _rand(gen, N, 1) - 1
    cmd = torch.zeros(N, 5)
    lx = torch.tensor(cfg.lin_vel_x)[c_idx]
    ly = torch.tensor(cfg.lin_vel_y)[c_idx]
    lw = torch.tensor(cfg.ang_vel_yaw)[c_idx]
    cmd[:, 0] = lx[:, 0] + (lx[:, 1] - lx[:, 0]) * _rand(gen, N)
    cmd[:, 1] = ly[:, 0] + (ly[:, 1] - ly[:, 0]) * _rand(gen, N)
    cmd[:, 2] = lw[:, 0] + (lw[:, 1] - lw[:, 0]) * _rand(gen, N)
    jump = (c_idx == C.DIM_C - 1).float()
    cmd[:, 3] = (cfg.jump_height[0] + (cfg.jump_height[1] - cfg.jump_height[0]) * _rand(gen, N)) * jump
    cmd[:, 4] = (cfg.locomotion_height[0] + (cfg.locomotion_height[1] - cfg.locomotion_height[0]) * _rand(gen, N)) * (1 - jump)
    s["commands"] = cmd
    # 60 % of the envs roughly track their command (so that most clipped rewards are > 0)
    track = _rand(gen, N) < 0.6
    v_body = torch.stack([cmd[:, 0], cmd[:, 1], torch.zeros(N)], dim=-1) + 0.2 * _randn(gen, N, 3)
    qv, qw = root[:, 3:6], root[:, 6:7]
    tt = 2.0 * torch.cross(qv, v_body, dim=-1)
    v_world = v_body + qw * tt + torch.cross(qv, tt, dim=-1)
    root[track, 7:10] = v_world[track]
    root[track, 12] = (cmd[:, 2] + 0.2 * _randn(gen, N))[track]

    s["obs_history_buf"] = 0.5 * _randn(gen, N, C.HISTORY_LEN, C.NUM_PROP)
    s["obs_disc_buf"] = 0.5 * _randn(gen, N, C.NUM_OBS_DISC)
    s["episode_sums"] = 0.1 * _randn(gen, C.NUM_REWARDS, N)
    s["feet_air_time"] = _rand(gen, N, 4)
    return s


def make_rng_draws(cfg: BbcEnvConfig, seed: int = 1234, step: int = 0,
                   num_clips_per_mode: Optional[list] = None) -> Dict[str, torch.Tensor]:
    """Dense per-env pre-drawn randoms for *parity mode* (SURVEY.md section 7 "RNG parity").

    The reference draws on variable-length index sets from three RNG sources (torch, numpy,
    multinomial; legged_robot.py:504-540, motion_loader.py:311-341).  In parity mode every env
    gets its own pre-drawn value; the kernels and the oracle consume only the entries of envs
    that actually resample / reset."""
    gen = torch.Generator().manual_seed(seed * 2000003 + step * 211 + 5)
    N = cfg.num_envs
    d = {}
    d["noise_u"] = _rand(gen, N, C.OBS_WIDTH)
    for tag in ("rs", "rt"):                       # rs: periodic resample site, rt: reset site
        d[f"{tag}_eps_u"] = torch.rand(N, generator=gen, dtype=torch.float64)
        d[f"{tag}_c_idx"] = torch.randint(0, C.DIM_C, (N,), generator=gen, dtype=torch.int32)
        d[f"{tag}_cmd_u"] = _rand(gen, N, 5)
    d["push_u"] = _rand(gen, N, 2)
    d["mocap_clip_u"] = torch.rand(N, generator=gen, dtype=torch.float64)   # -> clip within the env's mode
    d["mocap_time_u"] = torch.rand(N, generator=gen, dtype=torch.float64)   # np.random.uniform (float64)
    return d


# ------------------------------------------------------------------------------------------
# Seeded network weights with the reference's state_dict names / shapes (tsc/weights/bbc/model.pt layout).
# The GPU box has neither the reference nor its checkpoint; tests regenerate these from the seed.
# ------------------------------------------------------------------------------------------
AC_SHAPES = [("std", (12,)), ("priv_encoder.0.weight", (64, 29)), ("priv_encoder.0.bias", (64,)),
             ("priv_encoder.2.weight", (29, 64)), ("priv_encoder.2.bias", (29,)),
             ("history_encoder.encoder.0.weight", (30, 57)), ("history_encoder.encoder.0.bias", (30,)),
             ("history_encoder.conv_layers.0.weight", (20, 30, 4)), ("history_encoder.conv_layers.0.bias", (20,)),
             ("history_encoder.conv_layers.2.weight", (10, 20, 2)), ("history_encoder.conv_layers.2.bias", (10,)),
             ("history_encoder.linear_output.0.weight", (29, 30)), ("history_encoder.linear_output.0.bias", (29,)),
             ("actor_trunk.0.weight", (512, 101)), ("actor_trunk.0.bias", (512,)),
             ("actor_trunk.2.weight", (256, 512)), ("actor_trunk.2.bias", (256,)),
             ("actor_trunk.4.weight", (128, 256)), ("actor_trunk.4.bias", (128,)),
             ("actor_head.weight", (12, 128)), ("actor_head.bias", (12,)),
             ("critic_trunk.0.weight", (512, 671)), ("critic_trunk.0.bias", (512,)),
             ("critic_trunk.2.weight", (256, 512)), ("critic_trunk.2.bias", (256,)),
             ("critic_trunk.4.weight", (128, 256)), ("critic_trunk.4.bias", (128,)),
             ("critic_head.weight", (1, 128)), ("critic_head.bias", (1,))]
EST_SHAPES = [("estimator.0.weight", (128, 57)), ("estimator.0.bias", (128,)), ("estimator.2.weight", (64, 128)),
              ("estimator.2.bias", (64,)), ("estimator.4.weight", (4, 64)), ("estimator.4.bias", (4,))]
DISC_SHAPES = [("trunk.0.weight", (512, 98)), ("trunk.0.bias", (512,)), ("trunk.2.weight", (256, 512)),
               ("trunk.2.bias", (256,)), ("linear.weight", (1, 256)), ("linear.bias", (1,)),
               ("classifier.weight", (5, 256)), ("classifier.bias", (5,)), ("encoder_eps.weight", (1, 256)),
               ("encoder_eps.bias", (1,))]


def _make_sd(shapes, gen):
    sd = {}
    for name, shape in shapes:
        if name == "std":
            sd[name] = 0.6 + 0.4 * torch.rand(shape, generator=gen)
        elif name.endswith("bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=gen)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[name] = (2 * torch.rand(shape, generator=gen) - 1) * (1.7 / math.sqrt(fan_in))
    return sd


# TSC teacher (tsc/rsl_rl/modules/actor_critic.py:60-225 ActorCriticTSC with the go2 agility config): obs 800 =
# [prop 65 | scan 132 | priv explicit 4 | priv latent 29 | history 570]; actions = 1 mode index + 3 x 6 continuous
TSC_AC_SHAPES = [("std", (18,)),
                 ("actor.priv_encoder.0.weight", (64, 29)), ("actor.priv_encoder.0.bias", (64,)),
                 ("actor.priv_encoder.2.weight", (29, 64)), ("actor.priv_encoder.2.bias", (29,)),
                 ("actor.history_encoder.encoder.0.weight", (30, 57)), ("actor.history_encoder.encoder.0.bias", (30,)),
                 ("actor.history_encoder.conv_layers.0.weight", (20, 30, 4)), ("actor.history_encoder.conv_layers.0.bias", (20,)),
                 ("actor.history_encoder.conv_layers.2.weight", (10, 20, 2)), ("actor.history_encoder.conv_layers.2.bias", (10,)),
                 ("actor.history_encoder.linear_output.0.weight", (29, 30)), ("actor.history_encoder.linear_output.0.bias", (29,)),
                 ("actor.scan_encoder.0.weight", (128, 132)), ("actor.scan_encoder.0.bias", (128,)),
                 ("actor.scan_encoder.2.weight", (64, 128)), ("actor.scan_encoder.2.bias", (64,)),
                 ("actor.scan_encoder.4.weight", (32, 64)), ("actor.scan_encoder.4.bias", (32,)),
                 ("actor.actor_trunk.0.weight", (512, 130)), ("actor.actor_trunk.0.bias", (512,)),
                 ("actor.actor_trunk.2.weight", (256, 512)), ("actor.actor_trunk.2.bias", (256,)),
                 ("actor.actor_trunk.4.weight", (128, 256)), ("actor.actor_trunk.4.bias", (128,)),
                 ("actor.actor_d.weight", (3, 128)), ("actor.actor_d.bias", (3,)),
                 ("actor.actor_c.weight", (18, 128)), ("actor.actor_c.bias", (18,)),
                 ("critic.0.weight", (512, 800)), ("critic.0.bias", (512,)),
                 ("critic.2.weight", (256, 512)), ("critic.2.bias", (256,)),
                 ("critic.4.weight", (128, 256)), ("critic.4.bias", (128,)),
                 ("critic.6.weight", (1, 128)), ("critic.6.bias", (1,))]


def make_tsc_weights(seed: int = 0):
    """Returns dict(ac=ActorCriticTSC state_dict, est=Estimator(57 -> [128, 64] -> 4) state_dict)."""
    gen = torch.Generator().manual_seed(seed * 7919 + 5)
    return dict(ac=_make_sd(TSC_AC_SHAPES, gen), est=_make_sd(EST_SHAPES, gen))


def make_weights(seed: int = 0):
    """Returns dict(ac=..., est=..., disc=..., norm_mean (98,) f64, norm_var (98,) f64)."""
    gen = torch.Generator().manual_seed(seed * 31337 + 11)
    return dict(ac=_make_sd(AC_SHAPES, gen), est=_make_sd(EST_SHAPES, gen), disc=_make_sd(DISC_SHAPES, gen),
                norm_mean=0.2 * torch.randn(98, generator=gen, dtype=torch.float64),
                norm_var=0.2 + torch.rand(98, generator=gen, dtype=torch.float64))


# ------------------------------------------------------------------------------------------
# TSC (agility teacher) synthetic state: obstacle course goals, obstacle types, 132-point height scan, edge mask.
# Same role as make_static / make_snapshot above for the BBC env; distributions chosen so that every branch of
# tsc/legged_gym/envs/base/legged_robot.py:204-515 is exercised at N >= 64 (goal reached / left, all six termination
# causes, both tracking_goal_vel targets, history fill and shift).
# ------------------------------------------------------------------------------------------
TSC_NUM_BODIES = 17
TSC_FEET = [4, 8, 12, 16]
TSC_PENALISED = [0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 13, 14, 15]
TSC_TERMINATION = [0, 1, 2, 5, 6, 9, 10, 13, 14]
TSC_NUM_GOALS_TOTAL = 6 * 4 + 2


def make_tsc_static(num_envs: int, seed: int = 0, rows: int = 640, cols: int = 960):
    g = torch.Generator().manual_seed(seed * 104729 + 3)
    N = num_envs
    st = {}
    st["height_samples"] = torch.randint(-20, 120, (rows, cols), generator=g, dtype=torch.int16)
    st["x_edge_mask"] = torch.rand(rows, cols, generator=g) < 0.25
    x = torch.tensor([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1])
    y = torch.tensor([-0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5])
    gx, gy = torch.meshgrid(x, y, indexing="ij")
    pts = torch.zeros(N, gx.numel(), 3)
    pts[:, :, 0], pts[:, :, 1] = gx.flatten(), gy.flatten()
    st["height_points"] = pts
    # goals: a course of 24 way points inside the 22 m x 38 m terrain (border 5 m), + the last goal repeated twice
    start = torch.stack([2.0 + 16.0 * torch.rand(N, generator=g), 2.0 + 30.0 * torch.rand(N, generator=g)], dim=1)
    steps = torch.cat([0.9 * torch.rand(N, 24, 1, generator=g) - 0.2, 0.8 * torch.rand(N, 24, 1, generator=g) - 0.4], dim=2)
    xy = start[:, None, :] + torch.cumsum(steps, dim=1)
    goals = torch.cat([xy, 0.1 * torch.rand(N, 24, 1, generator=g)], dim=2)
    reps = []
    for k in range(2):
        last = goals[:, -1:, :].clone()
        last[:, :, 1] += 0.1 * (k + 1)
        reps.append(last)
    st["env_goals"] = torch.cat([goals] + reps, dim=1)
    st["obstacle_types"] = torch.stack([torch.randperm(6, generator=g) for _ in range(N)])
    flat = st["obstacle_types"].flatten()
    has_dof = (flat == 0) | (flat == 3) | (flat == 4)
    st["seesaw_dof_index"] = (flat[has_dof] == 3).nonzero(as_tuple=False).flatten()
    st["feet_indices"] = torch.tensor(TSC_FEET)
    st["key_body_ids"] = torch.tensor(TSC_FEET)
    st["penalised_contact_indices"] = torch.tensor(TSC_PENALISED)
    st["termination_contact_indices"] = torch.tensor(TSC_TERMINATION)
    st["default_dof_pos"] = torch.tensor([[0.0, 0.9, -1.8] * 4])
    st["p_gains"], st["d_gains"] = torch.full((12,), 40.0), torch.full((12,), 1.0)
    st["torque_limits"] = torch.tensor([23.7, 23.7, 45.43] * 4)
    st["gravity_vec"] = torch.tensor([[0.0, 0.0, -1.0]]).repeat(N, 1)
    st["mass_params"] = torch.cat([1.5 * torch.rand(N, 1, generator=g), 0.2 * torch.rand(N, 3, generator=g) - 0.1], dim=1)
    st["friction_coeffs"] = 0.6 + 1.4 * torch.rand(N, 1, generator=g)
    st["motor_strength"] = 0.8 + 0.4 * torch.rand(2, N, 12, generator=g)
    return st


def make_tsc_snapshot(num_envs: int, static, seed: int = 0, step: int = 0):
    g = torch.Generator().manual_seed(seed * 15485863 + step * 7 + 1)
    N, B = num_envs, TSC_NUM_BODIES
    r = lambda *s: torch.rand(*s, generator=g)                               # noqa: E731
    n = lambda *s: torch.randn(*s, generator=g)                              # noqa: E731
    s = {}
    cur = torch.randint(0, 24, (N,), generator=g)
    cur[r(N) < 0.04] = 24                                                   # reach_goal_cutoff (>= G - last_goal_repeat)
    plant = N >= 16                                                          # one env per termination cause, always
    if plant:
        cur[5] = 24
    s["cur_goal_idx"] = cur
    goals = static["env_goals"]
    gi = cur[:, None, None]
    s["cur_goals"] = goals.gather(1, gi.expand(-1, -1, 3)).squeeze(1)
    s["next_goals"] = goals.gather(1, (gi + 1).expand(-1, -1, 3)).squeeze(1)
    # root: mostly 0.2 .. 2 m from the current goal, 20 % inside the 0.4 m goal radius, 2 % further than 4 m
    dist = 0.2 + 1.8 * r(N)
    u = r(N)
    dist = torch.where(u < 0.2, 0.35 * r(N), dist)
    dist = torch.where(u > 0.98, 4.2 + r(N), dist)
    if plant:
        dist[1] = 4.5
    ang = 2 * math.pi * r(N)
    root = torch.zeros(N, 13)
    root[:, 0] = s["cur_goals"][:, 0] - dist * torch.cos(ang)
    root[:, 1] = s["cur_goals"][:, 1] - dist * torch.sin(ang)
    root[:, 2] = 0.22 + 0.3 * r(N)
    root[r(N) < 0.02, 2] = -0.3
    if plant:
        root[2, 2] = -0.3
    yaw = ang + 0.5 * n(N)
    roll, pitch = 0.15 * n(N), 0.15 * n(N)
    roll[r(N) < 0.015] = 1.6
    pitch[r(N) < 0.015] = -1.55
    if plant:
        roll[3], pitch[4] = 1.6, -1.55
    cy, sy, cr, sr, cp, sp = (torch.cos(yaw / 2), torch.sin(yaw / 2), torch.cos(roll / 2), torch.sin(roll / 2),
                              torch.cos(pitch / 2), torch.sin(pitch / 2))
    q = torch.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp,
                     cy * cr * cp + sy * sr * sp], dim=1)
    root[:, 3:7] = q / q.norm(dim=1, keepdim=True)
    root[:, 7:10] = 0.8 * n(N, 3)
    root[:, 7] += 1.0 * torch.cos(ang)
    root[:, 8] += 1.0 * torch.sin(ang)
    root[:, 10:13] = 0.8 * n(N, 3)
    s["root_states"] = root
    dof = torch.zeros(N, 12, 2)
    dof[:, :, 0] = static["default_dof_pos"] + 0.25 * n(N, 12)
    dof[:, :, 1] = 3.0 * n(N, 12)
    s["dof_state"] = dof.reshape(N * 12, 2)
    rb = torch.zeros(N, B, 13)
    rb[:, :, 0:3] = root[:, None, 0:3] + 0.2 * n(N, B, 3)
    rb[:, :, 2] = rb[:, :, 2].clamp(min=0.0)
    s["rigid_body_state"] = rb
    post = rb.clone()
    post[:, :, 0:3] += 0.05 * n(N, B, 3)                                    # what the physics step inside reset_idx leaves
    s["rigid_body_state_post"] = post
    cf = torch.zeros(N, B, 3)
    feet = static["feet_indices"]
    fz = torch.relu(40 + 40 * n(N, 4)) * (r(N, 4) < 0.5)
    cf[:, feet, 2] = fz
    cf[:, feet, 0:2] = 5 * n(N, 4, 2) * (fz > 0).unsqueeze(-1)
    hit = r(N, B) < 0.01
    hit[:, feet] = False
    mag = 1 + 49 * r(N, B)
    d = n(N, B, 3)
    cf = torch.where(hit.unsqueeze(-1), d / d.norm(dim=-1, keepdim=True) * mag.unsqueeze(-1), cf)
    s["contact_forces"] = cf
    ep = torch.randint(0, 2000, (N,), generator=g)
    ep[r(N) < 0.01] = 2000
    ep[r(N) < 0.03] = 0
    if plant:
        ep[6], ep[7] = 2000, 0
    s["episode_length_buf"] = ep
    s["common_step_counter"], s["global_counter"] = 100 + step, 100 + step
    s["last_root_vel"] = root[:, 7:13] + 0.1 * n(N, 6)
    s["last_contacts"] = r(N, 4) < 0.4
    s["reach_goal_timer"] = torch.randint(0, 3, (N,), generator=g).float()
    s["actions"], s["last_actions"] = n(N, 12), n(N, 12)
    s["torques_org"], s["last_torques_org"] = 8 * n(N, 12), 8 * n(N, 12)
    s["last_dof_vel"] = dof[:, :, 1] + 0.5 * n(N, 12)
    s["measured_heights"] = 0.3 * r(N, 132)
    s["delta_yaw"], s["delta_next_yaw"] = n(N), n(N)
    c_idx = torch.randint(0, 5, (N,), generator=g)
    s["latent_c"] = torch.nn.functional.one_hot(c_idx, 5).float()
    s["latent_eps"] = 2 * r(N, 1) - 1
    s["commands"] = torch.cat([2 * r(N, 1), 0.6 * r(N, 1) - 0.3, 2 * r(N, 1) - 1, 0.5 * r(N, 2)], dim=1)
    s["obs_history_buf"] = 0.5 * n(N, 10, 57)
    s["contact_buf"] = (r(N, 100, 4) < 0.5).float()
    s["action_history_buf"] = n(N, 8, 12)
    s["action_hl_history_buf"] = torch.cat([torch.randint(0, 3, (N, 8, 1), generator=g).float(), n(N, 8, 18)], dim=2)
    s["episode_sums"] = 0.3 * n(N, 8)
    s["feet_air_time"] = r(N, 4)
    s["obs_disc_buf"] = n(N, 49)
    s["obst_dof_state"] = 0.2 * n(3 * N, 2)
    return s


def make_tsc_draws(num_envs: int, seed: int = 0, step: int = 0):
    g = torch.Generator().manual_seed(seed * 32452843 + step * 11 + 2)
    return {k: torch.rand(num_envs, generator=g) for k in ("yaw_u", "x_u", "y_u")}


# ---- TSC depth student: formula-generated weights / inputs shared by the reference pin and the tests -----------
def load_student_weights(module, seed: int) -> None:
    """Deterministic weights for a student module, drawn key by key in `state_dict()` order (the 62 400 x 128 linear layer of
    the depth backbone makes a stored fixture impractical).  Tensors that appear under several keys are written several
    times; the last write wins, identically for any implementation with the same keys."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, v in module.state_dict().items():
            if not v.is_floating_point():
                continue
            if k.endswith("running_var"):
                v.copy_(1.0 + 0.1 * torch.rand(v.shape, generator=g))
            elif v.dim() >= 2:
                fan_in = v[0].numel()
                v.copy_(torch.randn(v.shape, generator=g) / fan_in ** 0.5)
            else:
                v.copy_(0.01 * torch.randn(v.shape, generator=g))


def make_student_inputs(num_envs: int, steps: int, seed: int):
    """obs (T,N,800) with plausible auxiliary lanes [57:59] = delta yaw, [59:65] = obstacle one-hot; depth images (T,N,58,87) in
    the normalised range of `process_depth_image`; teacher actions (T*N,19) = [mode index | 18 continuous]; yaw-ok masks."""
    g = torch.Generator().manual_seed(seed)
    T, N = steps, num_envs
    obs = 0.5 * torch.randn(T, N, 800, generator=g)
    obs[:, :, 57:59] = 0.3 * torch.randn(T, N, 2, generator=g)
    kind = torch.randint(0, 6, (T, N), generator=g)
    obs[:, :, 59:65] = torch.nn.functional.one_hot(kind, 6).float()
    depth = torch.rand(T, N, 58, 87, generator=g) - 0.5
    teacher = 0.5 * torch.randn(T * N, 19, generator=g)
    teacher[:, 0] = torch.randint(0, 3, (T * N,), generator=g).float()
    ok = torch.rand(T, N, generator=g) < 0.8
    return dict(obs=obs, depth=depth, actions_teacher=teacher, delta_yaw_ok=ok)

"""Expert demonstration sets for the discriminator update -- what `MotionLoader.pre_load_data` builds at start-up
(bbc/rsl_rl/datasets/motion_loader.py:190-249) and `feed_forward_generator_lb / _ulb` (:513-526) sample from:

  preloaded_s_lb   (n, disc_obs_len * 49)   discriminator-observation histories of the labelled clips
  preloaded_label  (n,)                     behaviour-mode label of each row
  preloaded_s_ulb  (n, disc_obs_len * 49)   the same for the unlabelled clips (concatenated into one trajectory, :177-183)

Start-up code, not the per-step hot path: the frame blending reuses the table the reset kernel reads (`MocapTable`) and
runs as a handful of batched torch ops on whatever device the table lives on (the reference loops over clips on the host
and takes minutes for 2 x 200 000 transitions).  Draws (clip index, time within the clip) are injectable for parity.
"""
import json
import os
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .mocap import FRAME_W, MocapTable, _normalise_quats, _reorder_pybullet_to_isaac

_EPS = float(np.finfo(float).eps * 4.0)


def _slerp(q0, q1, fraction):
    """rsl_rl/utils/utils.py:126-159 (spin 0, shortest path), incl. its 1/angle scaling (:154)."""
    out = torch.zeros_like(q0)
    zero_mask = torch.isclose(fraction, torch.zeros_like(fraction)).squeeze(-1)
    ones_mask = torch.isclose(fraction, torch.ones_like(fraction)).squeeze(-1)
    out[zero_mask] = q0[zero_mask]
    out[ones_mask] = q1[ones_mask]
    d = torch.sum(q0 * q1, dim=-1, keepdim=True)
    dist_mask = (torch.abs(torch.abs(d) - 1.0) < _EPS).squeeze(-1)
    out[dist_mask] = q0[dist_mask]
    neg = d < 0
    d = torch.where(neg, -d, d)
    q1 = torch.where(neg, -q1, q1)
    d = torch.clip(d, -1, 1)
    angle = torch.acos(d)
    angle_mask = (torch.abs(angle) < _EPS).squeeze(-1)
    out[angle_mask] = q0[angle_mask]
    final = ~(zero_mask | ones_mask | dist_mask | angle_mask)
    isin = 1.0 / angle
    mix = q0 * (torch.sin((1.0 - fraction) * angle) * isin) + q1 * (torch.sin(fraction * angle) * isin)
    out[final] = mix[final]
    return out


def frames_at_time(frames, start, lens, nframes, times):
    """MotionLoader.get_full_frame_at_time_batch (:410-447): `frames (F,49)`, per-row clip `start (n,) int64`,
    `lens / nframes / times (n,) float64` -> blended frames (n,49).  Index math in float64 like numpy."""
    p = times / lens
    pn = p * nframes
    lo, hi = torch.floor(pn).long(), torch.ceil(pn).long()
    f0, f1 = frames[start + lo], frames[start + hi]
    blend = (pn - lo.double()).to(torch.float32).unsqueeze(-1)
    pos = (1.0 - blend) * f0[:, 0:3] + blend * f1[:, 0:3]
    rot = _slerp(f0[:, 3:7], f1[:, 3:7], blend)
    traj = (1.0 - blend) * f0[:, 7:FRAME_W] + blend * f1[:, 7:FRAME_W]
    return torch.cat([pos, rot, traj], dim=-1)


def _rotate_inverse(q, v):
    qw, qv = q[:, 3:4], q[:, :3]
    a = v * (2.0 * qw ** 2 - 1.0)
    b = torch.cross(qv, v, dim=-1) * qw * 2.0
    c = qv * torch.sum(qv * v, dim=-1, keepdim=True) * 2.0
    return a - b + c


def _rotate(q, v):
    qw, qv = q[:, 3:4], q[:, :3]
    a = v * (2.0 * qw ** 2 - 1.0)
    b = torch.cross(qv, v, dim=-1) * qw * 2.0
    c = qv * torch.sum(qv * v, dim=-1, keepdim=True) * 2.0
    return a + b + c


def disc_obs_from_frames(fr, default_dof_pos, s: Dict[str, float]):
    """The 49-lane discriminator observation of a mocap frame (:194-216), same lane order as
    LeggedRobot.compute_observations' obs_disc_buf (legged_robot.py:268-275)."""
    q = fr[:, 3:7]
    lin, ang = _rotate_inverse(q, fr[:, 31:34]), _rotate_inverse(q, fr[:, 34:37])
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = torch.atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    pitch = torch.asin(torch.clip(2.0 * (w * y - z * x), -1, 1))
    key = fr[:, 19:31].reshape(-1, 4, 3)
    # compute_flat_key_pos (legged_robot.py:1377-1396): key-body positions in the heading frame
    ref_dir = torch.zeros_like(fr[:, 0:3])
    ref_dir[:, 0] = 1
    rd = _rotate(q, ref_dir)
    half = -torch.atan2(rd[:, 1], rd[:, 0]) / 2
    hq = torch.stack([torch.zeros_like(half), torch.zeros_like(half), torch.sin(half), torch.cos(half)], dim=-1)
    hq = hq / hq.norm(p=2, dim=-1).clamp(min=1e-9).unsqueeze(-1)
    local = (key - fr[:, None, 0:3]).reshape(-1, 3)
    flat_key = _rotate(hq.unsqueeze(1).repeat(1, 4, 1).reshape(-1, 4), local).reshape(-1, 12)
    contact = (key[:, :, -1] < 0.025).to(torch.float32)
    return torch.cat([torch.stack((roll, pitch), dim=1), fr[:, 2:3], lin * s["lin_vel_dist"], ang * s["ang_vel_dist"],
                      (fr[:, 7:19] - default_dof_pos) * s["dof_pos"], fr[:, 37:49] * s["dof_vel"],
                      flat_key * s["key_pos"], contact * s["foot_contact"]], dim=-1)


def load_unlabelled_clips(files: Sequence[str], frame_duration_scale: float = 1.0):
    """The unlabelled clips as ONE trajectory (:152-183): frames concatenated, length = sum of the clip lengths, frame
    count = sum, frame duration = the first clip's.  Returns (frames (F,49) f32, len_s, nframes, frame_dur)."""
    fl, lens, ns, durs = [], [], [], []
    for path in files:
        with open(path, "r") as fh:
            js = json.load(fh)
        f = _reorder_pybullet_to_isaac(np.array(js["Frames"]))
        _normalise_quats(f)
        fl.append(f[:, :FRAME_W])
        d = float(js["FrameDuration"]) * frame_duration_scale
        durs.append(d)
        lens.append((f.shape[0] - 1) * d)
        ns.append(float(f.shape[0]))
    return torch.tensor(np.concatenate(fl, axis=0), dtype=torch.float32), float(np.sum(lens)), float(np.sum(ns)), durs[0]


@dataclass
class ExpertData:
    preloaded_s_lb: torch.Tensor
    preloaded_label: torch.Tensor
    preloaded_s_ulb: torch.Tensor

    @staticmethod
    def build(table: MocapTable, ulb, num_preload: int, time_between_frames: float, default_dof_pos, obs_scales,
              disc_obs_len: int = 2, device="cuda", draws: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0):
        """`table`: the labelled clips; `ulb`: (frames, len_s, nframes, frame_dur) from `load_unlabelled_clips`.
        `draws` (parity): clip_idx_lb (n,) int, time_u_lb (n,) f64, time_u_ulb (n,) f64."""
        dev = torch.device(device)
        n = num_preload
        if draws is None:
            g = torch.Generator().manual_seed(seed)
            cdf = torch.cumsum(table.clip_weight.cpu(), 0)
            clip = torch.searchsorted(cdf, torch.rand(n, generator=g, dtype=torch.float64)).clamp(max=table.num_clips - 1)
            draws = dict(clip_idx_lb=clip, time_u_lb=torch.rand(n, generator=g, dtype=torch.float64),
                         time_u_ulb=torch.rand(n, generator=g, dtype=torch.float64))
        s = {k: float(obs_scales[k] if isinstance(obs_scales, dict) else getattr(obs_scales, k))
             for k in ("lin_vel_dist", "ang_vel_dist", "dof_pos", "dof_vel", "key_pos", "foot_contact")}
        dd = torch.as_tensor(default_dof_pos, dtype=torch.float32).reshape(1, 12).to(dev)
        ci = draws["clip_idx_lb"].long().to(dev)
        t = table.to(dev)

        def rollout(frames, start, lens, nfr, dur, u):
            subst = time_between_frames * disc_obs_len + dur                     # traj_time_sample_batch :333-341
            times = torch.maximum(torch.zeros_like(u) + 1e-7, (lens - subst) * u)
            out = []
            for _ in range(disc_obs_len):
                out.append(disc_obs_from_frames(frames_at_time(frames, start, lens, nfr, times), dd, s))
                times = times + time_between_frames
            return torch.cat(out, dim=-1)

        s_lb = rollout(t.frames, t.clip_start[ci].long(), t.clip_len_s[ci], t.clip_nframes[ci], t.clip_frame_dur[ci],
                       draws["time_u_lb"].to(dev))
        fu, len_u, n_u, dur_u = ulb
        one = torch.ones(n, dtype=torch.float64, device=dev)
        s_ulb = rollout(fu.to(dev), torch.zeros(n, dtype=torch.int64, device=dev), one * len_u, one * n_u, one * dur_u,
                        draws["time_u_ulb"].to(dev))
        return ExpertData(s_lb, t.clip_label[ci].long(), s_ulb)

    # reference-shaped generators (:513-526); `SSInfoGAIL.update_disc` draws its indices on the device instead
    def feed_forward_generator_lb(self, num_mini_batch, mini_batch_size):
        for _ in range(num_mini_batch):
            idx = torch.randint(self.preloaded_s_lb.shape[0], (mini_batch_size,), device=self.preloaded_s_lb.device)
            yield self.preloaded_s_lb[idx], self.preloaded_label[idx]

    def feed_forward_generator_ulb(self, num_mini_batch, mini_batch_size):
        for _ in range(num_mini_batch):
            idx = torch.randint(self.preloaded_s_ulb.shape[0], (mini_batch_size,), device=self.preloaded_s_ulb.device)
            yield self.preloaded_s_ulb[idx]

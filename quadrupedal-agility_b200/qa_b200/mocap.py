"""Labelled mocap clip table used for reference-state initialisation at reset.

Host-side, load-time counterpart of the reference's `MotionLoader(mocap_state_init=True)`
(bbc/rsl_rl/datasets/motion_loader.py:52-147): parse the clip JSON (61 floats/frame),
convert PyBullet leg order [FR,FL,RR,RL] to Isaac order [FL,FR,RL,RR] (`reorder`,
:251-302), normalise + standardise (w >= 0) the root quaternion (:121-128), keep the first
49 floats of every frame (`mocap_trajectory_full_lb`, :133-135) and remember per clip its
label, weight, frame duration, length in seconds and frame count (:136-143).

The result is ONE flat fp32 table `(total_frames, 49)` that lives in HBM (234 KB for the 17
shipped clips, L2-resident) plus small per-clip metadata; the fused reset path of the
post-physics kernel blends two rows of it per resetting env.
"""
import json
import os
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from . import config as C

FRAME_W = C.MOCAP_FRAME_WIDTH   # 49 = pos3 quat4 joint12 toe12 linvel3 angvel3 jointvel12


def _reorder_pybullet_to_isaac(frames: np.ndarray) -> np.ndarray:
    """motion_loader.py:251-302.  Input (F,61) float64 in PyBullet order, output (F,61)."""
    f = np.array(frames, dtype=np.float64, copy=True)
    pos, rot = f[:, 0:3], f[:, 3:7]
    leg = lambda a: [a[:, 3 * i:3 * i + 3].copy() for i in range(4)]   # noqa: E731
    fr, fl, rr, rl = leg(f[:, 7:19])
    for j in (fr, fl, rr, rl):
        j[:, 0] = -j[:, 0]                       # hip abduction sign flip
    joint_pos = np.hstack([fl, fr, rl, rr])
    tfr, tfl, trr, trl = leg(f[:, 19:31])
    mins = [np.min(t[:, -1]) for t in (tfl, tfr, trl, trr)]
    pos = pos.copy()
    pos[:, -1] -= np.mean(mins)                 # put the feet on the ground
    for t, m in zip((tfl, tfr, trl, trr), mins):
        t[:, -1] -= m
    toe_pos = np.hstack([tfl, tfr, trl, trr])
    lin, ang = f[:, 31:34], f[:, 34:37]
    vfr, vfl, vrr, vrl = leg(f[:, 37:49])
    for j in (vfr, vfl, vrr, vrl):
        j[:, 0] = -j[:, 0]
    joint_vel = np.hstack([vfl, vfr, vrl, vrr])
    wfr, wfl, wrr, wrl = leg(f[:, 49:61])
    toe_vel = np.hstack([wfl, wfr, wrl, wrr])
    return np.hstack([pos, rot, joint_pos, toe_pos, lin, ang, joint_vel, toe_vel])


def _normalise_quats(f: np.ndarray) -> None:
    """pose3d.QuaternionNormalize + motion_util.standardize_quaternion, motion_loader.py:121-128."""
    q = f[:, 3:7]
    n = np.linalg.norm(q, axis=1, keepdims=True)
    if np.any(np.isclose(n, 0.0)):
        raise ValueError("Quaternion may not be zero in a mocap frame")
    q = q / n
    q = np.where(q[:, 3:4] < 0, -q, q)
    f[:, 3:7] = q


@dataclass
class MocapTable:
    frames: torch.Tensor          # (F,49) f32
    clip_start: torch.Tensor      # (K,) i32 first row of each clip
    clip_nframes: torch.Tensor    # (K,) f64 (the reference keeps float(n), :142)
    clip_len_s: torch.Tensor      # (K,) f64 (n-1)*frame_duration, :140-141
    clip_frame_dur: torch.Tensor  # (K,) f64
    clip_label: torch.Tensor      # (K,) i32 index into mocap_category
    clip_weight: torch.Tensor     # (K,) f64 normalised over all clips, :146
    # per-mode clip lists (CSR) with the within-mode normalised CDF, for the in-kernel Philox path
    mode_offset: torch.Tensor     # (DIM_C+1,) i32
    mode_clips: torch.Tensor      # (K,) i32 clip ids grouped by mode
    mode_cdf: torch.Tensor        # (K,) f64 inclusive CDF within the mode
    names: List[str]

    @property
    def num_clips(self) -> int:
        return int(self.clip_start.numel())

    def to(self, device):
        kw = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return MocapTable(**kw)

    # ---- construction -------------------------------------------------------------------
    @staticmethod
    def _finish(frames_list, labels, weights, durs, names) -> "MocapTable":
        starts, n, acc = [], [], 0
        for fr in frames_list:
            starts.append(acc)
            n.append(fr.shape[0])
            acc += fr.shape[0]
        w = np.asarray(weights, dtype=np.float64)
        w = w / np.sum(w)
        durs = np.asarray(durs, dtype=np.float64)
        nf = np.asarray(n, dtype=np.float64)
        lens = (nf - 1.0) * durs
        labels = np.asarray(labels, dtype=np.int32)
        offs, clips, cdf = [0], [], []
        for m in range(C.DIM_C):
            ids = np.nonzero(labels == m)[0]
            # motion_loader.py:314-320: p = w[mode] / sum(w[mode])
            p = w[ids] / np.sum(w[ids]) if len(ids) else np.zeros(0)
            clips.extend(ids.tolist())
            cdf.extend(np.cumsum(p).tolist())
            offs.append(len(clips))
        table = np.concatenate(frames_list, axis=0)
        return MocapTable(
            frames=torch.tensor(table, dtype=torch.float32),
            clip_start=torch.tensor(starts, dtype=torch.int32),
            clip_nframes=torch.tensor(nf, dtype=torch.float64),
            clip_len_s=torch.tensor(lens, dtype=torch.float64),
            clip_frame_dur=torch.tensor(durs, dtype=torch.float64),
            clip_label=torch.tensor(labels, dtype=torch.int32),
            clip_weight=torch.tensor(w, dtype=torch.float64),
            mode_offset=torch.tensor(offs, dtype=torch.int32),
            mode_clips=torch.tensor(clips, dtype=torch.int32),
            mode_cdf=torch.tensor(cdf, dtype=torch.float64),
            names=list(names),
        )

    @staticmethod
    def from_json_files(files: Sequence[str], mocap_category=C.MOCAP_CATEGORY,
                        frame_duration_scale: float = 1.0) -> "MocapTable":
        frames_list, labels, weights, durs, names = [], [], [], [], []
        for path in files:
            name = os.path.basename(path)
            label = None
            for idx, cate in enumerate(mocap_category):       # last match wins, :113-115
                if cate in name:
                    label = idx
            if label is None:
                raise ValueError("Unsupported mocap category {}.".format(path))
            with open(path, "r") as fh:
                js = json.load(fh)
            f = _reorder_pybullet_to_isaac(np.array(js["Frames"]))
            _normalise_quats(f)
            frames_list.append(f[:, :FRAME_W])
            labels.append(label)
            weights.append(float(js["MotionWeight"]))
            durs.append(float(js["FrameDuration"]) * frame_duration_scale)
            names.append(name)
        return MocapTable._finish(frames_list, labels, weights, durs, names)

    @staticmethod
    def from_npz(path: str) -> "MocapTable":
        z = np.load(path, allow_pickle=False)
        starts = z["clip_start"].astype(np.int64)
        nf = z["clip_nframes"].astype(np.int64)
        frames_list = [z["frames"][s:s + n].astype(np.float64) for s, n in zip(starts, nf)]
        # the fp32 table is kept bit-exact: _finish round-trips f32 -> f64 -> f32
        return MocapTable._finish(frames_list, z["clip_label"], z["clip_weight_raw"], z["clip_frame_dur"],
                                  [str(s) for s in z["names"]])

    def save_npz(self, path: str, raw_weights: Sequence[float]) -> None:
        np.savez_compressed(
            path, frames=self.frames.numpy(), clip_start=self.clip_start.numpy(),
            clip_nframes=self.clip_nframes.numpy(), clip_label=self.clip_label.numpy(),
            clip_weight_raw=np.asarray(raw_weights, dtype=np.float64),
            clip_frame_dur=self.clip_frame_dur.numpy(), names=np.array(self.names))

    @staticmethod
    def synthetic(seed: int = 0, clips_per_mode: int = 3, frames_per_clip: int = 64) -> "MocapTable":
        """Smooth random clips with the shipped table's structure (for boxes without the data)."""
        rng = np.random.default_rng(seed)
        frames_list, labels, weights, durs, names = [], [], [], [], []
        for m in range(C.DIM_C):
            for k in range(clips_per_mode):
                n = frames_per_clip + 7 * k
                t = np.linspace(0, 2 * np.pi, n)[:, None]
                f = 0.3 * np.sin(t * rng.uniform(0.5, 2.0, (1, FRAME_W)) + rng.uniform(0, 6.28, (1, FRAME_W)))
                f[:, 2] = 0.3 + 0.05 * f[:, 2]
                f[:, 3:7] += np.array([[0, 0, 0, 1.0]])
                f61 = np.zeros((n, 61))
                f61[:, :FRAME_W] = f
                _normalise_quats(f61)
                frames_list.append(f61[:, :FRAME_W])
                labels.append(m)
                weights.append(1.0 + 0.5 * k)
                durs.append(1.0 / 30.0)
                names.append(f"{C.MOCAP_CATEGORY[m]}_{k}.json")
        return MocapTable._finish(frames_list, labels, weights, durs, names)

    # ---- host-side sampling helper (parity mode) ------------------------------------------
    def sample_clip(self, mode_idx: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
        """Inverse-CDF pick of a clip id within each env's mode (dense form of
        motion_loader.py:310-323).  mode_idx (N,) int, u (N,) f64 in [0,1)."""
        off = self.mode_offset.cpu().numpy()
        clips = self.mode_clips.cpu().numpy()
        cdf = self.mode_cdf.cpu().numpy()
        m = mode_idx.cpu().numpy().astype(np.int64)
        uu = u.cpu().numpy()
        out = np.zeros(len(m), dtype=np.int32)
        for i in range(len(m)):
            lo, hi = off[m[i]], off[m[i] + 1]
            j = int(np.searchsorted(cdf[lo:hi], uu[i], side="right"))
            out[i] = clips[lo + min(j, hi - lo - 1)]
        return torch.tensor(out, dtype=torch.int32)

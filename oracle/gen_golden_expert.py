"""Pin `qa_b200.expert.ExpertData.build` against the UNMODIFIED reference `MotionLoader.pre_load_data`
(bbc/rsl_rl/datasets/motion_loader.py:190-249) on the shipped labelled clips + the first three unlabelled clips, with the
numpy draws injected, and write tests/golden/expert_preload_n96.npz (incl. the three unlabelled clips' parsed frames, which
do not travel otherwise).  Build container only."""
import glob
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

from ref_harness import import_reference, REFERENCE_ROOT  # noqa: E402
from qa_b200.expert import ExpertData, load_unlabelled_clips  # noqa: E402
from qa_b200.mocap import MocapTable  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class _Rand:
    def __init__(self, queue):
        self.queue = list(queue)

    def choice(self, a, size=None, p=None, replace=True):
        kind, v = self.queue.pop(0)
        assert kind == "choice" and len(v) == size
        return np.asarray(a)[v] if len(a) > 1 else np.zeros(size, dtype=np.int64)

    def uniform(self, low=0.0, high=1.0, size=None):
        kind, v = self.queue.pop(0)
        assert kind == "uniform" and len(v) == size
        return v


def main():
    ref = import_reference("bbc")
    ML = ref.motion_loader
    n = 96
    files_lb = sorted(glob.glob(os.path.join(REFERENCE_ROOT, "bbc", "mocap_data", "mocap_all_lb", "*.json")))
    files_ulb = sorted(glob.glob(os.path.join(REFERENCE_ROOT, "bbc", "mocap_data", "mocap_all_ulb", "*.json")))[:3]
    rcfg = ref.envs.Go2LocomotionCfg()
    scales = rcfg.normalization.obs_scales
    dd = torch.tensor([[0.0, 0.9, -1.8] * 4])
    rng = np.random.default_rng(5)
    clip = rng.integers(0, len(files_lb), n)
    u_lb, u_ulb = rng.random(n), rng.random(n)
    u_lb[0], u_ulb[1] = 0.0, 0.999999
    real = ML.np.random
    ML.np.random = _Rand([("choice", clip), ("uniform", u_lb), ("choice", np.zeros(n, dtype=np.int64)), ("uniform", u_ulb)])
    try:
        loader = ML.MotionLoader("cpu", time_between_frames=0.02, mocap_state_init=False, motion_files_lb=files_lb,
                                 motion_files_ulb=files_ulb, mocap_category=rcfg.env.mocap_category_all,
                                 num_preload_transitions=n, compute_flat_key_pos=ref.legged_robot.compute_flat_key_pos,
                                 default_dof_pos=dd, obs_scales=scales, num_disc_obs=49, disc_obs_len=2,
                                 obs_disc_weight_step=0.0, frame_duration_scale=rcfg.env.frame_duration_scale)
    finally:
        ML.np.random = real
    table = MocapTable.from_json_files(files_lb, mocap_category=rcfg.env.mocap_category_all,
                                       frame_duration_scale=rcfg.env.frame_duration_scale)
    ulb = load_unlabelled_clips(files_ulb, rcfg.env.frame_duration_scale)
    s = {k: getattr(scales, k) for k in ("lin_vel_dist", "ang_vel_dist", "dof_pos", "dof_vel", "key_pos", "foot_contact")}
    draws = dict(clip_idx_lb=torch.from_numpy(clip), time_u_lb=torch.from_numpy(u_lb), time_u_ulb=torch.from_numpy(u_ulb))
    ex = ExpertData.build(table, ulb, n, 0.02, dd, s, disc_obs_len=2, device="cpu", draws=draws)
    for name, got, want in (("s_lb", ex.preloaded_s_lb, loader.preloaded_s_lb), ("s_ulb", ex.preloaded_s_ulb, loader.preloaded_s_ulb)):
        err = float((got - want).abs().max())
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-6), (name, err)
        print(f"  expert preload {name}: max abs err vs reference {err:.2e}")
    assert torch.equal(ex.preloaded_label, loader.preloaded_label.long())
    np.savez_compressed(os.path.join(GOLD, "expert_preload_n96.npz"), clip=clip, u_lb=u_lb, u_ulb=u_ulb,
                        ulb_frames=ulb[0].numpy(), ulb_meta=np.array(ulb[1:], dtype=np.float64),
                        scales=np.array([s[k] for k in ("lin_vel_dist", "ang_vel_dist", "dof_pos", "dof_vel", "key_pos", "foot_contact")]),
                        s_lb=loader.preloaded_s_lb.numpy(), label=loader.preloaded_label.numpy(), s_ulb=loader.preloaded_s_ulb.numpy())
    print("wrote tests/golden/expert_preload_n96.npz")


if __name__ == "__main__":
    main()

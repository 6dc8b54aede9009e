"""Expert preload (bbc/rsl_rl/datasets/motion_loader.py:190-249): `ExpertData.build` against the sets the
UNMODIFIED reference MotionLoader produced for the same injected draws (oracle/gen_golden_expert.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, mocap_table
from qa_b200.expert import ExpertData

KEYS = ("lin_vel_dist", "ang_vel_dist", "dof_pos", "dof_vel", "key_pos", "foot_contact")


def _build(device, z, draws=True):
    ulb = (torch.from_numpy(z["ulb_frames"]), *[float(v) for v in z["ulb_meta"]])
    s = dict(zip(KEYS, z["scales"].tolist()))
    d = dict(clip_idx_lb=torch.from_numpy(z["clip"]), time_u_lb=torch.from_numpy(z["u_lb"]),
             time_u_ulb=torch.from_numpy(z["u_ulb"])) if draws else None
    return ExpertData.build(mocap_table(), ulb, 96, 0.02, [0.0, 0.9, -1.8] * 4, s, device=device, draws=d, seed=3)


def test_preload_matches_reference_golden_cpu():
    z = np.load(os.path.join(GOLD, "expert_preload_n96.npz"))
    ex = _build("cpu", z)
    assert torch.equal(ex.preloaded_s_lb, torch.from_numpy(z["s_lb"]))
    assert torch.equal(ex.preloaded_s_ulb, torch.from_numpy(z["s_ulb"]))
    assert torch.equal(ex.preloaded_label, torch.from_numpy(z["label"]).long())
    assert ex.preloaded_s_lb.shape == (96, 98) and set(ex.preloaded_label.tolist()) <= set(range(5))


def test_seeded_draws_and_generators():
    z = np.load(os.path.join(GOLD, "expert_preload_n96.npz"))
    a, b = _build("cpu", z, draws=False), _build("cpu", z, draws=False)
    assert torch.equal(a.preloaded_s_lb, b.preloaded_s_lb) and torch.isfinite(a.preloaded_s_ulb).all()
    s, lab = next(a.feed_forward_generator_lb(2, 17))
    assert s.shape == (17, 98) and lab.shape == (17,)
    assert next(a.feed_forward_generator_ulb(1, 5)).shape == (5, 98)


@pytest.mark.gpu
def test_preload_on_device_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "expert_preload_n96.npz"))
    ex = _build("cuda:0", z)
    assert ex.preloaded_s_lb.is_cuda
    torch.testing.assert_close(ex.preloaded_s_lb.cpu(), torch.from_numpy(z["s_lb"]), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(ex.preloaded_s_ulb.cpu(), torch.from_numpy(z["s_ulb"]), rtol=1e-5, atol=2e-6)
    assert torch.equal(ex.preloaded_label.cpu(), torch.from_numpy(z["label"]).long())

"""Pin the TSC oracle pieces against the UNMODIFIED reference (`/root/reference/tsc`) and write their golden vectors
(build container only).  Run in its own process: bbc/ and tsc/ fork the same package names.

  python oracle/gen_golden_tsc.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

import tsc_depth as OD  # noqa: E402
from ref_harness import import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class _FakeGym:
    """Only what update_depth_buffer touches (:181-202)."""

    def __init__(self, images):
        self.images = images

    def step_graphics(self, sim): pass
    def render_all_camera_sensors(self, sim): pass
    def start_access_image_tensors(self, sim): pass
    def end_access_image_tensors(self, sim): pass

    def get_camera_image_gpu_tensor(self, sim, env, cam, kind):
        return self.images[env]


def depth_case(ref, N, seed):
    g = torch.Generator().manual_seed(seed)
    images = -(0.1 + 6.0 * torch.rand(N, 60, 106, generator=g))          # camera depth: negative metres, some beyond far
    images[0, :5] = -float("inf")                                        # sky pixels
    ep = torch.tensor([0, 1, 2, 17, 1, 250][:N] + [5] * max(0, N - 6), dtype=torch.int64)
    buf0 = torch.randn(N, 2, 58, 87, generator=g) * 0.2
    env = ref.LeggedRobot.__new__(ref.LeggedRobot)
    depth = types.SimpleNamespace(use_camera=True, update_interval=1, near_clip=0.3, far_clip=4, depth_noise=0.05, buffer_len=2)
    env.cfg = types.SimpleNamespace(depth=depth)
    env.global_counter, env.num_envs, env.device = 5, N, "cpu"
    env.sim, env.envs, env.cam_handles = None, list(range(N)), list(range(N))
    env.gym = _FakeGym(images)
    env.episode_length_buf = ep
    env.depth_buffer = buf0.clone()
    # the reference consumes the default CPU generator per env: rand(1), rand(1), rand_like(58x87) -- replay it
    torch.manual_seed(seed + 1)
    u1, u2, up = [], [], []
    for _ in range(N):
        u1.append(torch.rand(1)[0])
        u2.append(torch.rand(1)[0])
        up.append(torch.rand(58, 87))
    u1, u2, up = torch.stack(u1), torch.stack(u2), torch.stack(up)
    torch.manual_seed(seed + 1)
    env.update_depth_buffer()                                             # the reference's own code
    want = env.depth_buffer
    got = OD.update_depth_buffer(buf0, images, ep, 0.3, 4, 0.05, u1, u2, up)
    assert torch.equal(got, want), f"depth buffer: oracle != reference (max err {(got - want).abs().max()})"
    print(f"  depth: N={N}: oracle == reference update_depth_buffer (bit-exact), init envs={int((ep <= 1).sum())}")
    return dict(images=images, ep=ep, buf0=buf0, u1=u1, u2=u2, up=up, want=want)


def main():
    ref = import_reference("tsc")
    print("reference tsc imported from", ref.root)
    d = depth_case(ref, 6, seed=11)
    np.savez_compressed(os.path.join(GOLD, "tsc_depth_n6.npz"), **{k: v.numpy() for k, v in d.items()})
    print("wrote tests/golden/tsc_depth_n6.npz")


if __name__ == "__main__":
    main()

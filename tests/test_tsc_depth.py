"""TSC student depth path: oracle vs the reference-generated golden fixture (CPU), and the CUDA
kernel K14 vs the oracle / the fixture (GPU), bit-exact -- the arithmetic is clip / scale / add in fp32."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import tsc_depth as OD  # noqa: E402
from helpers import GOLD  # noqa: E402

NEAR, FAR, NOISE = 0.3, 4, 0.05


def _golden():
    z = np.load(os.path.join(GOLD, "tsc_depth_n6.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_oracle_matches_reference_golden():
    g = _golden()
    got = OD.update_depth_buffer(g["buf0"], g["images"], g["ep"], NEAR, FAR, NOISE, g["u1"], g["u2"], g["up"])
    assert torch.equal(got, g["want"])
    assert int((g["ep"] <= 1).sum()) == 3                      # both branches of :195-200 are in the fixture


def _synthetic(N, seed):
    g = torch.Generator().manual_seed(seed)
    images = -(0.05 + 7.0 * torch.rand(N, 60, 106, generator=g))
    images[::7, :3] = -float("inf")
    ep = torch.randint(0, 1001, (N,), generator=g)
    ep[::5] = 1
    ep[1::9] = 0
    buf0 = 0.3 * torch.randn(N, 2, 58, 87, generator=g)
    u1, u2, up = torch.rand(N, generator=g), torch.rand(N, generator=g), torch.rand(N, 58, 87, generator=g)
    return images, ep, buf0, u1, u2, up


@pytest.mark.gpu
def test_kernel_matches_reference_golden_pointer_table():
    from qa_b200.depth import DepthBuffer
    g = _golden()
    dev = "cuda:0"
    db = DepthBuffer(6, device=dev)
    db.depth_buffer.copy_(g["buf0"])
    cams = [g["images"][i].to(dev).contiguous() for i in range(6)]          # N separate allocations, like IsaacGym's
    db.set_camera_tensors(cams)
    db.set_parity_draws({"noise_scale_u": g["u1"].to(dev), "offset_u": g["u2"].to(dev), "pixel_u": g["up"].to(dev)})
    db.update_depth_buffer(g["ep"].to(dev), global_counter=5)
    assert torch.equal(db.depth_buffer.cpu(), g["want"])


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 33, 4096])
def test_kernel_matches_oracle_batched(N):
    from qa_b200.depth import DepthBuffer
    dev = "cuda:0"
    images, ep, buf0, u1, u2, up = _synthetic(N, seed=N)
    want = OD.update_depth_buffer(buf0, images, ep, NEAR, FAR, NOISE, u1, u2, up)
    db = DepthBuffer(N, device=dev)
    db.depth_buffer.copy_(buf0)
    db.set_batched_images(images.to(dev))
    db.set_parity_draws({"noise_scale_u": u1.to(dev), "offset_u": u2.to(dev), "pixel_u": up.to(dev)})
    db.update_depth_buffer(ep.to(dev))
    assert torch.equal(db.depth_buffer.cpu(), want)
    # second frame: the buffer shifts for running envs (:199)
    ep2 = ep + 1
    want2 = OD.update_depth_buffer(want, images, ep2, NEAR, FAR, NOISE, u1, u2, up)
    db.update_depth_buffer(ep2.to(dev))
    assert torch.equal(db.depth_buffer.cpu(), want2)


@pytest.mark.gpu
def test_philox_mode_is_bounded_deterministic_and_step_dependent():
    """Production RNG: values stay within the analytic range, two buffers with the same seed agree, steps differ."""
    from qa_b200.depth import DepthBuffer
    dev = "cuda:0"
    N = 64
    images, ep, _, _, _, _ = _synthetic(N, seed=3)
    outs = []
    for _ in range(2):
        db = DepthBuffer(N, device=dev, seed=77)
        db.set_batched_images(images.to(dev))
        db.update_depth_buffer(torch.ones(N, dtype=torch.int64, device=dev))
        outs.append(db.depth_buffer.clone())
    assert torch.equal(outs[0], outs[1])
    assert float(outs[0].min()) >= -0.5 - 3 * NOISE and float(outs[0].max()) <= 0.5 + 3 * NOISE
    first = outs[0][:, -1].clone()
    db.update_depth_buffer(torch.full((N,), 5, dtype=torch.int64, device=dev))
    assert torch.equal(db.depth_buffer[:, 0], first)            # shifted
    assert not torch.equal(db.depth_buffer[:, 1], first)        # new noise at the next step
    clean = OD.process_depth_images(images, NEAR, FAR, 0.0, torch.zeros(N), torch.zeros(N), torch.zeros(N, 58, 87))
    assert float((db.depth_buffer[:, 1].cpu() - clean).abs().max()) <= 3 * NOISE + 1e-6


def test_abi_rejects_bad_depth_arguments():
    import ctypes
    from qa_b200 import _abi
    lib = _abi.load()
    a = _abi.QaDepthArgs()
    a.num_envs = 4
    assert lib.qa_depth_update(ctypes.byref(a), None) == -1        # no images
    a.images, a.episode_length_buf, a.depth_buffer = 8, 8, 8
    a.in_h, a.in_w, a.out_h, a.out_w, a.buffer_len, a.clip_span = 60, 106, 58, 87, 2, 3.7
    a.crop_top, a.crop_left = 5, 10
    assert lib.qa_depth_update(ctypes.byref(a), None) == -2        # crop window leaves the image

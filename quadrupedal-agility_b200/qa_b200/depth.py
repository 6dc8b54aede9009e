"""TSC student depth path behind the reference's names (tsc/legged_gym/envs/base/legged_robot.py:154-202):
`DepthBuffer.update_depth_buffer()` is what `LeggedRobot.update_depth_buffer` does, for all envs in ONE kernel
launch (K14 `qa_depth_update`) instead of a Python loop over `num_envs` camera tensors with CPU random draws.

The camera tensors stay where IsaacGym allocates them: the kernel reads them through a device array of N pointers
(`set_camera_tensors`), or from one batched `(N,60,106)` tensor.
"""
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from . import ops


@dataclass
class DepthConfig:
    """cfg.depth of tsc/legged_gym/envs/base/legged_robot_config.py:63-84 (fields the hot path reads)."""
    use_camera: bool = True
    update_interval: int = 1
    original: tuple = (106, 60)        # (W, H)
    resized: tuple = (87, 58)          # (W, H)
    buffer_len: int = 2
    near_clip: float = 0.3
    far_clip: float = 4
    depth_noise: float = 0.05
    crop_top: int = 1                  # crop_depth_image: image[1:-1, 10:-9]  (:172-174)
    crop_left: int = 10


class DepthBuffer:
    def __init__(self, num_envs: int, cfg: DepthConfig = None, device="cuda", seed: int = 0):
        self.cfg = cfg or DepthConfig()
        self.num_envs, self.device = num_envs, torch.device(device)
        W, H = self.cfg.resized
        self.depth_buffer = torch.zeros(num_envs, self.cfg.buffer_len, H, W, device=self.device)   # :1098
        self.seed, self.step = seed, 0
        self._ptrs = None
        self._batched = None
        self._draws: Optional[Dict[str, torch.Tensor]] = None

    def set_camera_tensors(self, tensors: Sequence[torch.Tensor]) -> None:
        """The N wrapped `gym.get_camera_image_gpu_tensor(..., IMAGE_DEPTH)` tensors (:184-189); their addresses are
        stable for the life of the simulation, so the pointer table is built once."""
        W, H = self.cfg.original
        for t in tensors:
            if not t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != (H, W) or not t.is_contiguous():
                raise RuntimeError("qa_b200: camera tensors must be contiguous CUDA float32 (H,W) images")
        self._keep = list(tensors)
        self._ptrs = torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=self.device)
        self._batched = None

    def set_batched_images(self, images: torch.Tensor) -> None:
        W, H = self.cfg.original
        if tuple(images.shape) != (self.num_envs, H, W):
            raise RuntimeError("qa_b200: batched depth images must be (num_envs, H, W)")
        self._batched, self._ptrs = images, None

    def set_parity_draws(self, draws: Optional[Dict[str, torch.Tensor]]) -> None:
        """dense uniforms {noise_scale_u (N), offset_u (N), pixel_u (N,H,W)} replacing the in-kernel Philox stream."""
        self._draws = draws

    def update_depth_buffer(self, episode_length_buf: torch.Tensor, global_counter: int = 0) -> None:
        c = self.cfg
        if not c.use_camera or global_counter % c.update_interval != 0:                     # :177-181
            return
        self.step += 1
        d = self._draws or {}
        ops.depth_update(self._ptrs, self._batched, episode_length_buf, self.depth_buffer, c.original[1], c.original[0],
                         c.crop_top, c.crop_left, c.near_clip, c.far_clip, c.depth_noise, d.get("noise_scale_u"),
                         d.get("offset_u"), d.get("pixel_u"), self.seed, self.step)

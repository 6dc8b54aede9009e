"""CPU restatement of the TSC student depth path (TEST INFRASTRUCTURE -- only tests/, smoke() and bench.py's
cpu_baseline leg may import this).

Follows tsc/legged_gym/envs/base/legged_robot.py:154-202 (`normalize_depth_image`, `process_depth_image`,
`crop_depth_image`, `update_depth_buffer`) for ALL envs at once.  The reference draws three random sources per env
(`torch.rand(1)` twice on the CPU generator, `torch.rand_like(image)`); here they are dense per-env inputs
(`noise_scale_u (N)`, `offset_u (N)`, `pixel_u (N,58,87)`), the same interposition as the BBC oracle's.
Pinned bit-exact against the reference by oracle/gen_golden_tsc.py (fixture tests/golden/tsc_depth_n6.npz).
"""
import torch


def process_depth_images(images, near_clip, far_clip, depth_noise, noise_scale_u, offset_u, pixel_u):
    """images (N,60,106) camera depth (negative metres) -> (N,58,87) normalised noisy depth.  :161-174"""
    x = images[:, 1:-1, 10:-9]                                               # crop_depth_image :174
    x = torch.clip(x, -far_clip, -near_clip)                                 # :164
    x = x * -1                                                               # normalize_depth_image :155-158
    x = (x - near_clip) / (far_clip - near_clip) - 0.5
    n1 = (depth_noise * noise_scale_u).view(-1, 1, 1)                        # depth_noise = cfg * rand(1)[0]  :168
    x = x + (depth_noise * 2 * (offset_u - 0.5)).view(-1, 1, 1)              # :169
    x = x + n1 * 2 * (pixel_u - 0.5)                                         # :170
    return x


def update_depth_buffer(depth_buffer, images, episode_length_buf, near_clip, far_clip, depth_noise, noise_scale_u,
                        offset_u, pixel_u):
    """depth_buffer (N,L,58,87) -> new buffer: envs with episode_length_buf <= 1 are filled with the new frame,
    the others shift by one and append it.  :176-202"""
    new = process_depth_images(images, near_clip, far_clip, depth_noise, noise_scale_u, offset_u, pixel_u)
    L = depth_buffer.shape[1]
    init = (episode_length_buf <= 1).view(-1, 1, 1, 1)
    filled = new.unsqueeze(1).expand(-1, L, -1, -1)
    shifted = torch.cat([depth_buffer[:, 1:], new.unsqueeze(1)], dim=1)
    return torch.where(init, filled, shifted)

"""Host probe for K3 exactness: a numpy fp32 emulation of k_height_scan (no FMA) against torch-CPU get_heights indices."""
import sys, numpy as np, torch
sys.path[:0] = ["quadrupedal-agility_b200", "oracle", "tests"]
import bbc_env as O
from qa_b200 import synthetic
from qa_b200.config import BbcEnvConfig
torch.set_num_threads(4)
cfg = BbcEnvConfig(num_envs=4096)
st = synthetic.make_static(cfg, seed=5)
snap = synthetic.make_snapshot(cfg, seed=5, step=0)
rs = snap["root_states"]
N, P = rs.shape[0], st["height_points"].shape[0]
hp = st["height_points"].unsqueeze(0).expand(N, P, 3)
q = rs[:, 3:7]
pts = O.quat_apply_yaw(q.repeat(1, P), hp) + rs[:, :3].unsqueeze(1)
pts = pts + cfg.border_size
idx_t = (pts / cfg.horizontal_scale).long()[..., :2].numpy()

f = np.float32
def emulate(fma_cross=False, norm_mode="sqrt"):
    z, w = q[:, 2].numpy().astype(f), q[:, 3].numpy().astype(f)
    if norm_mode == "sqrt":
        n = np.sqrt(z * z + w * w, dtype=f)
    elif norm_mode == "sum4":
        n = np.sqrt((f(0) + f(0)) + (z * z + w * w), dtype=f)
    elif norm_mode == "f64":
        n = np.sqrt(z.astype(np.float64) ** 2 + w.astype(np.float64) ** 2).astype(f)
    n = np.maximum(n, f(1e-9))
    qz, qw = (z / n)[:, None], (w / n)[:, None]
    qx = qy = np.zeros_like(qz)
    b = st["height_points"].numpy().astype(f)[None]
    bx, by, bz = b[..., 0], b[..., 1], b[..., 2]
    def cross(ax, ay, az, bx, by, bz):
        if fma_cross:
            m = lambda a, b, c, d: (a.astype(np.float64) * b - (c * d).astype(f).astype(np.float64)).astype(f)
        else:
            m = lambda a, b, c, d: (a * b).astype(f) - (c * d).astype(f)
        return m(ay, bz, az, by), m(az, bx, ax, bz), m(ax, by, ay, bx)
    tx, ty, tz = cross(qx, qy, qz, bx, by, bz)
    tx, ty, tz = tx * f(2), ty * f(2), tz * f(2)
    ux, uy, uz = cross(qx, qy, qz, tx, ty, tz)
    rx = (bx + (qw * tx).astype(f)).astype(f) + ux
    ry = (by + (qw * ty).astype(f)).astype(f) + uy
    wx = (rx + rs[:, 0:1].numpy()).astype(f) + f(cfg.border_size)
    wy = (ry + rs[:, 1:2].numpy()).astype(f) + f(cfg.border_size)
    ix = (wx / f(cfg.horizontal_scale)).astype(np.int64)
    iy = (wy / f(cfg.horizontal_scale)).astype(np.int64)
    return np.stack([ix, iy], -1)
for fc in (False, True):
    for nm in ("sqrt", "sum4", "f64"):
        e = emulate(fc, nm)
        print("fma_cross", fc, "norm", nm, "index mismatches", int((e != idx_t).any(-1).sum()), "of", e.shape[0] * e.shape[1])

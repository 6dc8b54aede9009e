"""Shared test helpers: golden-fixture loading and tolerant comparison."""
import os

import numpy as np
import torch

from qa_b200 import synthetic
from qa_b200.config import BbcEnvConfig
from qa_b200.mocap import MocapTable

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# the parity bar of BASELINE.json: masks / indices bit-exact, floats within 1e-5 relative fp32
RTOL = 1e-5
ATOL = 2e-6      # absolute floor for values that cancel to ~0 (rewards near the >=0 clip, sin/cos of small angles)


def load_env_golden(name):
    """Returns (cfg, static, snap, draws, ref, meta) from tests/golden/bbc_env_<name>.npz."""
    z = np.load(os.path.join(GOLD, f"bbc_env_{name}.npz"))
    groups = {"static": {}, "snap": {}, "draws": {}, "ref": {}, "meta": {}}
    for k in z.files:
        g, key = k.split(".", 1)
        groups[g][key] = torch.from_numpy(z[k])
    meta = {k: v.item() for k, v in groups["meta"].items()}
    cfg = BbcEnvConfig(num_envs=int(meta["num_envs"]))
    # the 5 MB terrain grid is regenerated from the seed instead of being stored
    groups["static"]["height_samples"] = synthetic.make_static(cfg, seed=int(meta["seed"]))["height_samples"]
    return cfg, groups["static"], groups["snap"], groups["draws"], groups["ref"], meta


def mocap_table():
    return MocapTable.from_npz(os.path.join(GOLD, "mocap_lb_table.npz"))


def assert_close(name, got, want, rtol=RTOL, atol=ATOL):
    got = got.detach().cpu()
    want = want.detach().cpu()
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} != {tuple(want.shape)}"
    if want.dtype in (torch.bool, torch.uint8, torch.int32, torch.int64, torch.int16):
        assert torch.equal(got.to(torch.int64), want.to(torch.int64)), f"{name}: integer/mask mismatch"
        return
    g, w = got.double(), want.double()
    err = (g - w).abs()
    tol = atol + rtol * w.abs()
    bad = err > tol
    if bool(bad.any()):
        i = int(torch.argmax(err - tol))
        raise AssertionError(f"{name}: {int(bad.sum())}/{bad.numel()} beyond rtol={rtol} atol={atol}; "
                             f"worst flat idx {i}: got {g.flatten()[i].item():.9g} want {w.flatten()[i].item():.9g}")


def to_dev(d, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}

"""`Normalizer` / `RunningMeanStd` with the reference's pickle layout.

Checkpoints written by the reference pickle the `Normalizer` OBJECT (on_policy_runner.py:316), so the
attribute names (mean, var, count, epsilon, clip_obs) and the import path `rsl_rl.utils.utils.Normalizer`
are part of the on-disk contract: `install_pickle_alias()` registers this module under that path so that
`torch.load` of a shipped `model.pt` (e.g. tsc/weights/bbc/model.pt) resolves to these classes.
Moments stay float64 numpy like the reference (utils.py:51-103); a float32 device copy of mean / std is
cached for the per-step `normalize_torch` so that it is not re-uploaded on every call.
"""
import sys
import types

import numpy as np
import torch


class RunningMeanStd(object):
    def __init__(self, epsilon: float = 1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr) -> None:
        self.update_from_moments(np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count) -> None:
        """Chan et al. parallel merge of (mean, var, count), utils.py:68-83."""
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / tot
        self.mean = self.mean + delta * batch_count / tot
        self.var = m2 / tot
        self.count = tot
        self.__dict__.pop("_dev_cache", None)
        self.__dict__.pop("_dev_state", None)


class Normalizer(RunningMeanStd):
    def __init__(self, input_dim, epsilon=1e-4, clip_obs=10.0):
        super().__init__(shape=input_dim)
        self.epsilon = epsilon
        self.clip_obs = clip_obs

    def normalize(self, input):
        return np.clip((input - self.mean) / np.sqrt(self.var + self.epsilon), -self.clip_obs, self.clip_obs)

    def device_moments(self, device):
        """(mean, std) as float32 tensors on `device`, cached until the next update."""
        cache = self.__dict__.get("_dev_cache")
        if cache is None or cache[0] != str(device):
            mean = torch.tensor(self.mean, device=device, dtype=torch.float32)
            std = torch.sqrt(torch.tensor(self.var + self.epsilon, device=device, dtype=torch.float32))
            cache = (str(device), mean, std)
            self.__dict__["_dev_cache"] = cache
        return cache[1], cache[2]

    def normalize_torch(self, input, device):
        mean, std = self.device_moments(device)
        return torch.clamp((input - mean) / std, -self.clip_obs, self.clip_obs)

    # ---- device-side running moments (the reference round-trips every batch through numpy: 240 D2H syncs per iteration,
    #      gail.py:527-529) ------------------------------------------------------------------------------------------
    def _device_state(self, device):
        st = self.__dict__.get("_dev_state")
        if st is None or st[0] != str(device):
            mean32, std32 = self.device_moments(device)
            st = (str(device), torch.tensor(self.mean, device=device, dtype=torch.float64),
                  torch.tensor(self.var, device=device, dtype=torch.float64),
                  torch.tensor(float(self.count), device=device, dtype=torch.float64), mean32, std32)
            self.__dict__["_dev_state"] = st
        return st

    @torch.no_grad()
    def update_torch(self, x: torch.Tensor, world_size: int = 1) -> None:
        """RunningMeanStd.update (utils.py:63-83) on the device, float64 merge; refreshes the float32 (mean, std) that
        `normalize_torch` / K18 read IN PLACE (fixed addresses: CUDA-graph safe).  Call `sync_host()` before reading
        `.mean / .var / .count` or pickling."""
        _, mean, var, count, mean32, std32 = self._device_state(x.device)
        bm = x.mean(dim=0).double()                       # numpy's mean / var of a float32 batch are float32
        bv = x.var(dim=0, unbiased=False).double()
        bc = float(x.shape[0])
        if world_size > 1:
            # env-sharded ranks: pool the equal-sized per-rank batches into the moments of their union, so that every
            # rank holds the normaliser one big run would hold: mean_g = E_r[mean_r], var_g = E_r[var_r + mean_r^2] - mean_g^2
            import torch.distributed as dist
            pooled = torch.stack([bm, bv + bm.square()])
            dist.all_reduce(pooled, op=dist.ReduceOp.SUM)
            pooled /= world_size
            bm, bv, bc = pooled[0], pooled[1] - pooled[0].square(), bc * world_size
        delta = bm - mean
        tot = count + bc
        m2 = var * count + bv * bc + delta.square() * count * bc / tot
        mean.add_(delta * bc / tot)
        var.copy_(m2 / tot)
        count.copy_(tot)
        mean32.copy_(mean.float())
        std32.copy_(torch.sqrt((var + self.epsilon).float()))
        self.__dict__["_dev_dirty"] = True

    def load_moments(self, mean, var, count, epsilon=None, clip_obs=None) -> None:
        """Overwrite the running moments IN PLACE: the host arrays and, when they exist, the device tensors that captured
        CUDA graphs and K18 read (their addresses do not change)."""
        self.mean = np.array(mean, dtype=np.float64).reshape(self.mean.shape)
        self.var = np.array(var, dtype=np.float64).reshape(self.var.shape)
        self.count = float(count)
        if epsilon is not None:
            self.epsilon = epsilon
        if clip_obs is not None:
            self.clip_obs = clip_obs
        cache, st = self.__dict__.get("_dev_cache"), self.__dict__.get("_dev_state")
        if st is not None:
            dev = st[1].device
            st[1].copy_(torch.tensor(self.mean, dtype=torch.float64, device=dev))
            st[2].copy_(torch.tensor(self.var, dtype=torch.float64, device=dev))
            st[3].fill_(self.count)
            st[4].copy_(st[1].float())
            st[5].copy_(torch.sqrt((st[2] + self.epsilon).float()))
            self.__dict__["_dev_dirty"] = False
        elif cache is not None:
            cache[1].copy_(torch.tensor(self.mean, dtype=torch.float32, device=cache[1].device))
            cache[2].copy_(torch.sqrt(torch.tensor(self.var + self.epsilon, dtype=torch.float32, device=cache[2].device)))

    def sync_host(self) -> None:
        st = self.__dict__.get("_dev_state")
        if st is not None and self.__dict__.get("_dev_dirty"):
            self.mean, self.var, self.count = st[1].cpu().numpy(), st[2].cpu().numpy(), float(st[3].item())
            self.__dict__["_dev_dirty"] = False

    def __getstate__(self):
        self.sync_host()
        d = dict(self.__dict__)
        d.pop("_dev_cache", None)
        d.pop("_dev_state", None)
        d.pop("_dev_dirty", None)
        return d


def install_pickle_alias() -> None:
    """Make `rsl_rl.utils.utils.Normalizer` importable for un-pickling reference checkpoints."""
    if "rsl_rl.utils.utils" in sys.modules:
        return
    pkg = sys.modules.setdefault("rsl_rl", types.ModuleType("rsl_rl"))
    sub = sys.modules.setdefault("rsl_rl.utils", types.ModuleType("rsl_rl.utils"))
    me = sys.modules[__name__]
    sys.modules["rsl_rl.utils.utils"] = me
    pkg.utils = sub
    sub.utils = me


def reference_picklable(n):
    """The running normaliser as an object that pickles under the reference's class path `rsl_rl.utils.utils.Normalizer`
    (bbc/rsl_rl/utils/utils.py:86), so that a checkpoint written here un-pickles inside the reference (its `load` assigns the
    object as is, on_policy_runner.py:327) as well as here (through `install_pickle_alias`)."""
    if n is None:
        return None
    install_pickle_alias()
    mod = sys.modules["rsl_rl.utils.utils"]
    if mod is sys.modules[__name__]:
        Normalizer.__module__ = "rsl_rl.utils.utils"               # resolves to this very class through the alias
        return n
    obj = mod.Normalizer.__new__(mod.Normalizer)                   # the real reference package is importable
    obj.__dict__.update(n.__getstate__() if hasattr(n, "__getstate__") else n.__dict__)
    return obj

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

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_dev_cache", None)
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

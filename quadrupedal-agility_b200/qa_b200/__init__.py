"""qa_b200 -- B200-native (sm_100a) hot path of NJU-RLC/quadrupedal-agility.

Host side of the drop-in: Python classes that keep the reference's `VecEnv.step/reset`,
`RolloutStorage`, `ActorCritic` and `OnPolicyRunner` API over
PyTorch tensors, calling hand-written CUDA through the C-ABI library `libqa_b200.so`
(declared in `include/qa_b200.h`).  There is no CPU fallback: every hot-path entry point
raises if the CUDA library is missing.
"""
from .version import __version__  # noqa: F401

"""Stub of the un-installable `isaacgym` package (Isaac Gym Preview 4, closed binary).

TEST INFRASTRUCTURE ONLY.  It exists so that the reference's `legged_gym` /
`rsl_rl` modules can be *imported* in the survey container (no GPU, no
IsaacGym) by `oracle/gen_golden.py`.  Only `torch_utils` carries arithmetic
(a restatement of the public Preview-4 helpers, see that file); everything
else is an attribute sink that is never executed on the hot path.
"""
from . import gymapi, gymtorch, gymutil, terrain_utils  # noqa: F401

"""Stub for `pybullet_utils` (pybullet 3.2.5, requirements.txt:8): only the non-batched
`blend_frame_pose` and load-time `pose3d` helpers touch it; the batched hot path does not.
TEST INFRASTRUCTURE ONLY."""
from . import transformations  # noqa: F401

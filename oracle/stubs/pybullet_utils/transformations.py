"""Minimal `pybullet_utils.transformations` surface used at load time by the reference's
`rsl_rl/datasets/pose3d.py` / `motion_util.py`.  Off the batched hot path."""
import math

import numpy as np

_EPS = np.finfo(float).eps * 4.0


def quaternion_slerp(quat0, quat1, fraction, spin=0, shortestpath=True):
    q0 = np.array(quat0[:4], dtype=np.float64)
    q1 = np.array(quat1[:4], dtype=np.float64)
    q0 /= np.linalg.norm(q0)
    q1 /= np.linalg.norm(q1)
    if fraction == 0.0:
        return q0
    elif fraction == 1.0:
        return q1
    d = np.dot(q0, q1)
    if abs(abs(d) - 1.0) < _EPS:
        return q0
    if shortestpath and d < 0.0:
        d = -d
        q1 *= -1.0
    angle = math.acos(d) + spin * math.pi
    if abs(angle) < _EPS:
        return q0
    isin = 1.0 / math.sin(angle)
    q0 *= math.sin((1.0 - fraction) * angle) * isin
    q1 *= math.sin(fraction * angle) * isin
    q0 += q1
    return q0


def _raise(*a, **k):
    raise RuntimeError("pybullet_utils stub: function outside the oracle's scope")


quaternion_from_euler = euler_from_quaternion = quaternion_matrix = quaternion_from_matrix = _raise
quaternion_multiply = quaternion_conjugate = quaternion_inverse = quaternion_about_axis = _raise
euler_from_matrix = euler_matrix = _raise

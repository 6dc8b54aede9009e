"""Attribute sink for `isaacgym.gymapi` (never executed by the oracle)."""


class _Sink:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Sink()

    def __call__(self, *a, **k):
        return _Sink()


SIM_PHYSX = 1
SIM_FLEX = 0
UP_AXIS_Z = 1
KEY_ESCAPE = KEY_V = KEY_W = KEY_S = KEY_A = KEY_D = KEY_T = KEY_H = KEY_L = 0
KEY_SPACE = KEY_1 = KEY_2 = KEY_3 = KEY_4 = KEY_5 = KEY_LEFT = KEY_RIGHT = KEY_UP = KEY_DOWN = 0
IMAGE_DEPTH = 0
ENV_SPACE = 0
DOMAIN_SIM = 0


class SimParams(_Sink):
    pass


class Vec3(_Sink):
    pass


class Quat(_Sink):
    pass


class Transform(_Sink):
    pass


class PlaneParams(_Sink):
    pass


class HeightFieldParams(_Sink):
    pass


class TriangleMeshParams(_Sink):
    pass


class AssetOptions(_Sink):
    pass


class CameraProperties(_Sink):
    pass


def acquire_gym(*a, **k):
    return _Sink()

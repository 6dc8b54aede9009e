"""Attribute sink for `isaacgym.terrain_utils` (terrain generation is init-time, out of scope)."""


class SubTerrain:
    def __init__(self, *a, **k):
        raise RuntimeError("isaacgym stub: terrain generation is outside the oracle's scope")


def convert_heightfield_to_trimesh(*a, **k):
    raise RuntimeError("isaacgym stub: terrain generation is outside the oracle's scope")

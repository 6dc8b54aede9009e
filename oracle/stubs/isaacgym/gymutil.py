"""Attribute sink for `isaacgym.gymutil`."""


def parse_device_str(s):
    if ":" in s:
        a, b = s.split(":")
        return a, int(b)
    return s, 0


def parse_arguments(*a, **k):
    raise RuntimeError("isaacgym stub: CLI parsing is outside the oracle's scope")


def parse_sim_config(*a, **k):
    return None


class WireframeSphereGeometry:
    def __init__(self, *a, **k):
        pass


def draw_lines(*a, **k):
    pass

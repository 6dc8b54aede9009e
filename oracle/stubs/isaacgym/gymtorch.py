"""Attribute sink for `isaacgym.gymtorch` (tensors are injected directly by the oracle)."""


def wrap_tensor(t):
    return t


def unwrap_tensor(t):
    return t

"""Stub for the stdlib `turtle` module: the reference has a stray
`from turtle import forward` (bbc/rsl_rl/modules/estimator.py:1) that would pull in
tkinter.  TEST INFRASTRUCTURE ONLY."""


def forward(*a, **k):
    raise RuntimeError("turtle stub")

"""Stub for `matplotlib` (imported by the reference's play-time logger only)."""

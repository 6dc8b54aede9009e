"""Stub: scikit-image is not installable here; the reference only uses skimage.draw.polygon when it BUILDS obstacle
terrain (tsc/legged_gym/utils/obstacle.py), which is outside the hot path."""

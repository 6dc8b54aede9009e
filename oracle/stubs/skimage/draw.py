def polygon(*args, **kwargs):
    raise NotImplementedError("skimage.draw.polygon stub: obstacle terrain generation is outside the hot path")

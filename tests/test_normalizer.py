"""Running normaliser (bbc/rsl_rl/utils/utils.py:51-103): the device-side merge `update_torch` must track the reference's numpy
`update` (float32 batch moments, float64 running moments), and `normalize_torch` the reference's clip((x - mean) / sqrt(var + eps))."""
import numpy as np
import torch

from qa_b200.rsl_rl.utils import Normalizer


def test_update_torch_tracks_the_numpy_update_and_normalises_like_the_reference():
    g = torch.Generator().manual_seed(0)
    a, b = Normalizer(98), Normalizer(98)
    for n in (40, 1228, 7):
        x = 3.0 * torch.randn(n, 98, generator=g) + 1.5
        a.update(x.numpy())                      # the reference's path: Normalizer.update(x.cpu().numpy()) (gail.py:527-529)
        b.update_torch(x)
    b.sync_host()
    assert a.count == b.count
    # the batch moments are fp32 sums over up to 1228 rows: numpy's pairwise and torch's blocked summation differ in the last bits
    assert np.allclose(a.mean, b.mean, rtol=5e-6, atol=1e-6) and np.allclose(a.var, b.var, rtol=2e-5, atol=1e-6)
    x = 3.0 * torch.randn(16, 98, generator=g) + 1.5
    want = np.clip((x.numpy() - a.mean) / np.sqrt(a.var + a.epsilon), -a.clip_obs, a.clip_obs)
    got = b.normalize_torch(x, x.device)
    assert np.allclose(got.numpy(), want, rtol=1e-4, atol=1e-5)
    big = b.normalize_torch(torch.full((1, 98), 1e6), "cpu")
    assert float(big.max()) == b.clip_obs

"""GPU parity of K5 (GAE warp-scan + advantage normalisation) against the reference's own
RolloutStorage.compute_returns outputs (golden) and the CPU oracle at the BASELINE size."""
import numpy as np
import pytest
import torch

import trainer as OT
from helpers import GOLD, assert_close
from qa_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_gae(rewards, values, dones, last_values, gamma, lam):
    r, v, d, lv = (x.to(DEV).contiguous() for x in (rewards, values, dones, last_values))
    ret, adv = torch.empty_like(r), torch.empty_like(r)
    ws = torch.zeros(8, dtype=torch.float64, device=DEV)
    ops.gae(r, v, d, lv, ret, adv, ws, gamma, lam)
    torch.cuda.synchronize()
    return ret, adv


@pytest.mark.parametrize("name", ["gae_t24_n64", "gae_t24_n100", "gae_t5_n33"])
def test_gae_matches_reference_golden(name):
    z = np.load(f"{GOLD}/trainer_{name}.npz")
    t = lambda k: torch.from_numpy(z[k])          # noqa: E731
    ret, adv = run_gae(t("rewards"), t("values"), t("dones"), t("last_values"), float(z["gamma"]), float(z["lam"]))
    assert_close("returns", ret, t("ref_returns"))
    assert_close("advantages", adv, t("ref_advantages"), atol=1e-5)


def test_gae_matches_oracle_at_baseline_size():
    T, N = 24, 4096
    g = torch.Generator().manual_seed(1234)
    rewards = 0.05 * torch.rand(T, N, 1, generator=g)
    values = 1.5 + 0.5 * torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.015).byte()
    last_values = 1.5 + 0.5 * torch.randn(N, 1, generator=g)
    want_ret, want_adv = OT.compute_returns(rewards, values, dones, last_values, 0.99, 0.95)
    ret, adv = run_gae(rewards, values, dones, last_values, 0.99, 0.95)
    assert_close("returns", ret, want_ret)
    assert_close("advantages", adv, want_adv, atol=1e-5)
    # properties: normalised advantages have zero mean / unit (unbiased) std; a second call on the same
    # workspace gives the same answer (the workspace is re-zeroed by the library)
    a = adv.double()
    assert abs(float(a.mean())) < 1e-5 and abs(float(a.std()) - 1.0) < 1e-5
    ret2, adv2 = run_gae(rewards, values, dones, last_values, 0.99, 0.95)
    assert torch.equal(ret, ret2)
    assert_close("advantages again", adv2, adv, atol=1e-6)


def test_gae_all_done_and_none_done():
    T, N = 24, 96
    g = torch.Generator().manual_seed(5)
    rewards, values = torch.rand(T, N, 1, generator=g), torch.randn(T, N, 1, generator=g)
    last_values = torch.randn(N, 1, generator=g)
    for fill in (0, 1):
        dones = torch.full((T, N, 1), fill, dtype=torch.uint8)
        want_ret, want_adv = OT.compute_returns(rewards, values, dones, last_values, 0.99, 0.95)
        ret, adv = run_gae(rewards, values, dones, last_values, 0.99, 0.95)
        assert_close("returns", ret, want_ret)
        assert_close("advantages", adv, want_adv, atol=1e-5)
    with pytest.raises(RuntimeError, match="QA_ERANGE"):
        big = torch.zeros(40, 8, 1)
        run_gae(big, big, big.byte(), torch.zeros(8, 1), 0.99, 0.95)

"""Pin `oracle/trainer.py` against the UNMODIFIED reference trainer classes and write the trainer
golden vectors (build container only; needs /root/reference).  Called by `oracle/gen_golden.py`.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

import trainer as OT  # noqa: E402
from ref_harness import import_reference, REFERENCE_ROOT  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def synth_rollout(T, N, seed, done_frac=0.02):
    g = torch.Generator().manual_seed(seed)
    rewards = 0.05 * torch.rand(T, N, 1, generator=g) + 0.01 * torch.randn(T, N, 1, generator=g)
    values = 1.5 + 0.5 * torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < done_frac).byte()
    last_values = 1.5 + 0.5 * torch.randn(N, 1, generator=g)
    return rewards, values, dones, last_values


def gae_case(ref, name, T, N, seed, gamma=0.99, lam=0.95, save=True):
    rewards, values, dones, last_values = synth_rollout(T, N, seed)
    st = ref.RolloutStorage(N, T, [4], [4], [2], device="cpu")
    st.rewards.copy_(rewards)
    st.values.copy_(values)
    st.dones.copy_(dones)
    st.compute_returns(last_values, gamma, lam)                     # the reference's own code
    ret_o, adv_o = OT.compute_returns(rewards, values, dones, last_values, gamma, lam)
    assert torch.equal(st.returns, ret_o), "GAE returns: oracle != reference"
    assert torch.equal(st.advantages, adv_o), "GAE advantages: oracle != reference"
    print(f"  {name}: T={T} N={N} dones={int(dones.sum())}: oracle == reference RolloutStorage.compute_returns (bit-exact)")
    if save:
        np.savez_compressed(os.path.join(GOLD, f"trainer_{name}.npz"), rewards=rewards.numpy(), values=values.numpy(),
                            dones=dones.numpy(), last_values=last_values.numpy(), gamma=np.array(gamma),
                            lam=np.array(lam), ref_returns=st.returns.numpy(), ref_advantages=st.advantages.numpy())


def main():
    torch.set_num_threads(1)
    ref = import_reference("bbc")
    print("[gen_golden] trainer: GAE (reference RolloutStorage vs oracle)")
    gae_case(ref, "gae_t24_n64", 24, 64, 11)
    gae_case(ref, "gae_t24_n100", 24, 100, 12)          # ragged tile (N % 32 != 0)
    gae_case(ref, "gae_t5_n33", 5, 33, 13)              # short horizon
    gae_case(ref, "gae_t24_n4096_check", 24, 4096, 1234, save=False)
    try:
        import gen_golden_policy
    except ImportError:
        return
    gen_golden_policy.main(ref)


if __name__ == "__main__":
    main()

"""Pin the DAgger step (history-encoder adaptation, bbc/rsl_rl/algorithms/gail.py:543-575) of `oracle/trainer.py` against
the UNMODIFIED reference `SSInfoGAIL.update_dagger` and write tests/golden/trainer_dagger_seed3.npz.  Build container only."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

import trainer as OT  # noqa: E402
from gen_golden_policy import build_reference_nets, make_alg  # noqa: E402
from qa_b200 import synthetic  # noqa: E402
from ref_harness import import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref = import_reference("bbc")
    torch.set_num_threads(1)
    w = synthetic.make_weights(3)
    z = np.load(os.path.join(GOLD, "bbc_env_n64a.npz"))
    obs = torch.from_numpy(z["ref.obs_buf"]).clone()
    N = obs.shape[0]
    ac, est, disc, norm, env = build_reference_nets(ref, w)
    alg = make_alg(ref, ac, est)
    alg.optim_hist_encoder = torch.optim.Adam(ac.history_encoder.parameters(), lr=1e-4)
    alg.max_grad_norm, alg.num_learning_epochs, alg.num_mini_batches = 1.0, 2, 1
    st = ref.RolloutStorage(N, 1, [671], [671], [12], device="cpu")
    st.observations[0] = obs
    alg.storage = st
    loss_ref = alg.update_dagger()                                       # two epochs over the one minibatch
    sd = {k: v.clone().requires_grad_(True) for k, v in w["ac"].items()}
    enc = [v for k, v in sd.items() if k.startswith("history_encoder.")]
    opt = torch.optim.Adam(enc, lr=1e-4)
    losses = []
    for _ in range(2):
        with torch.no_grad():
            priv = OT.infer_priv_latent(sd, obs[:, 61:90])
        hist = OT.infer_hist_latent(sd, obs[:, 90:660])
        loss = (priv.detach() - hist).norm(p=2, dim=1).mean()
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(enc, 1.0)
        opt.step()
        losses.append(float(loss))
    assert abs(sum(losses) / 2 - loss_ref) <= 1e-6 * abs(loss_ref) + 1e-8, (losses, loss_ref)
    ref_sd = {k: v.detach() for k, v in ac.state_dict().items()}
    for k in ref_sd:
        assert torch.allclose(sd[k].detach(), ref_sd[k], rtol=1e-5, atol=1e-7), k
    enc_flat = torch.cat([v.reshape(-1) for k, v in ref_sd.items() if k.startswith("history_encoder.")])
    print(f"  dagger x2: oracle == reference update_dagger (mean loss {loss_ref:.5f})")
    np.savez_compressed(os.path.join(GOLD, "trainer_dagger_seed3.npz"), mean_loss=np.array(loss_ref), obs=obs.numpy(),
                        encoder_params=enc_flat.numpy(), other_params_unchanged=np.array(1))
    print("wrote tests/golden/trainer_dagger_seed3.npz")


if __name__ == "__main__":
    main()

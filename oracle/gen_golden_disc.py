"""Pin the discriminator-update half of `oracle/trainer.py` against the UNMODIFIED reference
`SSInfoGAIL.update_ss_info_gail` (bbc/rsl_rl/algorithms/gail.py:415-541) and write tests/golden/trainer_disc_seed3.npz.
Build container only.  python oracle/gen_golden_disc.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

import trainer as OT  # noqa: E402
from qa_b200 import synthetic  # noqa: E402
from ref_harness import import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
STRIDE = 53


def main():
    ref = import_reference("bbc")
    torch.set_num_threads(1)
    w = synthetic.make_weights(3)
    g = torch.Generator().manual_seed(9)
    B = 96
    env = types.SimpleNamespace(task_obs_weight_decay=True, task_obs_weight=0.8, dim_c=5, prior_parameters=torch.full((5,), 0.2))
    disc = ref.discriminator.Discriminator(env, 98, 49, 5, 0.02, "MSELoss", None, 1.0, 0.01, 0.2, 0.2, 2, 2, 0.0, [512, 256], "cpu")
    disc.load_state_dict(w["disc"])
    norm = ref.utils.Normalizer(98)
    norm.mean, norm.var, norm.count = w["norm_mean"].numpy().copy(), w["norm_var"].numpy().copy(), 5000.0
    G = ref.gail.SSInfoGAIL
    alg = G.__new__(G)
    alg.device, alg.env, alg.disc, alg.disc_normalizer = "cpu", env, disc, norm
    alg.disc_obs_len, alg.num_disc_obs, alg.obs_disc_weight_step, alg.dim_c = 2, 49, 0.0, 5
    alg.disc_loss_function = "MSELoss"
    alg.CE_loss, alg.MSELoss, alg.L1Loss = torch.nn.CrossEntropyLoss(), torch.nn.MSELoss(), torch.nn.L1Loss()
    alg.ss_coef, alg.info_max_coef_on, alg.disc_coef, alg.us_coef = 1.0, 0.3, 1.0, 1.0
    alg.disc_grad_penalty, alg.disc_logit_reg, alg.disc_weight_decay, alg.prior_soft_coef = 0.1, 0.05, 0.0001, 1e-3
    grp = lambda ps, name: {'params': ps, 'weight_decay': 1e-3, 'momentum': 0.9, 'name': name}      # noqa: E731  gail.py:107-122
    alg.optim_d = torch.optim.Adam([grp(disc.trunk.parameters(), 'trunk'), grp(disc.linear.parameters(), 'head')], lr=5e-4)
    alg.optim_q_eps = torch.optim.Adam([grp(disc.trunk.parameters(), 'trunk'), grp(disc.encoder_eps.parameters(), 'encoder_eps')], lr=1e-3)
    alg.optim_q_c = torch.optim.Adam([grp(disc.trunk.parameters(), 'trunk'), grp(disc.classifier.parameters(), 'classifier')], lr=1e-3)
    std0 = torch.tensor([0.02, 0.5, 0.9] * 4)
    alg.actor_critic = types.SimpleNamespace(fixed_std=False, std=torch.nn.Parameter(std0.clone()))
    alg.min_std = torch.tensor([0.05, 0.02, 0.05] * 4) * 1.5
    pol = torch.randn(B, 98, generator=g)
    pol_eps = 2 * torch.rand(B, 1, generator=g) - 1
    pol_c = torch.nn.functional.one_hot(torch.randint(0, 5, (B,), generator=g), 5).float()
    exp_lb, lab_lb = torch.randn(B, 98, generator=g), torch.randint(0, 5, (B,), generator=g)
    exp_ulb = torch.randn(B, 98, generator=g)
    out = {}
    sd = {k: v.clone().requires_grad_(True) for k, v in w["disc"].items()}
    opts = OT.disc_optimizers(sd)
    mean, var, count = w["norm_mean"].clone(), w["norm_var"].clone(), 5000.0
    prior = torch.full((5,), 0.2)
    for step in range(2):                                                    # two steps: moments of the 3 Adams interact
        ret = alg.update_ss_info_gail((pol, pol_eps, pol_c), (exp_lb, lab_lb), exp_ulb)      # the reference's own code
        xs = [OT.disc_prepare(x, 0.8, mean, var) for x in (pol, exp_lb, exp_ulb)]
        L = OT.disc_losses(sd, xs[0], pol_eps, pol_c, xs[1], lab_lb, xs[2], info_max_coef_on=0.3)
        prior = L["pred_c_ulb_mean"] * 1e-3 + prior * (1 - 1e-3)
        for o in opts:
            o.zero_grad()
        L["loss"].backward()
        for o in opts:
            o.step()
        for x in xs:
            mean, var, count = OT.normalizer_update(mean, var, count, x)
        names = ("ss_loss", "info_max_loss", "disc_loss", "us_loss", "grad_pen_loss", "disc_logit_loss", "disc_weight_decay",
                 "acc_lb", "acc_pi", "acc_exp", "acc_ulb")
        for k, rv in zip(names, ret):
            assert torch.allclose(L[k], rv.detach(), rtol=1e-5, atol=1e-7), (step, k, float(L[k]), float(rv))
            out[f"s{step}.{k}"] = rv.detach().clone()
    ref_sd = {k: v.detach() for k, v in disc.state_dict().items()}
    for k in ref_sd:
        d = (sd[k].detach() - ref_sd[k]).abs()
        assert float(d.max()) <= 2.5e-3 and float((d > 1e-6).float().mean()) < 5e-3, (k, float(d.max()), float((d > 1e-6).float().mean()))
    assert np.allclose(mean.numpy(), norm.mean, rtol=1e-6, atol=1e-9) and np.allclose(var.numpy(), norm.var, rtol=1e-6, atol=1e-9)
    assert abs(count - norm.count) < 1e-9
    assert torch.allclose(prior, env.prior_parameters, rtol=1e-6, atol=1e-9)
    assert torch.equal(alg.actor_critic.std.data, std0.clamp(min=alg.min_std))
    print(f"  disc update x2: oracle == reference update_ss_info_gail (ss={float(ret[0]):.4f}, disc={float(ret[2]):.4f}, "
          f"gp={float(ret[4]):.4f}, acc_lb={float(ret[7]):.2f})")
    out.update({"post.params_sampled": torch.cat([v.reshape(-1) for v in ref_sd.values()])[::STRIDE].clone(),
                "post.norm_mean": torch.from_numpy(norm.mean.copy()), "post.norm_var": torch.from_numpy(norm.var.copy()),
                "post.norm_count": torch.tensor(norm.count, dtype=torch.float64), "post.prior": env.prior_parameters.clone(),
                "post.std": alg.actor_critic.std.data.clone(),
                "in.pol": pol, "in.pol_eps": pol_eps, "in.pol_c": pol_c, "in.exp_lb": exp_lb, "in.lab_lb": lab_lb,
                "in.exp_ulb": exp_ulb, "in.std0": std0, "in.min_std": alg.min_std, "in.param_stride": torch.tensor(STRIDE)})
    np.savez_compressed(os.path.join(GOLD, "trainer_disc_seed3.npz"), **{k: v.numpy() for k, v in out.items()})
    print("wrote tests/golden/trainer_disc_seed3.npz")


if __name__ == "__main__":
    main()

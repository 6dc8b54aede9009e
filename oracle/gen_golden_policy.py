"""Pin the policy / discriminator-reward / PPO-step half of `oracle/trainer.py` against the UNMODIFIED
reference classes (bbc/rsl_rl), and write small golden vectors.  Build container only."""
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
from qa_b200.config import bbc_train_cfg  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
STRIDE = 97            # parameter sub-sampling stride for the post-step fixture


def build_reference_nets(ref, w):
    cfg = bbc_train_cfg()
    ac = ref.actor_critic.ActorCritic(101, 671, 12, 57, 10, 4, 29, 11, **cfg["policy"])
    ac.load_state_dict(w["ac"])
    est = ref.estimator.Estimator(input_dim=57, output_dim=4, hidden_dims=[128, 64])
    est.load_state_dict(w["est"])
    env = types.SimpleNamespace(task_obs_weight_decay=True, task_obs_weight=0.7, dim_c=5)
    disc = ref.discriminator.Discriminator(env, 98, 49, 5, 0.02, "MSELoss", None, 1.0, 0.01, 0.2, 0.2, 2, 2, 0.0,
                                           [512, 256], "cpu")
    disc.load_state_dict(w["disc"])
    norm = ref.utils.Normalizer(98)
    norm.mean, norm.var = w["norm_mean"].numpy().copy(), w["norm_var"].numpy().copy()
    return ac, est, disc, norm, env


def make_alg(ref, ac, est, lr_ac=1e-3):
    """SSInfoGAIL.__new__ + exactly the attributes act / update_actor_critic read (gail.py:176-197, 328-413)."""
    G = ref.gail.SSInfoGAIL
    alg = G.__new__(G)
    cfg = bbc_train_cfg()["algorithm"]
    alg.device = "cpu"
    alg.actor_critic, alg.estimator = ac, est
    alg.num_prop, alg.num_explicit, alg.num_latent, alg.num_hist, alg.num_command = 57, 4, 29, 10, 11
    alg.train_with_estimated_explicit = True
    alg.optim_ac = torch.optim.Adam([{'params': ac.parameters(), 'name': 'actor_critic'}], lr=lr_ac)
    alg.optim_estimator = torch.optim.Adam(est.parameters(), lr=1e-4)
    alg.priv_reg_coef_schedual, alg.priv_reg_counter = cfg["priv_reg_coef_schedual"], 1500
    alg.desired_kl, alg.schedule, alg.lr_ac = 0.01, "adaptive", lr_ac
    alg.clip_param, alg.use_clipped_value_loss = 0.2, True
    alg.surrogate_loss_coef, alg.value_loss_coef = 2.0, 5.0
    alg.bounds_loss_coef, alg.entropy_coef, alg.max_grad_norm = 0.0, 0.01, 1.0
    alg.transition = ref.RolloutStorage.Transition(64, [671], [671], [12], "cpu")
    return alg


def sample_params(sd):
    return torch.cat([v.reshape(-1) for v in sd.values()])[::STRIDE].clone()


def main(ref):
    torch.set_num_threads(1)
    torch.manual_seed(0)
    print("[gen_golden] policy / disc reward / PPO step (reference modules vs oracle)")
    z = np.load(os.path.join(GOLD, "bbc_env_n64a.npz"))
    obs = torch.from_numpy(z["ref.obs_buf"]).clone()                  # (64,671) real reference observations
    disc_now = torch.from_numpy(z["ref.obs_disc_buf"]).clone()
    disc_prev = torch.from_numpy(z["snap.obs_disc_buf"]).clone()
    N = obs.shape[0]
    g = torch.Generator().manual_seed(5)
    for tag, w in (("ckpt", None), ("seed3", synthetic.make_weights(3))):
        if w is None:                                                 # the shipped checkpoint (container only)
            ck = torch.load(os.path.join(ref.root, "..", "tsc", "weights", "bbc", "model.pt"), map_location="cpu",
                            weights_only=False)
            w = dict(ac=ck["actor_critic"], est=ck["estimator"], disc=ck["disc"],
                     norm_mean=torch.from_numpy(ck["disc_normalizer"].mean), norm_var=torch.from_numpy(ck["disc_normalizer"].var))
        ac, est, disc, norm, env = build_reference_nets(ref, w)
        alg = make_alg(ref, ac, est)
        draw = torch.randn(N, 12, generator=g)
        # ---- act (gail.py:176-197) with the N(0,1) draw injected into Normal.sample -------------------------
        Normal = torch.distributions.Normal
        orig = Normal.sample
        Normal.sample = lambda self, sample_shape=torch.Size(): (self.loc + self.scale * draw).detach()
        out = {}
        for he in (False, True):
            with torch.inference_mode():
                a = alg.act(obs.clone(), obs.clone(), hist_encoding=he)
            tr = alg.transition
            o = OT.act(w["ac"], w["est"], obs, obs, draw, hist_encoding=he)
            for k, rv in (("actions", a), ("values", tr.values), ("actions_log_prob", tr.actions_log_prob),
                          ("action_mean", tr.action_mean), ("action_sigma", tr.action_sigma)):
                assert torch.allclose(o[k], rv, rtol=1e-6, atol=1e-6), (tag, he, k, float((o[k] - rv).abs().max()))
                out[f"act{int(he)}.{k}"] = rv.clone()
        Normal.sample = orig
        # ---- predict_disc_reward (discriminator.py:71-118) -----------------------------------------------------
        hist = torch.stack([disc_prev, disc_now], dim=1)
        rew_t = torch.from_numpy(z["ref.rew_buf"]).clone().unsqueeze(1)
        r_ref = disc.predict_disc_reward(rew_t, obs, hist, normalizer=norm)
        r_or = OT.predict_disc_reward(w["disc"], rew_t, obs, hist, w["norm_mean"], w["norm_var"], 0.02, 0.7)
        for k, a_, b_ in zip(("rewards", "reward_i", "reward_us", "reward_ss", "reward_t"), r_ref, r_or):
            assert a_.dtype == b_.dtype, (k, a_.dtype, b_.dtype)
            assert torch.allclose(a_, b_, rtol=1e-6, atol=1e-7), (tag, k, float((a_ - b_).abs().max()))
            out[f"disc.{k}"] = a_.clone()
        assert r_ref[0].dtype == torch.float64 and r_ref[3].dtype == torch.float64      # SURVEY a' quirk
        # ---- one PPO minibatch step (gail.py:328-413) -----------------------------------------------------------
        o0 = OT.act(w["ac"], w["est"], obs, obs, draw, hist_encoding=False)
        batch = dict(obs=obs, critic_obs=obs, actions=o0["actions"], target_values=o0["values"],
                     advantages=torch.randn(N, 1, generator=g), returns=o0["values"] + 0.3 * torch.randn(N, 1, generator=g),
                     old_actions_log_prob=(o0["actions_log_prob"] + 0.05 * torch.randn(N, generator=g)).unsqueeze(1),
                     old_mu=o0["action_mean"] + 0.05 * torch.randn(N, 12, generator=g), old_sigma=o0["action_sigma"] * 1.05)
        sample = (batch["obs"], batch["critic_obs"], batch["actions"], batch["target_values"], batch["advantages"],
                  batch["returns"], batch["old_actions_log_prob"], batch["old_mu"], batch["old_sigma"], (None, None), None)
        Normal.sample = lambda self, sample_shape=torch.Size(): self.loc.detach()
        losses_ref = alg.update_actor_critic(sample)
        Normal.sample = orig
        # oracle: same step with torch.optim.Adam on leaf copies
        sd_ac = {k: v.clone().requires_grad_(True) for k, v in w["ac"].items()}
        sd_est = {k: v.clone().requires_grad_(True) for k, v in w["est"].items()}
        L = OT.ppo_losses(sd_ac, sd_est, batch, priv_reg_coef=OT.priv_reg_coef(1500))
        opt_e = torch.optim.Adam(list(sd_est.values()), lr=1e-4)
        L["estimator_loss"].backward()
        torch.nn.utils.clip_grad_norm_(list(sd_est.values()), 1.0)
        opt_e.step()
        lr_new = OT.adaptive_lr(1e-3, float(L["kl_mean"]))
        opt_a = torch.optim.Adam(list(sd_ac.values()), lr=lr_new)
        L["ppo_loss"].backward()
        torch.nn.utils.clip_grad_norm_(list(sd_ac.values()), 1.0)
        opt_a.step()
        names = ("surrogate_loss", "value_loss", "b_loss", "entropy", "priv_reg_loss", "estimator_loss")
        for k, rv in zip(names, losses_ref):
            assert torch.allclose(L[k], rv.mean(), rtol=1e-5, atol=1e-7), (tag, k, float(L[k]), float(rv.mean()))
            out[f"ppo.{k}"] = rv.mean().detach().clone()
        assert abs(alg.lr_ac - lr_new) < 1e-12, (alg.lr_ac, lr_new)
        ref_ac = {k: v.detach() for k, v in ac.state_dict().items()}
        ref_est = {k: v.detach() for k, v in est.state_dict().items()}
        for k in ref_ac:
            assert torch.allclose(sd_ac[k].detach(), ref_ac[k], rtol=1e-5, atol=1e-7), (tag, "post-step", k)
        for k in ref_est:
            assert torch.allclose(sd_est[k].detach(), ref_est[k], rtol=1e-5, atol=1e-7), (tag, "post-step est", k)
        out["ppo.kl_mean"] = L["kl_mean"].detach().clone()
        out["ppo.lr_new"] = torch.tensor(lr_new, dtype=torch.float64)
        out["ppo.ac_params_sampled"] = sample_params(ref_ac)
        out["ppo.est_params_sampled"] = sample_params(ref_est)
        print(f"  {tag}: act / predict_disc_reward / update_actor_critic: oracle == reference "
              f"(kl={float(L['kl_mean']):.4f}, lr {1e-3}->{lr_new:.6f})")
        if tag != "ckpt":                                             # checkpoint weights do not travel
            save = {k: v.numpy() for k, v in out.items()}
            save.update({"in.obs": obs.numpy(), "in.draw": draw.numpy(), "in.disc_hist": hist.numpy(),
                         "in.rew_t": rew_t.numpy(), "in.task_obs_weight": np.array(0.7), "in.weights_seed": np.array(3),
                         "in.priv_reg_counter": np.array(1500), "in.param_stride": np.array(STRIDE)})
            for k in ("advantages", "returns", "old_actions_log_prob", "old_mu", "old_sigma", "target_values", "actions"):
                save["in.batch." + k] = batch[k].numpy()
            np.savez_compressed(os.path.join(GOLD, "trainer_policy_seed3.npz"), **save)


if __name__ == "__main__":
    from ref_harness import import_reference
    main(import_reference("bbc"))

"""Opt-in batched discriminator minibatch step (`QA_DISC_BATCHED=1`, DESIGN.md open item 3): one input preparation and one
trunk pass over [policy | labelled | unlabelled] rows, the gradient penalty taken from the shared graph.  Host-logic check on
CPU: the 11 statistics of gail.py:415-541, the parameter gradients (incl. the double-backward term), the prior update and the
running normaliser must equal the three-pass path (itself pinned against the reference on the GPU) up to summation order."""
import types

import torch

from qa_b200.config import bbc_train_cfg
from qa_b200.rsl_rl import ActorCritic, Discriminator, Estimator, Normalizer, SSInfoGAIL


def _alg(batched, loss_fn):
    torch.manual_seed(3)
    cfg = bbc_train_cfg()
    ac = ActorCritic(101, 671, 12, 57, 10, 4, 29, 11, **cfg["policy"])
    est = Estimator(57, 4, hidden_dims=[128, 64])
    env = types.SimpleNamespace(task_obs_weight_decay=True, task_obs_weight=0.7, dim_c=5, num_obs_disc=49,
                                cfg=types.SimpleNamespace(), latent_eps=None, latent_c=None)
    disc = Discriminator(env, 98, 49, 5, 0.02, loss_fn, None, 1.0, 0.01, 0.2, 0.2, 2, 2, 0.0, [512, 256], "cpu")
    norm = Normalizer(98)
    g = torch.Generator().manual_seed(4)
    norm.mean[:] = 0.1 * torch.randn(98, generator=g).numpy()
    norm.var[:] = (0.5 + torch.rand(98, generator=g)).numpy()
    alg_cfg = dict(cfg["algorithm"], disc_replay_buffer_size=64, use_cuda_graph=False, fused_loss=False, disc_loss_function=loss_fn)
    alg = SSInfoGAIL(env, ac, disc, est, cfg["estimator"], None, norm, 2, 2, 49, 0.0, device="cpu", **alg_cfg)
    alg.disc_batched = batched
    alg._init_disc_update()
    alg._disc_optim_step = lambda *a: None                               # K8 is CUDA-only; the gradients are what is compared
    alg.info_max_coef_on = 0.3
    alg._info_max_coef_on.fill_(0.3)
    return alg, env, norm


def _batches(seed, n_pi=48, n_lb=40, n_ulb=56):
    g = torch.Generator().manual_seed(seed)
    c = torch.nn.functional.one_hot(torch.randint(0, 5, (n_pi,), generator=g), 5).float()
    return ((torch.randn(n_pi, 98, generator=g), torch.rand(n_pi, 1, generator=g) * 2 - 1, c),
            (torch.randn(n_lb, 98, generator=g), torch.randint(0, 5, (n_lb,), generator=g)),
            torch.randn(n_ulb, 98, generator=g))


def test_batched_step_equals_three_pass_step():
    for loss_fn in ("MSELoss", "BCEWithLogitsLoss"):
        res = []
        for batched in (False, True):
            alg, env, norm = _alg(batched, loss_fn)
            for step in range(2):                                        # the second step sees the updated normaliser and prior
                stats = alg.update_ss_info_gail(*_batches(10 + step))
            norm.sync_host()
            res.append((torch.stack(stats), alg.disc_flat.grad.clone(), env.prior_parameters.clone(),
                        torch.from_numpy(norm.mean.copy()), torch.from_numpy(norm.var.copy()), norm.count))
        (s0, g0, p0, m0, v0, c0), (s1, g1, p1, m1, v1, c1) = res
        assert float(g0.abs().max()) > 0
        assert torch.allclose(s1, s0, rtol=1e-5, atol=1e-6), (loss_fn, s0, s1)
        assert torch.allclose(g1, g0, rtol=1e-4, atol=1e-6 * float(g0.abs().max())), (loss_fn, float((g1 - g0).abs().max()))
        assert torch.allclose(p1, p0, rtol=1e-6, atol=1e-8)
        assert torch.allclose(m1, m0, rtol=1e-5, atol=1e-6) and torch.allclose(v1, v0, rtol=1e-5, atol=1e-6) and c0 == c1


def test_shared_priv_latent_pass_gives_the_same_ppo_step_gradients():
    """Opt-in `QA_SHARE_PRIV_LATENT=1`: the PPO minibatch step evaluates the privileged-latent encoder once for the actor and the
    regulariser (gail.py:338, :352 evaluate it twice).  Same losses, same gradients up to the accumulation order."""
    res = []
    for share in (False, True, "arena"):
        alg, env, norm = _alg(False, "MSELoss")
        alg.share_priv_latent = share is True
        if share == "arena":                                          # gradients re-homed into the single-all-reduce arena
            alg.use_grad_arena()
        alg.init_storage(64, 24, [671], [671], [12])
        alg._alloc_minibatch(384)
        if alg._grad_arena is None:
            alg._kl = torch.zeros(())
        alg._priv_reg_coef.fill_(0.07)
        g = torch.Generator().manual_seed(0)
        mb = alg._mb
        mb["obs"].copy_(0.5 * torch.randn(384, 671, generator=g))
        mb["critic_obs"].copy_(mb["obs"])
        mb["actions"].copy_(torch.randn(384, 12, generator=g))
        mb["old_mu"].copy_(0.1 * torch.randn(384, 12, generator=g))
        mb["old_sigma"].fill_(1.05)
        mb["old_actions_log_prob"].copy_(-12 + torch.randn(384, 1, generator=g))
        mb["advantages"].copy_(torch.randn(384, 1, generator=g))
        mb["returns"].copy_(torch.randn(384, 1, generator=g))
        mb["values"].copy_(torch.randn(384, 1, generator=g))
        mb["hist_latent"].copy_(0.3 * torch.randn(384, 29, generator=g))
        alg._forward_backward()
        res.append((alg.ac_flat.grad.clone(), alg.est_flat.grad.clone(), alg._ppo_stats.clone(), alg._aux_loss.clone()))
    (ga, ge, ps, ax), (gb, ge2, ps2, ax2), (gc, ge3, ps3, ax3) = res
    # the arena run is the default computation with the gradients landing in one shared buffer (autograd accumulates in place)
    assert torch.equal(gc, ga) and torch.equal(ge3, ge) and torch.equal(ps3, ps)
    n_ac = alg.ac_flat.numel
    assert torch.equal(alg._grad_arena[:n_ac], gc) and torch.equal(alg._grad_arena[n_ac:n_ac + alg.est_flat.numel], ge3)
    assert alg._kl.data_ptr() == alg._grad_arena[n_ac + alg.est_flat.numel:].data_ptr() and float(alg._kl) == float(ps3[3])
    assert float(ga.abs().max()) > 0 and torch.allclose(gb, ga, rtol=1e-5, atol=1e-7 * float(ga.abs().max()))
    assert torch.equal(ge, ge2) and torch.allclose(ps, ps2, rtol=1e-6, atol=1e-8) and torch.allclose(ax, ax2, rtol=1e-6, atol=1e-8)
    lo, n = alg.ac_flat.slices["priv_encoder.0.weight"]
    assert float(ga[lo:lo + n].abs().max()) > 0                   # the encoder does receive both gradient paths


def test_tsc_shared_priv_latent_pass_gives_the_same_ppo_step_gradients():
    """The same opt-in for the TSC `PPO` minibatch step (ppo.py:176, :186 evaluate the encoder twice)."""
    from qa_b200 import synthetic
    from qa_b200.config import tsc_train_cfg
    from qa_b200.rsl_rl import ActorCriticTSC, Estimator, PPO
    res = []
    for share in (False, True):
        torch.manual_seed(0)
        cfg = tsc_train_cfg()
        ac = ActorCriticTSC(65, 8, 132, 800, 29, 4, 10, 3, 6, device="cpu", **cfg["policy"])
        est = Estimator(input_dim=57, output_dim=4, hidden_dims=[128, 64])
        paras = dict(priv_states_dim=4, num_prop=57, num_auxiliary=8, num_scan=132, learning_rate=1e-4, train_with_estimated_states=True)
        alg = PPO(ac, None, est, paras, device="cpu", use_cuda_graph=False, fused_loss=False, **cfg["algorithm"])
        synthetic.load_student_weights(alg.actor_critic, 3)
        with torch.no_grad():
            alg.actor_critic.std.fill_(0.8)
        alg.share_priv_latent = share
        alg.init_storage(32, 8, [800], [None], [19])
        alg._alloc_minibatch(128)
        alg._priv_reg_coef.fill_(0.05)
        g = torch.Generator().manual_seed(1)
        mb = alg._mb
        mb["obs"].copy_(0.5 * torch.randn(128, 800, generator=g))
        mb["critic_obs"].copy_(mb["obs"])
        mb["actions"].copy_(torch.randn(128, 19, generator=g))
        mb["actions"][:, 0] = torch.randint(0, 3, (128,), generator=g).float()
        for k in ("old_actions_log_prob_d", "old_actions_log_prob_c", "advantages", "returns", "values"):
            mb[k].copy_(torch.randn(128, 1, generator=g))
        mb["old_mu"].copy_(0.1 * torch.randn(128, 18, generator=g))
        mb["old_sigma"].fill_(0.9)
        mb["hist_latent"].copy_(0.3 * torch.randn(128, 29, generator=g))
        alg._forward_backward()
        res.append((alg.ac_flat.grad.clone(), alg.est_flat.grad.clone()))
    (ga, ge), (gb, ge2) = res
    assert float(ga.abs().max()) > 0 and torch.allclose(gb, ga, rtol=1e-5, atol=1e-7 * float(ga.abs().max())) and torch.equal(ge, ge2)

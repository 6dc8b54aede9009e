"""TSC trainer: ActorCriticTSC / PPO behind the reference's API.

CPU: the oracle restatement and the product's fp32 forward against the golden vectors that
`oracle/gen_golden_tsc.py` produced by running the UNMODIFIED tsc/rsl_rl classes.
GPU: the loss kernel K15 against autograd on the oracle's loss graph, one PPO minibatch step (fused kernel path and
torch path) against the reference's post-update numbers, and the CUDA-graph update() against the eager one.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import tsc_trainer as OT  # noqa: E402
from helpers import GOLD, assert_close  # noqa: E402
from qa_b200 import synthetic  # noqa: E402
from qa_b200.rsl_rl import ActorCriticTSC, Estimator, PPO  # noqa: E402

POLICY = dict(scan_encoder_dims=[128, 64, 32], actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128],
              priv_encoder_dims=[64], activation="elu", init_noise_std=1.0, tanh_encoder_output=False)
EST = dict(priv_states_dim=4, num_prop=57, num_auxiliary=8, num_scan=132, learning_rate=1e-4, train_with_estimated_states=True)
NET_RTOL, NET_ATOL = 1e-4, 2e-5
ACT_KEYS = ("actions", "values", "actions_log_prob_d", "actions_log_prob_c", "action_mean", "action_sigma")


def golden():
    z = np.load(os.path.join(GOLD, "tsc_trainer_seed3.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def build(dev, **kw):
    w = synthetic.make_tsc_weights(3)
    ac = ActorCriticTSC(65, 8, 132, 800, 29, 4, 10, 3, 6, device=dev, **POLICY)
    ac.load_state_dict(w["ac"])
    est = Estimator(57, 4, hidden_dims=[128, 64])
    est.load_state_dict(w["est"])
    cfg = dict(device=dev, learning_rate=5e-4, schedule="adaptive", desired_kl=0.01, value_loss_coef=1.0, entropy_coef=0.01,
               num_learning_epochs=1, num_mini_batches=1, priv_reg_coef_schedual=[0, 0.1, 500, 1000], gamma=0.99, lam=0.95,
               use_clipped_value_loss=True, clip_param=0.2, max_grad_norm=1.0)
    cfg.update(kw)
    return PPO(ac, None, est, EST, **cfg), w


def batch_from(g, dev="cpu"):
    b = {k: g["in.batch." + k].to(dev) for k in ("advantages", "returns", "old_actions_log_prob_d", "old_actions_log_prob_c",
                                                 "old_mu", "old_sigma", "target_values", "actions")}
    b["obs"] = b["critic_obs"] = g["in.obs"].to(dev)
    return b


def test_state_dict_keys_match_reference_layout():
    alg, w = build("cpu")
    assert list(alg.actor_critic.state_dict().keys()) == [k for k, _ in synthetic.TSC_AC_SHAPES]
    for k, shape in synthetic.TSC_AC_SHAPES:
        assert tuple(alg.actor_critic.state_dict()[k].shape) == shape, k


def test_oracle_and_product_act_match_reference_golden_cpu():
    g = golden()
    alg, w = build("cpu")
    obs, draw, mode_u = g["in.obs"], g["in.draw"], g["in.mode_u"]
    for he in (False, True):
        o = OT.act(w["ac"], w["est"], obs, obs, draw, mode_u, hist_encoding=he)
        a = alg.act(obs.clone(), obs.clone(), None, hist_encoding=he, normal_draw=draw, mode_u=mode_u)
        tr = alg.transition
        got = dict(actions=a, values=tr.values, actions_log_prob_d=tr.actions_log_prob_d,
                   actions_log_prob_c=tr.actions_log_prob_c, action_mean=tr.action_mean, action_sigma=tr.action_sigma)
        for k in ACT_KEYS:
            assert_close(f"oracle act{int(he)}.{k}", o[k], g[f"act{int(he)}.{k}"], rtol=1e-6, atol=1e-6)
            assert_close(f"product act{int(he)}.{k}", got[k], g[f"act{int(he)}.{k}"], rtol=1e-5, atol=2e-6)
    assert set(g["act0.actions"][:, 0].long().tolist()) == {0, 1, 2}          # every mode is drawn in the fixture


def test_oracle_losses_match_reference_golden_cpu():
    g = golden()
    w = synthetic.make_tsc_weights(3)
    L = OT.ppo_losses(w["ac"], w["est"], batch_from(g), priv_reg_coef=OT.tsc_priv_reg_coef(int(g["in.counter"])))
    for k in ("value_loss", "surrogate_loss", "estimator_loss", "priv_reg_loss", "kl_mean", "entropy"):
        assert_close(f"ppo.{k}", L[k].detach(), g[f"ppo.{k}"], rtol=1e-5, atol=1e-7)
    assert OT.adaptive_lr(5e-4, float(L["kl_mean"])) == pytest.approx(float(g["ppo.lr_new"]), abs=1e-12)


@pytest.mark.gpu
def test_loss_kernel_matches_autograd_of_the_oracle_graph():
    """K15 forward statistics and all four gradients against torch autograd on oracle/tsc_trainer.py's loss terms."""
    from qa_b200 import ops
    dev = "cuda:0"
    gen = torch.Generator().manual_seed(0)
    M = 1000
    logits = (2.0 * torch.randn(M, 3, generator=gen)).to(dev).requires_grad_(True)
    logits.data[3] = torch.tensor([40.0, -40.0, 0.0])                    # saturated softmax: clamped-log branch
    mu = torch.randn(M, 18, generator=gen).to(dev).requires_grad_(True)
    value = torch.randn(M, 1, generator=gen).to(dev).requires_grad_(True)
    std = (0.5 + torch.rand(18, generator=gen)).to(dev).requires_grad_(True)
    actions = torch.cat([torch.randint(0, 3, (M, 1), generator=gen).float(), torch.randn(M, 18, generator=gen)], 1).to(dev)
    old_d, old_c = (-1.0 + 0.3 * torch.randn(M, generator=gen)).to(dev), (-25 + 3 * torch.randn(M, generator=gen)).to(dev)
    adv, ret, tv = (torch.randn(M, generator=gen).to(dev) for _ in range(3))
    old_mu, old_sigma = torch.randn(M, 18, generator=gen).to(dev), (0.5 + torch.rand(M, 18, generator=gen)).to(dev)
    # torch graph (same expressions as the oracle's ppo_losses)
    prob = torch.softmax(logits, -1)
    p, logit = OT.categorical(prob)
    logp_d = logit.gather(-1, actions[:, :1].long()).squeeze(-1)
    sigma = mu * 0. + std
    logp_c = OT.normal_log_prob(actions[:, 1:], mu, sigma)
    ent = (1.4189385332046727 + torch.log(sigma)).mean(-1) - (logit * p).sum(-1)

    def surr(lp, old):
        r = torch.exp(lp - old)
        return torch.max(-adv * r, -adv * torch.clamp(r, 0.8, 1.2)).mean()

    s_loss = surr(logp_d, old_d) + surr(logp_c, old_c)
    v = value.squeeze(1)
    vc = tv + (v - tv).clamp(-0.2, 0.2)
    v_loss = torch.max((v - ret).pow(2), (vc - ret).pow(2)).mean()
    kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5) + (old_sigma ** 2 + (old_mu - mu) ** 2) / (2.0 * sigma ** 2) - 0.5, -1).mean()
    loss = s_loss + 1.0 * v_loss - 0.01 * ent.mean()
    gl, gm, gv, gs = torch.autograd.grad(loss, (logits, mu, value, std))
    dl, dm = torch.empty(M, 4, device=dev)[:, :3], torch.empty(M, 20, device=dev)[:, :18]
    dv, ds, stats = torch.empty(M, device=dev), torch.empty(18, device=dev), torch.empty(4, device=dev)
    ops.ppo_loss_tsc(logits.detach(), mu.detach(), std.detach(), value.detach(), actions, old_d, old_c, adv, ret, tv, old_mu,
                     old_sigma, dl, dm, dv, ds, stats, 0.2, 1.0, 0.01, True)
    assert_close("stats", stats, torch.stack([s_loss, v_loss, ent.mean(), kl]).detach(), rtol=1e-4, atol=1e-6)
    assert_close("dlogits", dl, gl, rtol=1e-3, atol=1e-8)
    assert_close("dmu", dm, gm, rtol=1e-3, atol=1e-8)
    assert_close("dvalue", dv, gv.squeeze(1), rtol=1e-4, atol=1e-9)
    assert_close("dstd", ds, gs, rtol=1e-3, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("fused_loss", [False, True])
def test_ppo_minibatch_step_matches_reference_golden(fused_loss):
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda:0"
    g = golden()
    alg, w = build(dev, use_cuda_graph=False, fused_loss=fused_loss)
    alg.counter = int(g["in.counter"])
    alg.init_storage(64, 1, [800], [None], [19])
    alg._alloc_minibatch(64)
    mb, b = alg._mb, batch_from(g, dev)
    mb["obs"].copy_(b["obs"])
    mb["critic_obs"].copy_(b["obs"])
    for k_mb, k_b in (("actions", "actions"), ("values", "target_values"), ("returns", "returns"), ("advantages", "advantages"),
                      ("old_actions_log_prob_d", "old_actions_log_prob_d"), ("old_actions_log_prob_c", "old_actions_log_prob_c"),
                      ("old_mu", "old_mu"), ("old_sigma", "old_sigma")):
        mb[k_mb].copy_(b[k_b])
    with torch.no_grad():
        mb["hist_latent"].copy_(alg.actor_critic.actor.infer_hist_latent(mb["obs"]))
    alg._priv_reg_coef.fill_(OT.tsc_priv_reg_coef(alg.counter))
    alg._stats.zero_()
    alg._minibatch_step()
    torch.cuda.synchronize()
    s = alg._stats.cpu()
    for i, k in ((0, "surrogate_loss"), (1, "value_loss"), (2, "entropy"), (4, "priv_reg_loss"), (5, "estimator_loss")):
        assert_close(f"ppo.{k}", s[i], g[f"ppo.{k}"], rtol=NET_RTOL, atol=NET_ATOL)
    assert_close("kl_mean", s[6], g["ppo.kl_mean"], rtol=1e-3, atol=1e-5)
    assert abs(alg.learning_rate - float(g["ppo.lr_new"])) < 1e-9
    stride, lr = int(g["in.param_stride"]), float(g["ppo.lr_new"])
    for name, mod, key in (("ac", alg.actor_critic, "ppo.ac_params_sampled"), ("est", alg.estimator, "ppo.est_params_sampled")):
        flat = torch.cat([v.reshape(-1) for v in mod.state_dict().values()])[::stride].cpu()
        d = (flat - g[key]).abs()
        # Adam's first step is lr * sign-like: the few entries whose gradient is ~0 amplify summation-order noise
        assert float(d.max()) <= 2.5 * lr, (name, float(d.max()))
        assert float((d > 2e-5).float().mean()) < 5e-3, (name, float((d > 2e-5).float().mean()))


@pytest.mark.gpu
def test_update_cuda_graph_matches_eager_and_rollout_api():
    """act -> process_env_step over T steps, compute_returns (K5), update(): graph replay == eager, stats finite."""
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda:0"
    T, N = 4, 256
    results = []
    for graph in (False, True):
        alg, w = build(dev, use_cuda_graph=graph, num_learning_epochs=2, num_mini_batches=2)
        alg.init_storage(N, T, [800], [None], [19])
        gen = torch.Generator().manual_seed(5)
        for t in range(T):
            obs = (0.5 * torch.randn(N, 800, generator=gen)).to(dev)
            a = alg.act(obs, obs, None, normal_draw=torch.randn(N, 18, generator=gen).to(dev), mode_u=torch.rand(N, generator=gen).to(dev))
            assert a.shape == (N, 19) and float(a[:, 0].max()) <= 2
            rew = torch.rand(N, generator=gen).to(dev)
            dones = (torch.rand(N, generator=gen) < 0.05).to(dev)
            alg.process_env_step(rew, dones, {"time_outs": dones})
        alg.compute_returns((0.5 * torch.randn(N, 800, generator=gen)).to(dev))
        idx = torch.randperm(T * N, generator=gen).to(dev)
        out = alg.update(idx)
        assert alg.storage.step == 0 and alg.counter == 1
        results.append((out, torch.cat([v.reshape(-1) for v in alg.actor_critic.state_dict().values()]).clone()))
    for a, b in zip(results[0][0], results[1][0]):
        assert np.isfinite(a) and abs(a - b) <= 1e-4 * abs(a) + 1e-6, (results[0][0], results[1][0])
    assert_close("params graph vs eager", results[1][1], results[0][1], rtol=1e-4, atol=2e-5)

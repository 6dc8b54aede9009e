"""GPU parity of the trainer half (policy act, discriminator reward, gather, fused clip+Adam, one PPO
minibatch step) against golden vectors produced by the UNMODIFIED reference classes
(oracle/gen_golden_policy.py) and against the CPU oracle.

Tolerances: the dense layers run through cuBLAS/tensor cores in fp32 here (TF32 is NOT enabled), so the
comparison with the reference's CPU fp32 results uses rtol 1e-4 on network outputs (different summation order
over K up to 671) and 1e-5 on everything element-wise."""
import types

import numpy as np
import pytest
import torch

import trainer as OT
from helpers import GOLD, assert_close
from qa_b200 import ops, synthetic
from qa_b200.config import bbc_train_cfg
from qa_b200.rsl_rl import ActorCritic, Estimator, Discriminator, Normalizer, RolloutStorage, SSInfoGAIL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False     # parity path: full fp32 contractions
torch.backends.cudnn.allow_tf32 = False
NET_RTOL, NET_ATOL = 1e-4, 2e-5


def load_golden():
    z = np.load(f"{GOLD}/trainer_policy_seed3.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def build(w, n_envs=64, fused_loss=True):
    cfg = bbc_train_cfg()
    ac = ActorCritic(101, 671, 12, 57, 10, 4, 29, 11, **cfg["policy"])
    ac.load_state_dict(w["ac"])
    est = Estimator(57, 4, hidden_dims=[128, 64])
    est.load_state_dict(w["est"])
    env = types.SimpleNamespace(task_obs_weight_decay=True, task_obs_weight=0.7, dim_c=5, num_obs_disc=49,
                                cfg=types.SimpleNamespace(), latent_eps=None, latent_c=None)
    disc = Discriminator(env, 98, 49, 5, 0.02, "MSELoss", None, 1.0, 0.01, 0.2, 0.2, 2, 2, 0.0, [512, 256], DEV)
    disc.load_state_dict(w["disc"])
    norm = Normalizer(98)
    norm.mean, norm.var = w["norm_mean"].numpy().copy(), w["norm_var"].numpy().copy()
    alg_cfg = dict(cfg["algorithm"], disc_replay_buffer_size=1024, use_cuda_graph=False, fused_loss=fused_loss)
    alg = SSInfoGAIL(env, ac, disc, est, cfg["estimator"], None, norm, 2, 2, 49, 0.0, device=DEV, **alg_cfg)
    alg.init_storage(n_envs, 24, [671], [671], [12])
    return alg, env, norm


def test_act_and_disc_reward_match_reference_golden():
    g = load_golden()
    alg, env, norm = build(synthetic.make_weights(3))
    obs, draw = g["in.obs"].to(DEV), g["in.draw"].to(DEV)
    for he in (False, True):
        a = alg.act(obs.clone(), obs.clone(), hist_encoding=he, normal_draw=draw)
        tr = alg.transition
        for k, v in (("actions", a), ("values", tr.values), ("actions_log_prob", tr.actions_log_prob),
                     ("action_mean", tr.action_mean), ("action_sigma", tr.action_sigma)):
            assert_close(f"act{int(he)}.{k}", v, g[f"act{int(he)}.{k}"], rtol=NET_RTOL, atol=NET_ATOL)
    out = alg.disc.predict_disc_reward(g["in.rew_t"].to(DEV), obs, g["in.disc_hist"].to(DEV), normalizer=norm)
    for k, v in zip(("rewards", "reward_i", "reward_us", "reward_ss", "reward_t"), out):
        assert v.dtype == g[f"disc.{k}"].dtype, k                       # float64 quirk preserved
        assert_close(f"disc.{k}", v, g[f"disc.{k}"], rtol=NET_RTOL, atol=NET_ATOL)


@pytest.mark.parametrize("fused_loss", [False, True])
def test_ppo_minibatch_step_matches_reference_golden(fused_loss):
    g = load_golden()
    alg, env, norm = build(synthetic.make_weights(3), fused_loss=fused_loss)
    alg.priv_reg_counter = int(g["in.priv_reg_counter"])
    alg._alloc_minibatch(64)
    alg._kl = torch.zeros((), device=DEV)
    mb = alg._mb
    mb["obs"].copy_(g["in.obs"])
    mb["critic_obs"].copy_(g["in.obs"])
    for k_mb, k_g in (("actions", "actions"), ("values", "target_values"), ("returns", "returns"),
                      ("old_actions_log_prob", "old_actions_log_prob"), ("advantages", "advantages"),
                      ("old_mu", "old_mu"), ("old_sigma", "old_sigma")):
        mb[k_mb].copy_(g["in.batch." + k_g])
    with torch.no_grad():        # update() encodes the history of the whole rollout once; here: of this minibatch
        mb["hist_latent"].copy_(alg.actor_critic.infer_hist_latent(mb["obs"][:, 90:660]))
    alg._priv_reg_coef.fill_(OT.priv_reg_coef(1500))
    alg._stats.zero_()
    alg._minibatch_step()
    torch.cuda.synchronize()
    stats = alg._stats.cpu()
    for i, k in enumerate(("surrogate_loss", "value_loss", "b_loss", "entropy", "priv_reg_loss", "estimator_loss")):
        assert_close(f"ppo.{k}", stats[i], g[f"ppo.{k}"], rtol=NET_RTOL, atol=NET_ATOL)
    assert_close("kl_mean", stats[6], g["ppo.kl_mean"], rtol=1e-3, atol=1e-5)
    assert abs(alg.lr_ac - float(g["ppo.lr_new"])) < 1e-9
    stride = int(g["in.param_stride"])
    # post-step parameters: Adam's first step moves every weight by ~lr * sign(g); compare sampled entries
    ac_flat = torch.cat([v.reshape(-1) for v in alg.actor_critic.state_dict().values()])[::stride]
    est_flat = torch.cat([v.reshape(-1) for v in alg.estimator.state_dict().values()])[::stride]
    assert_close("ac params after step", ac_flat, g["ppo.ac_params_sampled"], rtol=1e-4, atol=2e-5)
    assert_close("est params after step", est_flat, g["ppo.est_params_sampled"], rtol=1e-4, atol=2e-6)


def test_gather_minibatch_matches_indexing():
    g = torch.Generator().manual_seed(0)
    R = 24 * 512
    srcs = [torch.randn(R, w, generator=g).to(DEV) for w in (671, 671, 12, 1, 1, 1, 1, 12, 12)]
    idx = torch.randperm(R, generator=g)[:3000].to(DEV)
    dsts = [torch.empty(3000, s.shape[1], device=DEV) for s in srcs]
    ops.gather_minibatch(idx, srcs, dsts)
    for s, d in zip(srcs, dsts):
        assert torch.equal(d, s[idx])
    ops.gather_minibatch(idx[:0], srcs, [d[:0] for d in dsts])         # empty minibatch


def test_gather_minibatch_padded_destination_pitch():
    g = torch.Generator().manual_seed(2)
    R = 4096
    srcs = [torch.randn(R, w, generator=g).to(DEV) for w in (671, 29, 12)]
    idx = torch.randperm(R, generator=g)[:1000].to(DEV)
    dsts = [torch.full((1000, (s.shape[1] + 3) // 4 * 4), 7.0, device=DEV)[:, :s.shape[1]] for s in srcs]
    ops.gather_minibatch(idx, srcs, dsts)
    for s, d in zip(srcs, dsts):
        assert d.stride(0) % 4 == 0
        assert torch.equal(d, s[idx])


@pytest.mark.parametrize("mode,width", [(0, 4), (1, 29), (1, 32), (0, 1)])
def test_row_loss_matches_torch_forward_and_gradient(mode, width):
    """K12 against the reference expressions gail.py:354 (mean row L2 distance) and :359 (MSE)."""
    g = torch.Generator().manual_seed(3)
    M = 3001
    a = torch.randn(M, width, generator=g).to(DEV).requires_grad_(True)
    b = torch.randn(M, width, generator=g).to(DEV)
    if mode == 1:
        with torch.no_grad():
            a[5] = b[5]                                                  # zero distance: subgradient 0, no NaN
    want = (a - b).pow(2).mean() if mode == 0 else (a - b).norm(p=2, dim=1).mean()
    (gw,) = torch.autograd.grad(want, a)
    da = torch.empty(M, (width + 3) // 4 * 4, device=DEV)[:, :width]
    loss = torch.zeros(1, device=DEV)
    ops.row_loss(a.detach(), b, da, loss, mode)
    assert_close("row loss", loss[0], want.detach(), rtol=1e-5, atol=1e-7)
    assert_close("row loss grad", da, gw, rtol=1e-5, atol=1e-9)


def test_ppo_scalars_adaptive_lr_and_running_stats():
    """K13 against the rule of gail.py:374-379."""
    std = torch.tensor([0.5, 1.0, 2.0], device=DEV)
    ent = float((1.4189385332046727 + torch.log(std)).sum())
    for kl, lr0, want_lr in ((0.05, 1e-3, 1e-3 / 1.5), (0.001, 1e-3, 1.5e-3), (0.01, 1e-3, 1e-3), (0.05, 1.2e-5, 1e-5),
                             (0.001, 9e-3, 1e-2), (0.0, 1e-3, 1e-3)):
        lr = torch.tensor([lr0], device=DEV)
        acc = torch.ones(7, device=DEV)
        ops.ppo_scalars(torch.tensor([1., 2., 3., kl], device=DEV), std, torch.tensor([4.], device=DEV),
                        torch.tensor([5.], device=DEV), torch.tensor([kl], device=DEV), 0.01, lr, acc)
        assert abs(float(lr) - want_lr) < 1e-9, (kl, lr0)
        assert_close("stats", acc, torch.tensor([2., 3., 4., 1. + ent, 5., 6., 1. + kl]), rtol=1e-6, atol=1e-7)
    lr = torch.tensor([1e-3], device=DEV)
    ops.ppo_scalars(torch.zeros(4, device=DEV), std, torch.zeros(1, device=DEV), torch.zeros(1, device=DEV),
                    torch.tensor([0.5], device=DEV), 0.0, lr, torch.zeros(7, device=DEV))     # fixed schedule
    assert float(lr) == pytest.approx(1e-3)


def test_clip_adam_matches_torch():
    g = torch.Generator().manual_seed(1)
    n = 735699
    p0 = torch.randn(n, generator=g)
    for max_norm, scale in ((1.0, 1.0), (1e9, 0.5)):
        p_ref = p0.clone().requires_grad_(True)
        opt = torch.optim.Adam([p_ref], lr=3e-4)
        p, m, v = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        lr, step = torch.full((1,), 3e-4, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
        ws, gn = torch.zeros(2, dtype=torch.float64, device=DEV), torch.zeros(1, device=DEV)
        for it in range(3):
            grad = torch.randn(n, generator=g) * (0.01 if it else 1.0)
            p_ref.grad = (grad * scale).clone()
            tn = torch.nn.utils.clip_grad_norm_([p_ref], max_norm)
            opt.step()
            ops.clip_adam(p, grad.to(DEV), m, v, lr, step, ws, max_grad_norm=max_norm, grad_scale=scale, grad_norm_out=gn)
            assert_close("grad norm", gn[0], tn, rtol=1e-5)
        assert int(step.item()) == 3
        assert_close("params", p, p_ref.detach(), rtol=1e-5, atol=1e-6)


def test_update_runs_with_and_without_cuda_graph_identically():
    """The captured-graph path of SSInfoGAIL.update is the same computation as the eager path."""
    outs = []
    for graph in (False, True):
        torch.manual_seed(0)
        alg, env, norm = build(synthetic.make_weights(4), n_envs=64)
        alg.use_cuda_graph = graph
        st = alg.storage
        g = torch.Generator().manual_seed(2)
        st.observations.copy_(torch.randn(24, 64, 671, generator=g))
        st.privileged_observations.copy_(st.observations)
        with torch.no_grad():
            for t in range(24):
                o = st.observations[t]
                alg.act(o, o, normal_draw=torch.randn(64, 12, generator=g).to(DEV))
                tr = alg.transition
                st.actions[t], st.values[t] = tr.actions, tr.values
                st.actions_log_prob[t, :, 0], st.mu[t], st.sigma[t] = tr.actions_log_prob, tr.action_mean, tr.action_sigma
        st.rewards.copy_(0.05 * torch.rand(24, 64, 1, generator=g))
        st.dones.copy_((torch.rand(24, 64, 1, generator=g) < 0.02).byte())
        st.compute_returns(torch.zeros(64, 1, device=DEV), 0.99, 0.95)
        idx = torch.randperm(24 * 64, generator=g).to(DEV)
        stats = alg.update(indices=idx)
        outs.append((stats, alg.ac_flat.data.clone(), alg.lr_ac))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(a)), (outs[0][0], outs[1][0])
    assert_close("params eager vs graph", outs[1][1], outs[0][1], rtol=1e-4, atol=1e-5)
    assert outs[0][2] == outs[1][2]


@pytest.mark.parametrize("act", [None, "elu", "relu"])
def test_act_bwd_matches_torch(act):
    """K9: gz = gy * act'(y) from the saved output, db = column sums (ragged M, N; strided views)."""
    g = torch.Generator().manual_seed(9)
    for M, N in ((24576, 512), (1000, 12), (257, 1), (4096, 128)):
        z = torch.randn(M, N + 4, generator=g).to(DEV)[:, :N]
        gy = torch.randn(M, N, generator=g).to(DEV)
        y = torch.nn.functional.elu(z) if act == "elu" else (torch.relu(z) if act == "relu" else z)
        want = gy * (torch.where(z > 0, 1.0, torch.exp(z)) if act == "elu" else ((z > 0).float() if act == "relu" else 1.0))
        gz = torch.empty(M, N, device=DEV)
        db = torch.full((N,), 7.0, device=DEV)
        ops.act_bwd(gy, y.contiguous() if act is None else y, act, gz=gz, db=db)
        assert_close("gz", gz, want, rtol=1e-5, atol=1e-6)
        assert_close("db", db, want.sum(0), rtol=1e-4, atol=1e-3)


def test_ppo_loss_kernel_matches_autograd():
    """K10 against torch autograd on the reference's formulas (gail.py:367-408), incl. samples inside and outside
    the clip range and on both sides of the value clip."""
    g = torch.Generator().manual_seed(4)
    M = 5000
    mu = (0.8 * torch.randn(M, 12, generator=g)).to(DEV).requires_grad_(True)
    value = torch.randn(M, 1, generator=g).to(DEV).requires_grad_(True)
    std = (0.5 + torch.rand(12, generator=g)).to(DEV).requires_grad_(True)
    actions = (mu.detach() + 0.7 * torch.randn(M, 12, generator=g).to(DEV))
    sigma = mu * 0. + std
    logp = (-((actions - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - 0.9189385332046727).sum(-1)
    old_logp = (logp.detach() + 0.3 * torch.randn(M, generator=g).to(DEV))
    adv = torch.randn(M, generator=g).to(DEV)
    returns = torch.randn(M, generator=g).to(DEV)
    tv = value.detach().view(-1) + 0.3 * torch.randn(M, generator=g).to(DEV)
    old_mu = mu.detach() + 0.1 * torch.randn(M, 12, generator=g).to(DEV)
    old_sigma = (sigma.detach() * 1.1).contiguous()
    clip, cs, cv, cb, ce = 0.2, 2.0, 5.0, 0.3, 0.01
    ratio = torch.exp(logp - old_logp)
    surr = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
    v = value.view(-1)
    vc = tv + (v - tv).clamp(-clip, clip)
    vl = torch.max((v - returns) ** 2, (vc - returns) ** 2).mean()
    bl = (torch.clamp(mu + 1, max=0.) ** 2 + torch.clamp(mu - 1, min=0.) ** 2).sum(-1).mean()
    ent = (1.4189385332046727 + torch.log(sigma)).sum(-1).mean()
    kl = torch.sum(torch.log(sigma / old_sigma + 1e-5) + (old_sigma ** 2 + (old_mu - mu) ** 2) / (2 * sigma ** 2) - 0.5, -1).mean()
    (cs * surr + cv * vl + cb * bl - ce * ent).backward()
    dmu, dvalue = torch.empty(M, 12, device=DEV), torch.empty(M, device=DEV)
    dstd, stats = torch.empty(12, device=DEV), torch.empty(4, device=DEV)
    ops.ppo_loss(mu.detach(), std.detach(), value.detach(), actions, old_logp, adv, returns, tv, old_mu, old_sigma, dmu,
                 dvalue, dstd, stats, clip, cs, cv, cb, ce, True)
    assert_close("stats", stats, torch.stack([surr, vl, bl, kl]).detach(), rtol=1e-4, atol=1e-6)
    assert_close("dmu", dmu, mu.grad, rtol=1e-4, atol=1e-8)
    assert_close("dvalue", dvalue, value.grad.view(-1), rtol=1e-4, atol=1e-8)
    assert_close("dstd", dstd, std.grad, rtol=1e-3, atol=1e-6)


def test_fused_history_encoder_matches_module():
    """K11 against the (reference-named) torch module it replaces, incl. a ragged tail block and the obs-row view."""
    w = synthetic.make_weights(5)
    ac = ActorCritic(101, 671, 12, 57, 10, 4, 29, 11, **bbc_train_cfg()["policy"]).to(DEV)
    ac.load_state_dict(w["ac"])
    ac.flatten_parameters()
    g = torch.Generator().manual_seed(8)
    for M in (24576, 4096, 37):
        obs = torch.randn(M, 672, generator=g).to(DEV)[:, :671]
        hist = obs[:, 90:660]
        with torch.no_grad():
            got = ac.infer_hist_latent(hist)                                    # fused kernel
            want = ac.history_encoder(hist.reshape(-1, 10, 57))                  # cuBLAS + cuDNN path
        assert got.shape == want.shape == (M, 29)
        assert_close(f"hist latent M={M}", got, want, rtol=1e-4, atol=1e-5)
    with torch.enable_grad():                                                   # gradients requested -> module path
        out = ac.infer_hist_latent(hist)
        assert out.requires_grad


def test_fused_rollout_step_matches_torch_rollout_step():
    """K18 + K19 around the discriminator GEMMs against the torch restatement of on_policy_runner.py:163-181 /
    discriminator.py:71-118 / gail.py:199-212 (which is itself pinned against the reference golden above)."""
    import bench
    from qa_b200.pipeline import BbcIteration
    torch.backends.cuda.matmul.allow_tf32 = False       # `import bench` switches TF32 on; this is the fp32 parity path
    N, T = 256, 4
    cfg, static, snaps, table = bench.build_workload(0, DEV, n_envs=N, steps=T)
    out = []
    for fused in (False, True):
        it = BbcIteration(cfg, static, snaps, table, device=DEV, seed=77, use_cuda_graph=False)
        it.runner.fused_rollout = fused
        it.env.task_obs_weight = 0.7
        torch.manual_seed(5)
        it._rollout_eager(host=False)
        st = it.runner.alg.storage
        ds = it.runner.alg.disc_storage
        out.append(dict(rewards=st.rewards.clone(), dones=st.dones.clone(), values=st.values.clone(),
                        obs=st.observations.clone(), hist=it.runner._disc_hist.clone(),
                        replay=ds.states[:T * N].clone(), replay_eps=ds.latent_eps[:T * N].clone(), n=ds.num_samples))
    a, b = out
    assert int(a["dones"].sum()) > 0 and a["n"] == b["n"] == T * N
    assert torch.equal(a["dones"], b["dones"])
    assert_close("obs", b["obs"], a["obs"])
    assert_close("values", b["values"], a["values"], rtol=NET_RTOL, atol=NET_ATOL)
    assert_close("rewards", b["rewards"], a["rewards"], rtol=NET_RTOL, atol=NET_ATOL)
    assert_close("disc history", b["hist"], a["hist"])
    assert_close("replay states", b["replay"], a["replay"])
    assert_close("replay eps", b["replay_eps"], a["replay_eps"])


def test_discriminator_update_matches_reference_golden():
    """Two consecutive `update_ss_info_gail` steps (gail.py:415-541) against the numbers the UNMODIFIED reference
    produced (oracle/gen_golden_disc.py): the 11 returned statistics, post-update parameters (three interleaved Adam
    optimisers with weight decay on the shared trunk), running-normaliser moments, prior estimate, std floor."""
    torch.backends.cuda.matmul.allow_tf32 = False
    z = np.load(f"{GOLD}/trainer_disc_seed3.npz")
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    alg, env, norm = build(synthetic.make_weights(3))
    env.task_obs_weight, env.prior_parameters = 0.8, torch.full((5,), 0.2, device=DEV)
    norm.count = 5000.0
    alg.min_std = g["in.min_std"].to(DEV)
    with torch.no_grad():
        alg.actor_critic.std.copy_(g["in.std0"].to(DEV))
    alg._init_disc_update()
    alg._info_max_coef_on.fill_(0.3)
    dev = lambda k: g[k].to(DEV)                                               # noqa: E731
    names = ("ss_loss", "info_max_loss", "disc_loss", "us_loss", "grad_pen_loss", "disc_logit_loss", "disc_weight_decay",
             "acc_lb", "acc_pi", "acc_exp", "acc_ulb")
    for step in range(2):
        out = alg.update_ss_info_gail((dev("in.pol"), dev("in.pol_eps"), dev("in.pol_c")), (dev("in.exp_lb"), dev("in.lab_lb")),
                                      dev("in.exp_ulb"))
        for k, v in zip(names, out):
            assert_close(f"s{step}.{k}", v, g[f"s{step}.{k}"], rtol=NET_RTOL, atol=NET_ATOL)
    stride = int(g["in.param_stride"])
    flat = torch.cat([v.reshape(-1) for v in alg.disc.state_dict().values()])[::stride].cpu()
    d = (flat - g["post.params_sampled"]).abs()
    assert float(d.max()) <= 2.5e-3 and float((d > 2e-5).float().mean()) < 1e-2, (float(d.max()), float((d > 2e-5).float().mean()))
    norm.sync_host()
    assert_close("normaliser mean", torch.from_numpy(norm.mean), g["post.norm_mean"], rtol=1e-5, atol=1e-8)
    assert_close("normaliser var", torch.from_numpy(norm.var), g["post.norm_var"], rtol=1e-5, atol=1e-8)
    assert abs(norm.count - float(g["post.norm_count"])) < 1e-6
    assert_close("prior", env.prior_parameters, g["post.prior"], rtol=1e-5, atol=1e-8)
    assert_close("std floor", alg.actor_critic.std.detach(), g["post.std"])


def test_update_disc_runs_over_replay_and_expert_sets():
    """`update_disc` end to end: 16 minibatch steps over a filled replay buffer and synthetic expert sets; finite
    statistics, parameters move, the normaliser count grows by 3 x minibatch x steps."""
    alg, env, norm = build(synthetic.make_weights(3), n_envs=64)
    env.task_obs_weight, env.prior_parameters = 1.0, torch.full((5,), 0.2, device=DEV)
    gen = torch.Generator().manual_seed(4)
    alg.disc_storage.insert(torch.randn(900, 98, generator=gen).to(DEV), torch.rand(900, 1, generator=gen).to(DEV),
                            torch.nn.functional.one_hot(torch.randint(0, 5, (900,), generator=gen), 5).float().to(DEV))
    expert = types.SimpleNamespace(preloaded_s_lb=torch.randn(500, 98, generator=gen).to(DEV),
                                   preloaded_label=torch.randint(0, 5, (500,), generator=gen).to(DEV),
                                   preloaded_s_ulb=torch.randn(700, 98, generator=gen).to(DEV))
    before = torch.cat([v.reshape(-1) for v in alg.disc.state_dict().values()]).clone()
    c0 = norm.count
    stats = alg.update_disc(expert, num_updates=16)
    assert len(stats) == 11 and all(np.isfinite(s) for s in stats)
    after = torch.cat([v.reshape(-1) for v in alg.disc.state_dict().values()])
    assert float((after - before).abs().max()) > 1e-5
    norm.sync_host()
    mb = 64 * 24 // (5 * 4 * 4)
    assert abs(norm.count - (c0 + 3 * mb * 16)) < 1e-6


def test_update_disc_cuda_graph_matches_eager():
    res = []
    for graph in (False, True):
        alg, env, norm = build(synthetic.make_weights(3), n_envs=64)
        alg.use_cuda_graph = graph
        env.task_obs_weight, env.prior_parameters = 0.9, torch.full((5,), 0.2, device=DEV)
        gen = torch.Generator().manual_seed(4)
        alg.disc_storage.insert(torch.randn(900, 98, generator=gen).to(DEV), torch.rand(900, 1, generator=gen).to(DEV),
                                torch.nn.functional.one_hot(torch.randint(0, 5, (900,), generator=gen), 5).float().to(DEV))
        expert = types.SimpleNamespace(preloaded_s_lb=torch.randn(500, 98, generator=gen).to(DEV),
                                       preloaded_label=torch.randint(0, 5, (500,), generator=gen).to(DEV),
                                       preloaded_s_ulb=torch.randn(700, 98, generator=gen).to(DEV))
        torch.manual_seed(11)
        stats = alg.update_disc(expert, num_updates=6)
        norm.sync_host()
        res.append((stats, torch.cat([v.reshape(-1) for v in alg.disc.state_dict().values()]).clone(), norm.mean.copy(),
                    env.prior_parameters.clone()))
    for a, b in zip(res[0][0], res[1][0]):
        assert abs(a - b) <= 1e-4 * abs(a) + 1e-6, (res[0][0], res[1][0])
    assert_close("params", res[1][1], res[0][1], rtol=1e-4, atol=2e-5)
    assert np.allclose(res[0][2], res[1][2], rtol=1e-6, atol=1e-9)
    assert_close("prior", res[1][3], res[0][3])


@pytest.mark.parametrize("fused_loss", [False, True])
def test_update_dagger_matches_reference_golden(fused_loss):
    """History-encoder adaptation (gail.py:543-575): two epochs over one 64-row minibatch against the reference's mean loss
    and post-update encoder parameters; every other actor-critic parameter must stay untouched."""
    torch.backends.cuda.matmul.allow_tf32 = False
    z = np.load(f"{GOLD}/trainer_dagger_seed3.npz")
    w = synthetic.make_weights(3)
    alg, env, norm = build(w, n_envs=64, fused_loss=fused_loss)
    alg.num_learning_epochs, alg.num_mini_batches = 2, 1
    alg.init_storage(64, 1, [671], [671], [12])
    alg.storage.observations[0].copy_(torch.from_numpy(z["obs"]).to(DEV))
    before = {k: v.clone() for k, v in alg.actor_critic.state_dict().items()}
    loss = alg.update_dagger()
    assert abs(loss - float(z["mean_loss"])) <= 1e-4 * abs(float(z["mean_loss"])) + 1e-6
    sd = alg.actor_critic.state_dict()
    enc = torch.cat([v.reshape(-1) for k, v in sd.items() if k.startswith("history_encoder.")])
    # two Adam steps at lr = 1e-4: a step is lr * sign-like, so the few entries whose gradient is ~0 amplify
    # summation-order noise up to ~lr per step (same rule as tests/test_tsc_trainer.py); everything else within 1e-4 rel
    d = (enc.cpu() - torch.from_numpy(z["encoder_params"])).abs()
    tol = 2e-6 + 1e-4 * torch.from_numpy(z["encoder_params"]).abs()
    assert float(d.max()) <= 2.5 * 1e-4, float(d.max())
    assert float((d > tol).float().mean()) < 5e-3, float((d > tol).float().mean())
    for k, v in sd.items():
        if not k.startswith("history_encoder."):
            assert torch.equal(v, before[k]), k
    assert alg.storage.step == 0

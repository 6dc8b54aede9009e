"""GPU parity of the static discriminator schedule (`qa_b200.rsl_rl.disc_plan.DiscStepPlan`, kernels K24-K30 in
csrc/qa_disc_update.cu around the tcgen05 trunk GEMMs): each kernel against the torch statement of the reference lines it
implements (the stand-ins of tests/test_disc_plan_host.py, which the host test pins against autograd on the reference-shaped
`update_ss_info_gail`), and the whole step against the numbers the UNMODIFIED reference produced
(tests/golden/trainer_disc_seed3.npz, oracle/gen_golden_disc.py; bbc/rsl_rl/algorithms/gail.py:415-541).

Tolerances: the kernels K24-K30 are fp32 (moments fp64) and are held to fp32 bars.  The trunk layers are TF32 tensor-core
contractions (operand error <= 2^-10 each, see tests/test_ppo_plan_gpu.py); the step's statistics are batch means of smooth
functions of the trunk output and are asserted at rtol 1e-2 against the fp32 golden (observed ~1e-3); the gradient penalty is
a squared norm of a product of two TF32 GEMMs (rtol 2e-2)."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLD, assert_close
from qa_b200 import ops, synthetic
from qa_b200.rsl_rl import linear
from qa_b200.rsl_rl.disc_plan import DiscStepPlan
from test_disc_plan_host import FakeDiscOps
from test_trainer_gpu import build

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ("ss_loss", "info_max_loss", "disc_loss", "us_loss", "grad_pen_loss", "disc_logit_loss", "disc_weight_decay",
         "acc_lb", "acc_pi", "acc_exp", "acc_ulb")


@pytest.fixture()
def tc_mode():
    linear.set_mode("tc")
    yield
    linear.set_mode("fp32")


def _sets(gen, n_pi=900, n_lb=500, n_ulb=700):
    replay = types.SimpleNamespace(states=torch.randn(n_pi, 98, generator=gen).to(DEV),
                                   latent_eps=(torch.rand(n_pi, 1, generator=gen) * 2 - 1).to(DEV),
                                   latent_c=F.one_hot(torch.randint(0, 5, (n_pi,), generator=gen), 5).float().to(DEV))
    expert = types.SimpleNamespace(preloaded_s_lb=torch.randn(n_lb, 98, generator=gen).to(DEV),
                                   preloaded_label=torch.randint(0, 5, (n_lb,), generator=gen).to(DEV),
                                   preloaded_s_ulb=torch.randn(n_ulb, 98, generator=gen).to(DEV))
    return replay, expert


@pytest.mark.parametrize("B,decay", [(1228, True), (37, False)])
def test_k24_prepare_matches_torch(B, decay):
    gen = torch.Generator().manual_seed(B)
    replay, expert = _sets(gen)
    idx = [torch.randint(0, n, (B,), generator=gen).to(DEV) for n in (900, 500, 700)]
    mean, std = torch.randn(98, generator=gen).to(DEV) * 0.1, (torch.rand(98, generator=gen) + 0.5).to(DEV)
    w = torch.tensor(0.6, device=DEV)
    outs = []
    for fn in (ops.disc_prepare, FakeDiscOps.disc_prepare):
        x = torch.full((3 * B, 100), 9.0, device=DEV)[:, :98]
        te, tc, tl = torch.zeros(B, device=DEV), torch.zeros(B, device=DEV, dtype=torch.int32), torch.zeros(B, device=DEV, dtype=torch.int32)
        fn(B, replay, expert, *idx, decay, w, 0.02, mean, std, 5.0, x, te, tc, tl, 49)
        outs.append((x.clone(), te, tc, tl))
    a, b = outs
    assert_close("x", a[0], b[0], rtol=1e-6, atol=1e-6)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])


def _disc(seed=3):
    alg, env, norm = build(synthetic.make_weights(seed))
    env.prior_parameters = torch.full((5,), 0.2, device=DEV)
    alg._init_disc_update()
    return alg, env, norm


@pytest.mark.parametrize("B", [1228, 50])
def test_k25_heads_losses_and_gradients_match_autograd(B):
    gen = torch.Generator().manual_seed(B + 1)
    res = []
    for fn in (ops.disc_heads_loss, FakeDiscOps.disc_heads_loss):
        alg, _, _ = _disc()
        g = torch.Generator().manual_seed(B + 1)
        h2 = torch.randn(3 * B, 256, generator=g).clamp_(min=0).to(DEV)
        te = (torch.rand(B, generator=g) * 2 - 1).to(DEV)
        tc = torch.randint(0, 5, (B,), generator=g).int().to(DEV)
        tl = torch.randint(0, 5, (B,), generator=g).int().to(DEV)
        with torch.no_grad():                                            # spread the class logits: non-trivial soft-max
            alg.disc.classifier.weight.mul_(8.0)
        alg.disc_flat.grad.zero_()
        gz2, v2 = torch.zeros(3 * B, 256, device=DEV), torch.zeros(B, 256, device=DEV)
        stats, prior = torch.zeros(11, device=DEV), torch.zeros(8, device=DEV)
        fn(B, h2, alg.disc, te, tc, tl, 1.0, 1.0, 0.5, torch.tensor(0.3, device=DEV), gz2, v2, stats, prior)
        res.append((gz2, v2, stats, prior, alg.disc_flat.grad.clone()))
    a, b = res
    assert_close("gz2", a[0], b[0], rtol=1e-4, atol=1e-8)
    assert torch.equal(a[1], b[1])
    assert_close("stats", a[2], b[2], rtol=2e-5, atol=1e-6)
    assert_close("prior", a[3][:5], b[3][:5], rtol=1e-5, atol=1e-7)
    assert_close("head / bias-2 gradients", a[4], b[4], rtol=1e-4, atol=2e-6)


def test_k27_k28_k29_k30_match_torch():
    gen = torch.Generator().manual_seed(9)
    B = 1228
    g = torch.randn(B, 100, generator=gen).to(DEV)[:, :98]
    ga, gb, sa, sb = g.clone(), g.clone(), torch.zeros(11, device=DEV), torch.zeros(11, device=DEV)
    ops.disc_gp_loss(ga, 0.7, sa)
    FakeDiscOps.disc_gp_loss(gb, 0.7, sb)
    assert_close("gp", ga, gb, rtol=1e-6, atol=0)
    assert_close("gp stat", sa, sb, rtol=1e-5)
    res = []
    for fn in (ops.disc_reg, FakeDiscOps.disc_reg):
        alg, _, _ = _disc()
        alg.disc_flat.grad.fill_(0.25)
        st = torch.zeros(11, device=DEV)
        sl = alg.disc_flat.slices
        fn(alg.disc_flat, [sl["trunk.0.weight"], sl["trunk.2.weight"], sl["linear.weight"]], 0.01, 0.002, st)
        res.append((st, alg.disc_flat.grad.clone()))
    assert_close("reg stats", res[0][0], res[1][0], rtol=2e-5)
    assert_close("reg grads", res[0][1], res[1][1], rtol=1e-6, atol=1e-9)
    x = (torch.randn(3 * B, 100, generator=gen) * 2 + 1).to(DEV)[:, :98]
    ma, mb = torch.zeros(3, 2, 98, device=DEV, dtype=torch.float64), torch.zeros(3, 2, 98, device=DEV, dtype=torch.float64)
    ops.norm_moments(x, B, 3, ma)
    FakeDiscOps.norm_moments(x, B, 3, mb)
    assert_close("moments", ma, mb, rtol=1e-12, atol=1e-14)
    out = []
    for fn in (ops.norm_merge, FakeDiscOps.norm_merge):
        mean = torch.randn(98, generator=torch.Generator().manual_seed(1)).double().to(DEV)
        var = (torch.rand(98, generator=torch.Generator().manual_seed(2)) + 0.5).double().to(DEV)
        count = torch.tensor(5000.0, device=DEV, dtype=torch.float64)
        m32, s32 = torch.zeros(98, device=DEV), torch.zeros(98, device=DEV)
        prior, pb = torch.full((5,), 0.2, device=DEV), torch.tensor([0.1, 0.2, 0.3, 0.25, 0.15, 0, 0, 0], device=DEV) * 2
        std, floor = torch.linspace(0.1, 1.2, 12, device=DEV), torch.full((12,), 0.5, device=DEV)
        fn(B, 3, 2, ma * 2, mean, var, count, m32, s32, 1e-4, prior=prior, prior_batch=pb, prior_soft_coef=0.05, std=std, min_std=floor)
        out.append((mean, var, count, m32, s32, prior, std))
    for k, (u, v) in enumerate(zip(*out)):
        assert_close(f"merge[{k}]", u, v, rtol=1e-12 if u.dtype == torch.float64 else 1e-6, atol=0)


def _golden_alg():
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
    B = g["in.pol"].shape[0]
    alg.disc_storage.insert(g["in.pol"].to(DEV), g["in.pol_eps"].to(DEV), g["in.pol_c"].to(DEV))
    expert = types.SimpleNamespace(preloaded_s_lb=g["in.exp_lb"].to(DEV), preloaded_label=g["in.lab_lb"].to(DEV),
                                   preloaded_s_ulb=g["in.exp_ulb"].to(DEV))
    return alg, env, norm, g, expert, B


def test_disc_plan_two_steps_match_reference_golden(tc_mode):
    """The scheduled step on the reference's inputs (identity index vectors): the reference's 11 statistics of both steps, the
    post-update parameters, normaliser, prior and std floor."""
    alg, env, norm, g, expert, B = _golden_alg()
    assert DiscStepPlan.supported(alg) is None
    plan = DiscStepPlan(alg, B)
    idx = torch.arange(B, device=DEV)
    for step in range(2):
        alg._disc_stats.zero_()
        plan.step(expert, idx, idx, idx)
        got = alg._disc_stats.cpu()
        for k, name in enumerate(NAMES):
            want = float(g[f"s{step}.{name}"])
            rtol = 2e-2 if name == "grad_pen_loss" else 1e-2
            atol = 2.0 / B if name.startswith("acc") else 1e-5          # an accuracy may flip on a TF32-perturbed near-tie
            assert abs(float(got[k]) - want) <= atol + rtol * abs(want), (step, name, float(got[k]), want)
    stride = int(g["in.param_stride"])
    flat = torch.cat([v.reshape(-1) for v in alg.disc.state_dict().values()])[::stride].cpu()
    d = (flat - g["post.params_sampled"]).abs()
    # Adam's first steps move every weight by ~lr regardless of the gradient's size (the trunk by lr_disc + 2 lr_q = 2.5e-3 per
    # step: three optimisers): where a TF32-perturbed pre-activation flips a ReLU mask the unit's gradient row changes by O(1)
    # and small entries change sign, so a minority of weights differ by up to 2 steps' worth = 1e-2
    assert float(d.max()) <= 1e-2 and float((d > 5e-5).float().mean()) < 0.1, (float(d.max()), float((d > 5e-5).float().mean()))
    norm.sync_host()
    assert_close("normaliser mean", torch.from_numpy(norm.mean), g["post.norm_mean"], rtol=1e-5, atol=1e-7)
    assert_close("normaliser var", torch.from_numpy(norm.var), g["post.norm_var"], rtol=1e-5, atol=1e-7)
    assert abs(norm.count - float(g["post.norm_count"])) < 1e-6
    assert_close("prior", env.prior_parameters, g["post.prior"], rtol=2e-3, atol=1e-6)
    assert_close("std floor", alg.actor_critic.std.detach(), g["post.std"])


def _loss_with_masks(alg, P, x, m1, m2, B, tgt_eps, tgt_label, info_coef):
    """gail.py:454-515 in fp64 on the prepared input `x`, with the ReLU masks given (constants, as relu'' = 0 makes them in the
    reference's double backward); P = fp64 leaf copies of the discriminator parameters."""
    h1 = (x @ P["trunk.0.weight"].t() + P["trunk.0.bias"]) * m1
    h2 = (h1 @ P["trunk.2.weight"].t() + P["trunk.2.bias"]) * m2
    d = h2 @ P["linear.weight"].t() + P["linear.bias"]
    eps = h2 @ P["encoder_eps.weight"].t() + P["encoder_eps.bias"]
    c = torch.clamp(torch.softmax(h2 @ P["classifier.weight"].t() + P["classifier.bias"], -1), 1e-20, torch.inf)
    ss = F.cross_entropy(c[B:2 * B], tgt_label.long())
    cu = c[2 * B:]
    info = torch.mean(-torch.sum(cu * torch.log(cu + 1e-20), -1))
    dl = 0.5 * (F.mse_loss(d[:B], -torch.ones_like(d[:B])) + F.mse_loss(d[2 * B:], torch.ones_like(d[2 * B:])))
    us = F.l1_loss(eps[:B].view(-1), tgt_eps.double())
    g = (((m2[2 * B:] * P["linear.weight"]) @ P["trunk.2.weight"]) * m1[2 * B:]) @ P["trunk.0.weight"]      # dD/dx, :492-500
    gp = torch.mean(torch.sum(g ** 2, -1))
    logit = torch.sum(P["linear.weight"] ** 2)
    wd = torch.sum(P["trunk.0.weight"] ** 2) + torch.sum(P["trunk.2.weight"] ** 2) + logit
    return (alg.ss_coef * ss + info_coef * info + alg.disc_coef * dl + alg.us_coef * us + alg.disc_grad_penalty * gp +
            alg.disc_logit_reg * logit + alg.disc_weight_decay * wd)


@pytest.mark.parametrize("B", [96, 1228])
def test_disc_plan_gradients_match_fp64_autograd_on_the_same_masks(tc_mode, B):
    """Gradient buffer of the scheduled step (hand-derived backward incl. the gradient penalty's double backward, on TF32 GEMMs)
    against fp64 autograd.  relu' is discontinuous, so a pre-activation within TF32 rounding of zero would make an O(1)
    per-unit difference that says nothing about the kernels: the fp64 statement takes the activation MASKS from the step's own
    buffers; everything else (input preparation aside, K24 is tested above) is recomputed."""
    alg, env, norm, g, expert, _ = _golden_alg()
    alg._disc_optim_step = lambda scale=1.0: None
    gen = torch.Generator().manual_seed(B)
    replay, expert = _sets(gen)
    alg.disc_storage.insert(replay.states, replay.latent_eps, replay.latent_c)
    idx = [torch.randint(0, n, (B,), generator=gen).to(DEV) for n in (900, 500, 700)]
    plan = DiscStepPlan(alg, B)
    plan.step(expert, *idx)
    P = {n: p.detach().double().requires_grad_(True) for n, p in alg.disc.named_parameters()}
    loss = _loss_with_masks(alg, P, plan.x.double(), (plan.h1 > 0).double(), (plan.h2 > 0).double(), B, plan.tgt_eps, plan.tgt_label, 0.3)
    loss.backward()
    for name, p in alg.disc.named_parameters():
        want = P[name].grad
        err = (p.grad.double() - want)
        rel_f = float(err.norm() / want.norm().clamp(min=1e-12))
        rel_max = float(err.abs().max() / want.abs().max().clamp(min=1e-12))
        assert rel_f < 3e-3 and rel_max < 5e-3, (name, rel_f, rel_max)


def _run_update_disc(graph, iters, num_updates):
    alg, env, norm = build(synthetic.make_weights(3), n_envs=64)
    alg.use_cuda_graph = graph
    env.task_obs_weight, env.prior_parameters = 1.0, torch.full((5,), 0.2, device=DEV)
    gen = torch.Generator().manual_seed(4)
    replay, expert = _sets(gen)
    alg.disc_storage.insert(replay.states, replay.latent_eps, replay.latent_c)
    torch.manual_seed(11)
    out = []
    for it in range(iters):
        env.task_obs_weight = 1.0 - 0.3 * it
        out.append(alg.update_disc(expert, num_updates=num_updates))
    assert alg._disc_plan is not None
    norm.sync_host()
    return (out, torch.cat([v.reshape(-1) for v in alg.disc.state_dict().values()]).clone(), norm.mean.copy(), env.prior_parameters.clone())


def test_update_disc_plan_graph_matches_plan_eager(tc_mode):
    """`update_disc` over the schedule: captured graph == eager launches (same kernels; the fp32 atomic / split-K accumulation
    order differs between a graph's parallel branches and stream-ordered launches, and Adam's g / (|g| + eps) turns a last-bit
    difference of a near-zero gradient element into a fraction of lr -- measured: <= 2e-5 on ~2 of 183815 weights after one
    step, growing ~2x per step on 19-row minibatches; eager-vs-eager shows the same growth from its own 1e-7).  One step is
    held tight; two iterations of six steps are held on the statistics, and show that the decayed task_obs_weight reaches the
    captured step through its device scalar (an undecayed weight moves disc_loss by > 1e-2 here)."""
    a, b = _run_update_disc(False, 1, 1), _run_update_disc(True, 1, 1)
    d = (a[1] - b[1]).abs()
    assert float(d.max()) < 1e-4 and float((d > 1e-6).float().mean()) < 1e-3, (float(d.max()), float((d > 1e-6).float().mean()))
    np.testing.assert_allclose(a[0][0], b[0][0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-9, atol=1e-12)
    assert_close("prior", a[3], b[3], rtol=1e-5, atol=1e-8)
    a, b = _run_update_disc(False, 2, 6), _run_update_disc(True, 2, 6)
    for sa, sb in zip(a[0], b[0]):
        np.testing.assert_allclose(sa[:7], sb[:7], rtol=2e-3, atol=1e-5)
        np.testing.assert_allclose(sa[7:], sb[7:], atol=2.5 / 114)
    assert float((a[1] - b[1]).abs().max()) < 1e-2
    np.testing.assert_allclose(a[2], b[2], rtol=1e-6, atol=1e-8)
    assert_close("prior", a[3], b[3], rtol=1e-3, atol=1e-6)


def test_k8c_adam_chain_equals_the_five_sequential_adam_steps():
    """The discriminator's three optimisers (five parameter groups, the shared trunk stepped three times with separate
    moments, weight decay 1e-3) as ONE launch: parameters, moments and step counters bit-equal to five K8 calls, three steps."""
    import os
    res = []
    for chain in ("0", "1"):
        os.environ["QA_ADAM_CHAIN"] = chain
        try:
            alg, _, _ = _disc()
            g = torch.Generator().manual_seed(8)
            for _ in range(3):
                alg.disc_flat.grad.copy_(torch.randn(alg.disc_flat.grad.shape, generator=g) * 0.05)
                alg._disc_optim_step(0.5)
            torch.cuda.synchronize()
            opts = alg.optim_d + alg.optim_q_eps + alg.optim_q_c
            res.append((alg.disc_flat.data.clone(), [o.exp_avg.clone() for o in opts], [o.exp_avg_sq.clone() for o in opts],
                        [int(o.step_count) for o in opts]))
        finally:
            os.environ.pop("QA_ADAM_CHAIN", None)
    a, b = res
    assert a[3] == b[3] == [3] * 5
    assert torch.equal(a[0], b[0])
    for x, y in zip(a[1] + a[2], b[1] + b[2]):
        assert torch.equal(x, y)

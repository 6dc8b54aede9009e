"""Host check of the static discriminator schedule's LOGIC (qa_b200/rsl_rl/disc_plan.py): the hand-derived backward -- including
the double backward of the gradient penalty through the ReLU masks -- against torch autograd on the reference-shaped
`SSInfoGAIL.update_ss_info_gail` (itself pinned against the reference golden on the GPU, tests/test_trainer_gpu.py).  The
libqa_b200 entry points are replaced by torch stand-ins with the documented semantics (include/qa_b200.h); the kernels themselves
are tested on the device (tests/test_disc_plan_gpu.py)."""
import types

import pytest
import torch
import torch.nn.functional as F

from qa_b200 import ops
from qa_b200.rsl_rl import disc_plan
from test_ppo_plan_host import FakeOps, _build


class FakeDiscOps:
    @staticmethod
    def disc_prepare(B, replay, expert, idx_pi, idx_lb, idx_ulb, decay, w, step, mean, std, clip, x, tgt_eps, tgt_c, tgt_label, obs_dim):
        rows = torch.cat([replay.states[idx_pi], expert.preloaded_s_lb[idx_lb], expert.preloaded_s_ulb[idx_ulb]], 0)
        v = rows.view(3 * B, -1, obs_dim).clone()
        if decay:
            v[:, :, 3:9] *= w
            v[:, :, 33:] *= w
        L = v.shape[1]
        v = v * (torch.arange(L, dtype=torch.float32, device=v.device) * step + 1).view(1, L, 1)
        x.copy_(torch.clamp((v.reshape(3 * B, -1) - mean) / std, -clip, clip))
        tgt_eps.copy_(replay.latent_eps[idx_pi].view(-1))
        tgt_c.copy_(torch.argmax(replay.latent_c[idx_pi], -1).int())
        tgt_label.copy_(expert.preloaded_label[idx_lb].int())

    @staticmethod
    def disc_heads_loss(B, h2, disc, tgt_eps, tgt_c, tgt_label, ss_coef, disc_coef, us_coef, info_coef, gz2, v2, stats, prior_batch):
        with torch.enable_grad():
            h = h2.detach().clone().requires_grad_(True)
            P = {n: getattr(disc, n) for n in ("linear", "encoder_eps", "classifier")}
            w = {n: (m.weight.detach().clone().requires_grad_(True), m.bias.detach().clone().requires_grad_(True)) for n, m in P.items()}
            d = F.linear(h, *w["linear"])
            eps = F.linear(h, *w["encoder_eps"])
            c = torch.clamp(torch.softmax(F.linear(h, *w["classifier"]), -1), 1e-20, torch.inf)
            ss = F.cross_entropy(c[B:2 * B], tgt_label.long())
            cu = c[2 * B:]
            info = torch.mean(-torch.sum(cu * torch.log(cu + 1e-20), -1))
            dl = 0.5 * (F.mse_loss(d[:B], -torch.ones_like(d[:B])) + F.mse_loss(d[2 * B:], torch.ones_like(d[2 * B:])))
            us = F.l1_loss(eps[:B].view(-1), tgt_eps)
            (ss_coef * ss + float(info_coef) * info + disc_coef * dl + us_coef * us).backward()
        g = h.grad * (h2 > 0)
        gz2.copy_(g)
        v2.copy_((h2[2 * B:] > 0) * disc.linear.weight.detach())
        for n, m in P.items():
            m.weight.grad += w[n][0].grad
            m.bias.grad += w[n][1].grad
        disc.trunk[2].bias.grad += g.sum(0)
        with torch.no_grad():
            stats[0] += ss
            stats[1] += info
            stats[2] += dl
            stats[3] += us
            stats[7] += (torch.argmax(c[B:2 * B], -1) == tgt_label).float().mean()
            stats[8] += (d[:B] < 0).float().mean()
            stats[9] += (d[2 * B:] > 0).float().mean()
            stats[10] += (torch.argmax(c[:B], -1) == tgt_c).float().mean()
            prior_batch[:5] += cu.mean(0)

    @staticmethod
    def disc_gp_loss(g, coef, stats):
        stats[4] += (g ** 2).sum(-1).mean()
        g.mul_(2 * coef / g.shape[0])

    @staticmethod
    def disc_reg(flat, segments, c_logit, c_wd, stats):
        for k, (o, n) in enumerate(segments):
            w = flat.data[o:o + n]
            stats[6] += (w ** 2).sum()
            if k == 2:
                stats[5] += (w ** 2).sum()
            flat.grad[o:o + n] += 2 * (c_wd + (c_logit if k == 2 else 0.0)) * w

    @staticmethod
    def norm_moments(x, B, nb, moments):
        for b in range(nb):
            xb = x[b * B:(b + 1) * B].double()
            moments[b, 0] = xb.mean(0)
            moments[b, 1] = (xb ** 2).mean(0)

    @staticmethod
    def norm_merge(B, nb, world, moments, mean, var, count, mean32, std32, eps, prior=None, prior_batch=None, prior_soft_coef=0.0,
                   std=None, min_std=None):
        for b in range(nb):
            bm, ex2 = moments[b, 0] / world, moments[b, 1] / world
            bv, bc = ex2 - bm ** 2, float(B * world)
            delta, tot = bm - mean, count + bc
            m2 = var * count + bv * bc + delta ** 2 * count * bc / tot
            mean += delta * bc / tot
            var.copy_(m2 / tot)
            count.copy_(tot)
        mean32.copy_(mean.float())
        std32.copy_(torch.sqrt((var + eps).float()))
        if prior is not None:
            prior.mul_(1 - prior_soft_coef).add_(prior_batch[:5] / world * prior_soft_coef)
        if std is not None:
            std.copy_(torch.maximum(std, min_std))


@pytest.fixture()
def fake_ops(monkeypatch):
    for name in ("zero_", "linear_fwd", "linear_bwd", "act_bwd"):
        monkeypatch.setattr(ops, name, getattr(FakeOps, name))
    for name in ("disc_prepare", "disc_heads_loss", "disc_gp_loss", "disc_reg", "norm_moments", "norm_merge"):
        monkeypatch.setattr(ops, name, getattr(FakeDiscOps, name))
    yield


def _setup(seed=3):
    alg = _build(seed, M=96)
    alg.env.prior_parameters = torch.full((5,), 0.2)
    alg.env.task_obs_weight = 0.6
    g = torch.Generator().manual_seed(4)
    alg.disc_storage.insert(torch.randn(60, 98, generator=g), torch.rand(60, 1, generator=g) * 2 - 1,
                            F.one_hot(torch.randint(0, 5, (60,), generator=g), 5).float())
    expert = types.SimpleNamespace(preloaded_s_lb=torch.randn(50, 98, generator=g), preloaded_label=torch.randint(0, 5, (50,), generator=g),
                                   preloaded_s_ulb=torch.randn(70, 98, generator=g))
    alg._init_disc_update()
    alg._info_max_coef_on.fill_(0.3)
    alg.min_std = torch.full((12,), 1.5)                    # above the initial std of 1: the floor must bite
    alg._disc_optim_step = lambda scale=1.0: None          # K8 is a device kernel; gradients are what is compared here
    B = 24
    idx = [torch.randint(0, n, (B,), generator=g) for n in (60, 50, 70)]
    return alg, expert, idx, B


def test_disc_schedule_gradients_statistics_and_state_equal_autograd(fake_ops):
    ref, expert, idx, B = _setup()
    ds = ref.disc_storage
    out = ref.update_ss_info_gail((ds.states[idx[0]], ds.latent_eps[idx[0]], ds.latent_c[idx[0]]),
                                  (expert.preloaded_s_lb[idx[1]], expert.preloaded_label[idx[1]]), expert.preloaded_s_ulb[idx[2]])
    want_stats = torch.stack([o.detach().float() for o in out])
    alg, expert, idx, B = _setup()
    plan = disc_plan.DiscStepPlan(alg, B)
    alg.disc_flat.grad.fill_(7.0)                           # the schedule zeroes the gradient buffer itself
    alg._disc_stats.zero_()
    with torch.no_grad():
        plan.step(expert, *idx)
    assert torch.allclose(alg._disc_stats, want_stats, rtol=2e-5, atol=1e-6), (alg._disc_stats, want_stats)
    for (name, p), (_, q) in zip(alg.disc.named_parameters(), ref.disc.named_parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=2e-4, atol=2e-7), f"{name}: max |d| {float((p.grad - q.grad).abs().max()):.3e}"
    assert torch.allclose(alg.env.prior_parameters, ref.env.prior_parameters, rtol=1e-6, atol=1e-8)
    assert torch.equal(alg.actor_critic.std.data, torch.full((12,), 1.5)) and torch.equal(ref.actor_critic.std.data, torch.full((12,), 1.5))
    a, b = alg.disc_normalizer._device_state("cpu"), ref.disc_normalizer._device_state("cpu")
    for k in (1, 2, 3):                 # the autograd path takes the batch moments in fp32 (like numpy on a float32 batch), K29 in fp64
        assert torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-7), k
    assert torch.allclose(a[4], b[4], rtol=1e-5, atol=1e-7) and torch.allclose(a[5], b[5], rtol=1e-5)

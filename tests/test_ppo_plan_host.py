"""Host check of the static PPO schedule's LOGIC (qa_b200/rsl_rl/ppo_plan.py): which buffer feeds which launch, the column
windows, the hand-written backward order and where every gradient lands.  The libqa_b200 entry points are replaced by torch
stand-ins with the same signatures and the documented semantics (include/qa_b200.h), so the schedule runs on CPU and its flat
gradients / statistics are compared with the autograd path (`SSInfoGAIL._forward_backward`, itself pinned against the
reference golden on the GPU).  The kernels themselves are tested on the device (tests/test_ppo_plan_gpu.py)."""
import types

import pytest
import torch
import torch.nn.functional as F

from qa_b200 import ops, synthetic
from qa_b200.config import bbc_train_cfg
from qa_b200.rsl_rl import ActorCritic, Discriminator, Estimator, Normalizer, SSInfoGAIL
from qa_b200.rsl_rl import ppo_plan


def _act(y, act):
    return F.elu(y) if act == "elu" else F.relu(y) if act == "relu" else y


def _dact(y, act):
    if act == "elu":
        return torch.where(y > 0, torch.ones_like(y), y + 1)
    if act == "relu":
        return (y > 0).to(y.dtype)
    return torch.ones_like(y)


class FakeOps:
    """torch stand-ins for the entry points the schedule calls (semantics of include/qa_b200.h)."""

    @staticmethod
    def zero_(t):
        t.zero_()

    @staticmethod
    def copy_(dst, src):
        dst.copy_(src.reshape(dst.shape))

    @staticmethod
    def linear_fwd(x, w, b, y, act, x_col0=0, y_col0=0):
        N, K = w.shape
        z = x[:, x_col0:x_col0 + K] @ w.t()
        y[:, y_col0:y_col0 + N] = _act(z if b is None else z + b, act)

    @staticmethod
    def head_fwd(h, w, b, y):
        y.copy_(h @ w.t() + b)

    @staticmethod
    def head_bwd(gz, h, w, act, gz_prev=None, dw=None, db=None, db_prev=None, gz_scale=1.0):
        g = (gz.unsqueeze(1) if gz.dim() == 1 else gz) * gz_scale
        gp = (g @ w) * _dact(h, act)
        if gz_prev is not None:
            gz_prev.copy_(gp)
        if dw is not None:
            dw += g.t() @ h
        if db is not None:
            db += g.sum(0)
        if db_prev is not None:
            db_prev += gp.sum(0)

    @staticmethod
    def linear_bwd(gz, x, w, dx=None, dw=None, act_prev=None, y_prev=None, db_prev=None, db_accumulate=False, x_col0=0,
                   w_col0=0, K=None):
        if K is None:
            K = dx.shape[1] if dx is not None else dw.shape[1]
        if dx is not None:
            d = gz @ w[:, w_col0:w_col0 + K]
            if act_prev is not None:
                d = d * _dact(y_prev, act_prev)
                if db_prev is not None:
                    if not db_accumulate:
                        db_prev.zero_()
                    db_prev += d.sum(0)
            dx.copy_(d)
        if dw is not None:
            dw += gz.t() @ x[:, x_col0:x_col0 + K]

    @staticmethod
    def act_bwd(gy, y, act, gz=None, db=None, zero_db=True, addend=None, addend_scale=None):
        g = gy if addend is None else gy + (1.0 if addend_scale is None else addend_scale) * addend
        g = g * _dact(y, act) if act is not None else g
        if gz is not None:
            gz.copy_(g)
        if db is not None:
            if zero_db:
                db.zero_()
            db += g.sum(0)

    @staticmethod
    def row_loss(a, b, da, loss, mode):
        with torch.enable_grad():
            a = a.detach().clone().requires_grad_(True)
            val = (a - b).pow(2).mean() if mode == 0 else (a - b).norm(p=2, dim=1).mean()
            val.backward()
        da.copy_(a.grad)
        loss.copy_(val.detach().reshape(1))

    @staticmethod
    def ppo_loss(mu, std, value, actions, old_logp, adv, returns, target_values, old_mu, old_sigma, dmu, dvalue, dstd, stats,
                 clip, c_surr, c_value, c_bound, c_entropy, clipped_value):
        with torch.enable_grad():
            mu_, std_, v_ = (t.detach().clone().requires_grad_(True) for t in (mu, std, value))
            sg = std_.expand_as(mu_)
            logp = (-((actions - mu_) ** 2) / (2 * sg ** 2) - torch.log(sg) - 0.9189385332046727).sum(-1)
            ratio = torch.exp(logp - old_logp)
            surr = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
            v = v_.reshape(-1)
            if clipped_value:
                vc = target_values + (v - target_values).clamp(-clip, clip)
                vl = torch.max((v - returns).pow(2), (vc - returns).pow(2)).mean()
            else:
                vl = (returns - v).pow(2).mean()
            bl = (torch.clamp(mu_ + 1.0, max=0.) ** 2 + torch.clamp(mu_ - 1.0, min=0.) ** 2).sum(-1).mean()
            ent = (1.4189385332046727 + torch.log(sg)).sum(-1).mean()
            (c_surr * surr + c_value * vl + c_bound * bl - c_entropy * ent).backward()
        with torch.no_grad():
            kl = torch.sum(torch.log(sg / old_sigma + 1.e-5) + (old_sigma ** 2 + (old_mu - mu_) ** 2) / (2.0 * sg ** 2) - 0.5, -1).mean()
        dmu.copy_(mu_.grad)
        dvalue.copy_(v_.grad.reshape(-1))
        dstd.copy_(std_.grad)
        stats.copy_(torch.stack([surr.detach(), vl.detach(), bl.detach(), kl]))


@pytest.fixture()
def fake_ops(monkeypatch):
    for name in ("zero_", "copy_", "linear_fwd", "head_fwd", "head_bwd", "linear_bwd", "act_bwd", "row_loss", "ppo_loss"):
        monkeypatch.setattr(ops, name, getattr(FakeOps, name))
    yield


def _build(seed=3, M=96):
    torch.manual_seed(seed)
    cfg = bbc_train_cfg()
    w = synthetic.make_weights(seed)
    ac = ActorCritic(101, 671, 12, 57, 10, 4, 29, 11, **cfg["policy"])
    ac.load_state_dict(w["ac"])
    est = Estimator(57, 4, hidden_dims=[128, 64])
    est.load_state_dict(w["est"])
    env = types.SimpleNamespace(task_obs_weight_decay=True, task_obs_weight=0.7, dim_c=5, num_obs_disc=49, cfg=types.SimpleNamespace(),
                                latent_eps=None, latent_c=None)
    disc = Discriminator(env, 98, 49, 5, 0.02, "MSELoss", None, 1.0, 0.01, 0.2, 0.2, 2, 2, 0.0, [512, 256], "cpu")
    alg_cfg = dict(cfg["algorithm"], disc_replay_buffer_size=64, use_cuda_graph=False, fused_loss=False)
    alg = SSInfoGAIL(env, ac, disc, est, cfg["estimator"], None, Normalizer(98), 2, 2, 49, 0.0, device="cpu", **alg_cfg)
    alg.init_storage(M // 24 if M >= 24 else 1, 24, [671], [671], [12])
    return alg


def _minibatch(M, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)                                     # noqa: E731
    obs = 0.6 * r(M, 671)
    return (obs, obs + 0.05 * r(M, 671), r(M, 12), r(M, 1), r(M, 1), r(M, 1), -12 + r(M, 1), 0.3 * r(M, 12),
            0.8 + 0.2 * torch.rand(M, 12, generator=g), (None, None), None)


def test_schedule_gradients_equal_autograd(fake_ops):
    M = 96
    sample = _minibatch(M)
    ref = _build(M=M)
    ref._alloc_minibatch(M)
    ref._kl = torch.zeros(())
    for k, v in zip(("obs", "critic_obs", "actions", "values", "advantages", "returns", "old_actions_log_prob", "old_mu", "old_sigma"),
                    sample[:9]):
        ref._mb[k].copy_(v.reshape(ref._mb[k].shape))
    with torch.no_grad():
        hl = ref.actor_critic.infer_hist_latent(ref._mb["obs"][:, 90:660])
    ref._mb["hist_latent"].copy_(hl)
    ref._priv_reg_coef.fill_(0.07)
    ref._forward_backward()

    alg = _build(M=M)
    alg._kl = torch.zeros(())
    assert ppo_plan.PpoStepPlan.supported(alg) == "not a CUDA device"                  # the product path refuses the host
    plan = ppo_plan.PpoStepPlan(alg, M, 2)
    plan.load(1, sample)
    plan.sets[1]["hist_latent"].copy_(hl)
    alg._priv_reg_coef.fill_(0.07)
    # poison the gradient buffers: the schedule must zero them itself
    alg.ac_flat.grad.fill_(3.0)
    alg.est_flat.grad.fill_(3.0)
    with torch.no_grad():
        plan.forward_backward(1)
    for (name, p), (_, q) in zip(list(alg.actor_critic.named_parameters()) + list(alg.estimator.named_parameters()),
                                 list(ref.actor_critic.named_parameters()) + list(ref.estimator.named_parameters())):
        assert torch.allclose(p.grad, q.grad, rtol=2e-4, atol=2e-6), f"{name}: max |d| {float((p.grad - q.grad).abs().max()):.3e}"
    assert torch.allclose(alg._ppo_stats, ref._ppo_stats, rtol=1e-5, atol=1e-6)
    assert torch.allclose(alg._aux_loss, ref._aux_loss, rtol=1e-5, atol=1e-6)
    assert torch.allclose(alg._kl, ref._kl, rtol=1e-5, atol=1e-7)
    # padding columns of the flat gradient buffers stay zero (K8 must not move padding weights)
    assert float(alg.ac_flat.grad.abs().sum()) == pytest.approx(float(sum(p.grad.abs().sum() for p in alg.actor_critic.parameters())), rel=1e-6)


def test_schedule_gather_windows_layout(monkeypatch):
    """`plan.gather` asks K6 for exactly the windows the schedule later reads (stand-in gather on the host)."""
    def fake_gather(idx, entries):
        for s, sc, d, dc, w in entries:
            d[:, dc:dc + w] = s[idx][:, sc:sc + w]
    monkeypatch.setattr(ops, "gather_minibatch_windows", fake_gather)
    alg = _build(M=96)
    st = alg.storage
    g = torch.Generator().manual_seed(1)
    st.observations.copy_(torch.randn(st.observations.shape, generator=g))
    st.privileged_observations.copy_(torch.randn(st.observations.shape, generator=g))
    st.actions.copy_(torch.randn(st.actions.shape, generator=g))
    st.advantages.copy_(torch.randn(st.advantages.shape, generator=g))
    plan = ppo_plan.PpoStepPlan(alg, 48, 2)
    idx = torch.randperm(96, generator=g)[:48]
    lat = torch.randn(96, 29, generator=g)
    plan.gather(1, idx, lat)
    s, v = plan.sets[1], st.flat_views()
    assert torch.equal(s["obs"], v["obs"][idx]) and torch.equal(s["critic_obs"], v["critic_obs"][idx])
    assert torch.equal(s["xa"][:, :61], v["obs"][idx][:, :61]) and torch.equal(s["xa"][:, 90:], v["obs"][idx][:, 660:])
    assert torch.equal(s["hist_latent"], lat[idx]) and torch.equal(s["actions"], v["actions"][idx])
    assert torch.equal(s["advantages"], v["advantages"][idx])
    assert torch.equal(s["lat_in"], v["obs"][idx][:, 61:90])
    assert all(t.stride(0) % 4 == 0 for t in (s["obs"], s["critic_obs"], s["xa"], s["hist_latent"], s["lat_in"]))

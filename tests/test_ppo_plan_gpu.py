"""GPU parity of the path `bench.py` times: the PPO minibatch step as a static schedule of libqa_b200 launches
(`qa_b200.rsl_rl.ppo_plan.PpoStepPlan`: tcgen05 TF32 GEMMs with column windows + fused activation backward, fp32 head kernels,
fused losses) inside a CUDA graph -- against the CPU oracle (`oracle/trainer.py`, pinned on the reference's
`update_actor_critic`, bbc/rsl_rl/algorithms/gail.py:328-413) and the reference golden.

Tolerance of the tensor-core layers, derived: `tcgen05.mma.kind::tf32` reads fp32 operands with a 10-bit mantissa (the low 13
bits are dropped), so each operand carries a relative error <= 2^-10 and each product <= 2^-9 (first order); a dot product
sum_k x_k w_k is therefore within 2^-9 * sum_k |x_k||w_k| of the exact one, accumulation is fp32.  Tests of a single layer
assert exactly that bound (`scaled error` = |y - ref| / (|x| @ |w|^T + |b|) <= 2e-3 ~ 2^-9).  Through a depth-4 trunk of
1-Lipschitz activations the bound compounds to <= 4 * 2^-9 ~ 8e-3 relative to the layer-wise absolute-value products; the loss
statistics are means over the minibatch of smooth functions of the outputs, asserted at rtol 1e-2 (observed ~1e-3).  Everything that
is NOT a tensor-core contraction (head layers, losses, gathers, Adam) is full fp32 and held to the fp32 bars of the other tests."""
import numpy as np
import pytest
import torch

import trainer as OT
from helpers import GOLD, assert_close
from qa_b200 import ops, synthetic
from qa_b200.rsl_rl import linear

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TF32_BOUND = 2e-3          # 2^-9 = 1.95e-3, see the module docstring
TC_STAT_RTOL = 1e-2


@pytest.fixture()
def tc_mode():
    prev = linear.get_mode()
    linear.set_mode("tc")
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    linear.set_mode(prev)
    torch.backends.cuda.matmul.allow_tf32 = tf32


def _padded(rows, cols, fill=None, gen=None):
    t = torch.zeros(rows, (cols + 3) // 4 * 4, device=DEV)
    if gen is not None:
        t.copy_(torch.randn(t.shape, generator=gen))
    if fill is not None:
        t.fill_(fill)
    return t[:, :cols]


def _scaled_err(y, ref, x, w, b=None):
    scale = x.double().abs() @ w.double().abs().t()
    if b is not None:
        scale = scale + b.double().abs()
    return float(((y.double() - ref) .abs() / scale.clamp(min=1e-6)).max())


# ---- K20 / K21 ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,Kh,act", [(24576, 12, 128, "elu"), (4096, 1, 128, "elu"), (3001, 4, 64, "elu"), (515, 16, 32, "relu"),
                                        (7, 5, 128, None)])
def test_head_fwd_bwd_match_torch(M, N, Kh, act):
    g = torch.Generator().manual_seed(M + N)
    h = _padded(M, Kh, gen=g)
    if act == "relu":
        h = h.clamp_(min=0)
    w = (torch.randn(N, Kh, generator=g) / Kh ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    y = _padded(M, N)
    ops.head_fwd(h, w, b, y)
    ref = h.double() @ w.double().t() + b.double()
    assert_close("head_fwd", y, ref.float(), rtol=1e-5, atol=1e-5)
    gz = torch.randn(M, N, generator=g).to(DEV)
    gz_prev = _padded(M, Kh)
    dw, db, dbp = torch.ones(N, Kh, device=DEV), torch.ones(N, device=DEV), torch.ones(Kh, device=DEV)     # accumulate onto 1
    ops.head_bwd(gz, h, w, act, gz_prev=gz_prev, dw=dw, db=db, db_prev=dbp, gz_scale=0.5)
    gzd = 0.5 * gz.double()
    dh = gzd @ w.double()
    if act == "elu":
        dh = torch.where(h.double() > 0, dh, dh * (h.double() + 1))
    elif act == "relu":
        dh = torch.where(h.double() > 0, dh, torch.zeros_like(dh))
    tol = dict(rtol=2e-5, atol=2e-5 * max(1.0, M ** 0.5 / 16))
    assert_close("gz_prev", gz_prev, dh.float(), rtol=1e-5, atol=1e-5)
    assert_close("dw", dw, (1 + gzd.t() @ h.double()).float(), **tol)
    assert_close("db", db, (1 + gzd.sum(0)).float(), **tol)
    assert_close("db_prev", dbp, (1 + dh.sum(0)).float(), **tol)


# ---- K7 with column windows / fused activation backward -----------------------------------------------------------------------
def test_linear_fwd_reads_and_writes_column_windows():
    """Input window at a 16-byte aligned column of a wide row (TMA coordinates); output window at an ODD column (lanes 61..89 of
    the 101-wide actor input row: plain-store epilogue), neighbouring lanes untouched; unaligned INPUT windows are refused."""
    g = torch.Generator().manual_seed(1)
    M = 5000
    obs = _padded(M, 671, gen=g)
    w1, b1 = _padded(64, 30, gen=g), torch.randn(64, generator=g).to(DEV)
    w1.mul_(1 / 30 ** 0.5)
    h = _padded(M, 64)
    ops.linear_fwd(obs, w1, b1, h, "elu", x_col0=60)
    x = obs[:, 60:90]
    ref = torch.nn.functional.elu(x.double() @ w1.double().t() + b1.double())
    assert _scaled_err(h, ref, x, w1, b1) < TF32_BOUND
    with pytest.raises(RuntimeError, match="QA_EINVAL"):
        ops.linear_fwd(obs, w1, b1, h, "elu", x_col0=61)
    w2, b2 = (torch.randn(29, 64, generator=g) / 8).to(DEV), torch.randn(29, generator=g).to(DEV)
    xa = _padded(M, 101, fill=7.0)
    ops.linear_fwd(h, w2, b2, xa, "elu", y_col0=61)
    ref2 = torch.nn.functional.elu(h.double() @ w2.double().t() + b2.double())
    assert _scaled_err(xa[:, 61:90], ref2, h, w2, b2) < TF32_BOUND
    assert bool((xa[:, :61] == 7.0).all()) and bool((xa[:, 90:] == 7.0).all())
    # a wide unaligned output window (several column chunks and n tiles)
    w3, b3 = _padded(200, 64, gen=g), torch.randn(200, generator=g).to(DEV)
    w3.mul_(1 / 8)
    wide = _padded(M, 260, fill=-3.0)
    ops.linear_fwd(h, w3, b3, wide, None, y_col0=37)
    ref3 = h.double() @ w3.double().t() + b3.double()
    assert _scaled_err(wide[:, 37:237], ref3, h, w3, b3) < TF32_BOUND
    assert bool((wide[:, :37] == -3.0).all()) and bool((wide[:, 237:] == -3.0).all())


@pytest.mark.parametrize("M,N,K,act", [(24576, 256, 512, "elu"), (4100, 128, 256, "elu"), (1000, 29, 64, "elu"), (777, 64, 128, "relu")])
def test_linear_bwd_dx_fused_activation_backward_tma_prefetch(M, N, K, act):
    """dx epilogue = act'(y_prev) (y_prev slabs prefetched by TMA) + bias gradient, accumulate and overwrite modes."""
    g = torch.Generator().manual_seed(N)
    gz, w = _padded(M, N, gen=g), _padded(N, K, gen=g)
    w.mul_(1 / N ** 0.5)
    yprev = _padded(M, K, gen=g)
    if act == "relu":
        yprev.clamp_(min=0)
    else:
        yprev.copy_(torch.nn.functional.elu(yprev))
    dx = _padded(M, K)
    db = torch.full((K,), 3.0, device=DEV)
    ops.linear_bwd(gz, None, w, dx=dx, act_prev=act, y_prev=yprev, db_prev=db, db_accumulate=True)
    raw = gz.double() @ w.double()
    d = torch.where(yprev.double() > 0, torch.ones_like(raw), (yprev.double() + 1) if act == "elu" else torch.zeros_like(raw))
    ref = raw * d
    scale = (gz.double().abs() @ w.double().abs()).clamp(min=1e-6) * d.abs().clamp(min=1e-3)
    assert float(((dx.double() - ref).abs() / scale).max()) < TF32_BOUND
    assert_close("db accumulate", db, (3.0 + dx.double().sum(0)).float(), rtol=1e-4, atol=1e-3 * M ** 0.5 / 16)
    ops.linear_bwd(gz, None, w, dx=dx, act_prev=act, y_prev=yprev, db_prev=db)
    assert_close("db overwrite", db, dx.double().sum(0).float(), rtol=1e-4, atol=1e-3 * M ** 0.5 / 16)


def test_linear_bwd_column_windows():
    """dX w.r.t. a 16-byte aligned window of the input row (actor layer 1 -> the window that covers the 29 latent lanes) and dW
    from an aligned window of a wider input; unaligned windows are refused."""
    g = torch.Generator().manual_seed(4)
    M = 6000
    gz, w = _padded(M, 512, gen=g), _padded(512, 101, gen=g)
    dx = _padded(M, 32, fill=9.0)
    ops.linear_bwd(gz, None, w, dx=dx, w_col0=60, K=32)
    ref = gz.double() @ w.double()[:, 60:92]
    scale = gz.double().abs() @ w.double().abs()[:, 60:92]
    assert float(((dx.double() - ref).abs() / scale.clamp(min=1e-6)).max()) < TF32_BOUND
    with pytest.raises(RuntimeError, match="QA_EINVAL"):
        ops.linear_bwd(gz, None, w, dx=dx, w_col0=61, K=29)
    obs = _padded(M, 671, gen=g)
    gz1 = _padded(M, 64, gen=g)
    dw = _padded(64, 29)
    ops.linear_bwd(gz1, obs, None, dw=dw, x_col0=60, K=29)
    ref = gz1.double().t() @ obs.double()[:, 60:89]
    scale = gz1.double().abs().t() @ obs.double().abs()[:, 60:89]
    assert float(((dw.double() - ref).abs() / scale.clamp(min=1e-6)).max()) < TF32_BOUND


def test_act_bwd_with_second_upstream_gradient():
    g = torch.Generator().manual_seed(5)
    M, N = 4099, 29
    gy, y, add = _padded(M, N, gen=g), _padded(M, 101, gen=g)[:, 61:90], _padded(M, N, gen=g)
    coef = torch.full((), 0.37, device=DEV)
    gz, db = _padded(M, N), torch.full((N,), 2.0, device=DEV)
    ops.act_bwd(gy, y, "elu", gz=gz, db=db, zero_db=False, addend=add, addend_scale=coef)
    up = gy.double() + 0.37 * add.double()
    ref = torch.where(y.double() > 0, up, up * (y.double() + 1))
    assert_close("gz", gz, ref.float(), rtol=1e-5, atol=1e-6)
    assert_close("db", db, (2.0 + ref.sum(0)).float(), rtol=1e-4, atol=1e-3)


def test_gather_windows_builds_actor_input_row():
    g = torch.Generator().manual_seed(6)
    R, M = 5000, 1234
    obs = torch.randn(R, 671, generator=g).to(DEV)
    lat = torch.randn(R, 29, generator=g).to(DEV)
    idx = torch.randperm(R, generator=g)[:M].to(DEV)
    xa, ob, hl = _padded(M, 101, fill=5.0), _padded(M, 671), _padded(M, 29)
    li = _padded(M, 29)
    ops.gather_minibatch_windows(idx, [(obs, 0, ob, 0, 671), (obs, 0, xa, 0, 61), (obs, 660, xa, 90, 11), (lat, 0, hl, 0, 29),
                                       (obs, 61, li, 0, 29)])
    assert torch.equal(ob, obs[idx]) and torch.equal(hl, lat[idx]) and torch.equal(li, obs[idx][:, 61:90])
    assert torch.equal(xa[:, :61], obs[idx][:, :61]) and torch.equal(xa[:, 90:], obs[idx][:, 660:])
    assert bool((xa[:, 61:90] == 5.0).all())


# ---- the scheduled step -------------------------------------------------------------------------------------------------------
def _load_golden():
    z = np.load(f"{GOLD}/trainer_policy_seed3.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _sample_from_golden(g):
    b = {k: g["in.batch." + k].to(DEV) for k in ("actions", "target_values", "advantages", "returns", "old_actions_log_prob",
                                                  "old_mu", "old_sigma")}
    obs = g["in.obs"].to(DEV)
    return (obs, obs, b["actions"], b["target_values"], b["advantages"], b["returns"], b["old_actions_log_prob"], b["old_mu"],
            b["old_sigma"], (None, None), None)


def test_plan_step_matches_reference_golden(tc_mode):
    """`update_actor_critic(sample)` on the static schedule (tcgen05 layers) against the UNMODIFIED reference's statistics,
    adaptive learning rate and post-step parameters (tests/golden/trainer_policy_seed3.npz)."""
    from test_trainer_gpu import build
    g = _load_golden()
    alg, env, norm = build(synthetic.make_weights(3))
    alg.priv_reg_counter = int(g["in.priv_reg_counter"])
    out = alg.update_actor_critic(_sample_from_golden(g))
    assert alg._plan is not None, "the static schedule must be the path under test"
    torch.cuda.synchronize()
    for v, k in zip(out, ("surrogate_loss", "value_loss", "b_loss", "entropy", "priv_reg_loss", "estimator_loss")):
        assert_close(f"ppo.{k}", v.cpu(), g[f"ppo.{k}"], rtol=TC_STAT_RTOL, atol=1e-4)
    assert abs(alg.lr_ac - float(g["ppo.lr_new"])) < 1e-9
    stride = int(g["in.param_stride"])
    ac_flat = torch.cat([v.reshape(-1) for v in alg.actor_critic.state_dict().values()])[::stride].cpu()
    est_flat = torch.cat([v.reshape(-1) for v in alg.estimator.state_dict().values()])[::stride].cpu()
    # Adam's first step is lr * g / (|g| + eps): sign-like, so an entry whose gradient is ~0 can flip; bound = 2 * lr per entry,
    # and >= 99 % of the sampled entries within the TF32-perturbed fp32 bar
    lr = 1e-3
    for name, got, want in (("ac", ac_flat, g["ppo.ac_params_sampled"]), ("est", est_flat, g["ppo.est_params_sampled"])):
        err = (got - want).abs()
        assert float(err.max()) <= 2.5 * lr, name
        assert float((err <= 2e-4 + 1e-3 * want.abs()).float().mean()) > 0.99, name


def test_plan_gradients_match_fp32_autograd_path(tc_mode):
    """Every flat gradient entry of the scheduled step (forward + hand-written backward) against the autograd path on fp32
    cuBLAS layers, same minibatch: relative L2 error of each parameter tensor's gradient below the TF32 bound compounded over
    the depth of the network."""
    from test_trainer_gpu import build
    g = _load_golden()
    w = synthetic.make_weights(3)
    sample = _sample_from_golden(g)
    coef = OT.priv_reg_coef(1500)
    # fp32 autograd
    linear.set_mode("fp32")
    ref, _, _ = build(w)
    ref._alloc_minibatch(64)
    ref._kl = torch.zeros((), device=DEV)
    for k, v in zip(("obs", "critic_obs", "actions", "values", "advantages", "returns", "old_actions_log_prob", "old_mu", "old_sigma"),
                    sample[:9]):
        ref._mb[k].copy_(v.reshape(ref._mb[k].shape))
    with torch.no_grad():
        ref._mb["hist_latent"].copy_(ref.actor_critic.infer_hist_latent(ref._mb["obs"][:, 90:660]))
    ref._priv_reg_coef.fill_(coef)
    ref._forward_backward()
    # static schedule
    linear.set_mode("tc")
    alg, _, _ = build(w)
    alg._kl = torch.zeros((), device=DEV)
    plan = alg._ensure_plan(64)
    assert plan is not None
    plan.load(0, sample)
    with torch.no_grad():
        plan.sets[0]["hist_latent"].copy_(alg.actor_critic.infer_hist_latent(plan.sets[0]["obs"][:, 90:660]))
    alg._priv_reg_coef.fill_(coef)
    plan.forward_backward(0)
    torch.cuda.synchronize()
    for (name, p), (_, q) in zip(list(alg.actor_critic.named_parameters()) + list(alg.estimator.named_parameters()),
                                 list(ref.actor_critic.named_parameters()) + list(ref.estimator.named_parameters())):
        a, b = p.grad.double(), q.grad.double()
        if float(b.norm()) == 0.0:
            assert float(a.norm()) == 0.0, name          # frozen history encoder: no gradient in either path
            continue
        rel = float((a - b).norm() / b.norm())
        assert rel < 1e-2, f"{name}: relative L2 gradient error {rel:.3e}"
    assert_close("ppo stats", alg._ppo_stats, ref._ppo_stats, rtol=TC_STAT_RTOL, atol=1e-5)
    assert_close("aux losses", alg._aux_loss, ref._aux_loss, rtol=TC_STAT_RTOL, atol=1e-5)


def test_plan_update_full_size_in_cuda_graph_matches_oracle(tc_mode):
    """BASELINE config 0 at full size on the path the bench times: `update()` over a recorded 4096 x 24 storage = K5, K11, four
    K6 gathers, 20 replays of the captured static-schedule graphs.  The mean statistics of the FIRST epoch's first minibatch
    are checked against the CPU oracle's loss graph (later minibatches see updated parameters); the whole update is checked for
    self-consistency against the eager (un-captured) schedule: same statistics and same final parameters."""
    from test_trainer_gpu import build
    N, T = 4096, 24
    w = synthetic.make_weights(3)
    runs = []
    for graph in (True, False):
        alg, env, norm = build(w, n_envs=N)
        alg.use_cuda_graph = graph
        st = alg.storage
        g = torch.Generator().manual_seed(2)
        st.observations.copy_(0.5 * torch.randn(T, N, 671, generator=g))
        st.privileged_observations.copy_(st.observations)
        linear.set_mode("fp32")
        with torch.no_grad():
            for t in range(T):
                o = st.observations[t]
                alg.act(o, o, normal_draw=torch.randn(N, 12, generator=g).to(DEV))
                tr = alg.transition
                st.actions[t], st.values[t] = tr.actions, tr.values
                st.actions_log_prob[t, :, 0], st.mu[t], st.sigma[t] = tr.actions_log_prob, tr.action_mean, tr.action_sigma
        linear.set_mode("tc")
        st.mu.add_(0.05 * torch.randn(T, N, 12, generator=g).to(DEV))
        st.sigma.mul_(1.05)
        st.actions_log_prob.add_(0.05 * torch.randn(T, N, 1, generator=g).to(DEV))
        st.rewards.copy_(0.05 * torch.rand(T, N, 1, generator=g))
        st.dones.copy_((torch.rand(T, N, 1, generator=g) < 0.02).byte())
        st.compute_returns(torch.zeros(N, 1, device=DEV), 0.99, 0.95)
        idx = torch.randperm(T * N, generator=g).to(DEV)
        alg.priv_reg_counter = 1500
        if graph:
            # oracle on the first minibatch with the initial weights
            rows = idx[:T * N // 4].cpu()
            v = st.flat_views()
            batch = dict(obs=v["obs"].cpu()[rows], critic_obs=v["critic_obs"].cpu()[rows], actions=v["actions"].cpu()[rows],
                         target_values=v["values"].cpu()[rows], advantages=v["advantages"].cpu()[rows], returns=v["returns"].cpu()[rows],
                         old_actions_log_prob=v["old_actions_log_prob"].cpu()[rows], old_mu=v["old_mu"].cpu()[rows],
                         old_sigma=v["old_sigma"].cpu()[rows])
            with torch.no_grad():
                L = OT.ppo_losses(w["ac"], w["est"], batch, priv_reg_coef=OT.priv_reg_coef(1500))
            # one scheduled step on that minibatch (eager), statistics only -- then restore and run the full update
            alg._alloc_minibatch(T * N // 4)
            alg._kl = torch.zeros((), device=DEV)
            alg._encode_history()
            plan = alg._ensure_plan(T * N // 4)
            plan.gather(0, idx[:T * N // 4], alg._hist_latent_all)
            alg._priv_reg_coef.fill_(OT.priv_reg_coef(1500))
            plan.forward_backward(0)
            torch.cuda.synchronize()
            ps, ax = alg._ppo_stats.cpu(), alg._aux_loss.cpu()
            for got, k in ((ps[0], "surrogate_loss"), (ps[1], "value_loss"), (ps[2], "b_loss"), (ax[0], "priv_reg_loss"),
                           (ax[1], "estimator_loss"), (ps[3], "kl_mean")):
                assert_close(f"first minibatch {k}", got, L[k].float(), rtol=TC_STAT_RTOL, atol=1e-5)
        stats = alg.update(indices=idx)
        torch.cuda.synchronize()
        assert alg._plan is not None and (not graph or len(alg._graphs) == 4)
        runs.append((stats, alg.ac_flat.data.clone(), alg.est_flat.data.clone(), alg.lr_ac))
    (s0, a0, e0, lr0), (s1, a1, e1, lr1) = runs
    assert lr0 == lr1
    for x, y in zip(s0, s1):
        assert abs(x - y) <= 1e-4 + 1e-3 * abs(y)         # split-K / atomic summation order differs between runs
    assert float((a0 - a1).abs().max()) <= 2.5e-3 and float(((a0 - a1).abs() < 1e-4).float().mean()) > 0.98
    assert float((e0 - e1).abs().max()) <= 2.5e-3


# ---- rollout schedule -----------------------------------------------------------------------------------------------------
def test_head_fwd_wide_input_for_the_discriminator_heads():
    g = torch.Generator().manual_seed(9)
    M, N, Kh = 4096, 7, 256
    h = _padded(M, Kh, gen=g).clamp_(min=0)
    w, b = (torch.randn(8, Kh, generator=g) / 16).to(DEV), torch.randn(8, generator=g).to(DEV)
    y = torch.zeros(M, 8, device=DEV)
    ops.head_fwd(h, w[:N], b[:N], y[:, :N])
    assert_close("heads", y[:, :N], (h.double() @ w[:N].double().t() + b[:N].double()).float(), rtol=1e-5, atol=1e-5)
    assert bool((y[:, N:] == 0).all())


def test_policy_sample_kernel():
    """K22 against Normal(mean, std).sample() / log_prob(.).sum(-1) with the draw injected (gail.py:186-196); the in-kernel Philox
    stream is standard normal, a function of (seed, step, env) only, and advances with the device step counter."""
    g = torch.Generator().manual_seed(3)
    M, A = 4096, 12
    mu, std = _padded(M, A, gen=g), (0.5 + torch.rand(A, generator=g)).to(DEV)
    noise = torch.randn(M, A, generator=g).to(DEV)
    out = {k: torch.zeros(M, A, device=DEV) for k in ("a", "a_st", "mu_st", "sg_st")}
    lp, lp_st = torch.zeros(M, device=DEV), torch.zeros(M, device=DEV)
    ops.policy_sample(mu, std, out["a"], noise=noise, logp=lp, actions_st=out["a_st"], logp_st=lp_st, mu_st=out["mu_st"], sigma_st=out["sg_st"])
    dist = torch.distributions.Normal(mu, std.expand_as(mu))
    want_a = mu + std * noise
    assert_close("actions", out["a"], want_a, rtol=1e-6, atol=1e-6)
    assert_close("log_prob", lp, dist.log_prob(want_a).sum(-1), rtol=1e-5, atol=1e-5)
    assert torch.equal(out["a"], out["a_st"]) and torch.equal(lp, lp_st)
    assert torch.equal(out["mu_st"], mu.contiguous()) and torch.equal(out["sg_st"], std.expand(M, A).contiguous())
    state = torch.tensor([41], device=DEV, dtype=torch.int64)
    draws = []
    for s_ in (41, 41, 42):
        state.fill_(s_)
        a = torch.zeros(M, A, device=DEV)
        ops.policy_sample(torch.zeros(M, A, device=DEV), torch.ones(A, device=DEV), a, rng_seed=7, step_state=state)
        draws.append(a)
    assert torch.equal(draws[0], draws[1]) and not torch.equal(draws[0], draws[2])
    b = torch.zeros(M, A, device=DEV)
    ops.policy_sample(torch.zeros(M, A, device=DEV), torch.ones(A, device=DEV), b, rng_seed=7, rng_step=42)     # host counter: same stream
    assert torch.equal(b, draws[0])
    z = draws[0].double()
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02 and float(z.abs().max()) < 6.5
    assert abs(float((z ** 3).mean())) < 0.05 and abs(float((z ** 4).mean()) - 3.0) < 0.15
    cols = torch.corrcoef(z.t())
    assert float((cols - torch.eye(A, dtype=torch.float64, device=DEV)).abs().max()) < 0.06


def test_rollout_schedule_matches_torch_rollout(tc_mode):
    """A 4-step rollout through `rollout_plan.RolloutPlan` (tcgen05 layers, head kernels, fused sample / storage writes, K18/K19)
    against the torch restatement of on_policy_runner.py:156-181 on fp32 cuBLAS layers, same action noise (torch generator):
    masks / replay bookkeeping exact, network-dependent quantities within the TF32 bound of the module docstring."""
    import bench
    from qa_b200.pipeline import BbcIteration
    torch.backends.cuda.matmul.allow_tf32 = False
    N, T = 256, 4
    cfg, static, snaps, table = bench.build_workload(0, DEV, n_envs=N, steps=T)
    out = []
    for plan in (False, True):
        linear.set_mode("tc" if plan else "fp32")
        it = BbcIteration(cfg, static, snaps, table, device=DEV, seed=77, use_cuda_graph=False)
        it.runner.fused_rollout = plan
        it.env.task_obs_weight = 0.7
        if plan:
            rp = it.runner._ensure_rollout_plan()
            assert rp is not None
            rp.use_torch_generator = True
        torch.manual_seed(5)
        it._rollout_eager(host=False)
        torch.cuda.synchronize()
        st, ds = it.runner.alg.storage, it.runner.alg.disc_storage
        assert st.step == T
        out.append({k: getattr(st, k).clone() for k in ("rewards", "dones", "values", "observations", "privileged_observations", "actions",
                                                        "actions_log_prob", "mu", "sigma")}
                   | dict(hist=it.runner._disc_hist.clone(), replay=ds.states[:T * N].clone(), replay_eps=ds.latent_eps[:T * N].clone(),
                          replay_c=ds.latent_c[:T * N].clone(), n=ds.num_samples))
    a, b = out
    assert int(a["dones"].sum()) > 0 and a["n"] == b["n"] == T * N
    assert torch.equal(a["dones"], b["dones"]) and torch.equal(a["sigma"], b["sigma"])
    for k in ("values", "mu", "actions"):
        assert_close(k, b[k], a[k], rtol=TC_STAT_RTOL, atol=5e-3)
    assert_close("log_prob", b["actions_log_prob"], a["actions_log_prob"], rtol=TC_STAT_RTOL, atol=2e-2)
    assert_close("rewards", b["rewards"], a["rewards"], rtol=TC_STAT_RTOL, atol=1e-3)
    # observations: only the last-action lanes depend on the policy output
    assert_close("obs", b["observations"], a["observations"], rtol=TC_STAT_RTOL, atol=5e-3)
    quiet = [i for i in range(671) if not (29 <= i < 41 or 90 <= i < 660)]
    assert torch.equal(b["observations"][..., quiet], a["observations"][..., quiet])
    assert torch.equal(b["observations"], b["privileged_observations"])
    assert_close("disc history", b["hist"], a["hist"])
    assert_close("replay states", b["replay"], a["replay"])
    assert torch.equal(b["replay_eps"], a["replay_eps"]) and torch.equal(b["replay_c"], a["replay_c"])


@pytest.mark.parametrize("graph", [False, True])
def test_rollout_deferred_reward_tail_equals_in_order_tail(tc_mode, graph):
    """The reward tail of step t (discriminator trunk, heads, K19) on its own stream next to step t+1 -- with K18's snapshots of
    rew_buf / reset_buf / latched time-outs and the labels read from the storage slot -- fills the storage with exactly what
    the in-order schedule does (same kernels on the same values: bit-equal), eagerly and as one captured rollout graph."""
    import bench
    from qa_b200.pipeline import BbcIteration
    N, T = 512, 8
    cfg, static, snaps, table = bench.build_workload(0, DEV, n_envs=N, steps=T)
    out = []
    for defer in (False, True):
        it = BbcIteration(cfg, static, snaps, table, device=DEV, seed=77, use_cuda_graph=graph)
        it.runner.full_rollouts = defer
        rp = it.runner._ensure_rollout_plan()
        assert rp is not None and rp.defer_reward_tail == defer
        for _ in range(2):                                        # two rollouts: the second starts from carried state
            it.runner.alg.storage.clear()
            it._rollout(host=False)
        torch.cuda.synchronize()
        st = it.runner.alg.storage
        out.append({k: getattr(st, k).clone() for k in ("rewards", "dones", "values", "observations", "actions")}
                   | dict(hist=it.runner._disc_hist.clone()))
    a, b = out
    assert int(a["dones"].sum()) > 0 and float(a["rewards"].abs().sum()) > 0
    for k in a:
        assert torch.equal(a[k], b[k]), k

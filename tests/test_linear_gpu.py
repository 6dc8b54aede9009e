"""GPU numerics of K7 (tcgen05 TF32 GEMM + fused bias/activation epilogue) against a plain PyTorch fp32
reference of the same op.  TF32 keeps 10 mantissa bits of each operand (fp32 accumulate), so the tolerance is
2e-3 relative to the row's scale -- stated here, not the 1e-5 bar of the env/GAE kernels."""
import pytest
import torch
import torch.nn.functional as F

from qa_b200 import ops, synthetic
from qa_b200.config import bbc_train_cfg
from qa_b200.rsl_rl import ActorCritic, linear

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

SHAPES = [  # (M, N, K, x_pitch, w_pitch)
    (24576, 512, 671, 672, 672), (4096, 512, 101, 104, 104), (4096, 256, 512, 512, 512), (4096, 128, 256, 256, 256),
    (4096, 12, 128, 128, 128), (4096, 1, 128, 128, 128), (300, 64, 57, 672, 60), (128, 29, 64, 64, 64),
    (1, 16, 32, 32, 32), (1000, 512, 98, 100, 100), (129, 5, 256, 256, 256), (4096, 4, 64, 64, 64),
    # tall problems run on CTA pairs (cta_group::2, M >= 8192): ragged last pair (second CTA partly / wholly out of range),
    # N that is not a multiple of the tile, every pair tile width
    (10000, 512, 671, 672, 672), (8200, 200, 300, 300, 300), (9001, 64, 128, 128, 128), (8192 + 130, 100, 57, 60, 60),
    (24576, 256, 512, 512, 512), (24576, 128, 256, 256, 256),
]


def ref(x, w, b, act):
    y = F.linear(x.double(), w.double(), None if b is None else b.double())
    if act == "elu":
        y = F.elu(y)
    elif act == "relu":
        y = F.relu(y)
    return y


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("act", [None, "elu", "relu"])
def test_linear_fwd_matches_fp32_reference(shape, act):
    M, N, K, xp, wp = shape
    g = torch.Generator().manual_seed(M + 7 * N + 13 * K)
    xbuf = torch.randn(M, xp, generator=g).to(DEV)
    wbuf = (torch.randn(N, wp, generator=g) / K ** 0.5).to(DEV)
    b = (0.1 * torch.randn(N, generator=g)).to(DEV)
    x, w = xbuf[:, :K], wbuf[:, :K]
    assert ops.linear_tc_ok(x, w)
    y = torch.full((M, (N + 3) // 4 * 4), float("nan"), device=DEV)[:, :N]      # 16-byte row pitch (TMA store)
    ops.linear_fwd(x, w, b, y, act)
    torch.cuda.synchronize()
    want = ref(x, w, b, act)
    scale = (x.double().abs() @ w.double().abs().t() + b.double().abs()).clamp(min=1e-6)
    err = ((y.double() - want).abs() / scale).max().item()
    assert torch.isfinite(y).all()
    assert err < 2e-3, f"{shape} act={act}: max scaled err {err:.3e}"
    # without bias
    ops.linear_fwd(x, w, None, y, act)
    err = ((y.double() - ref(x, w, None, act)).abs() / scale).max().item()
    assert err < 2e-3


def test_linear_rejects_unaligned_operands():
    x = torch.randn(64, 101, device=DEV)            # pitch 101 floats: not a 16-byte multiple
    w = torch.randn(32, 101, device=DEV)
    assert not ops.linear_tc_ok(x, w)
    with pytest.raises(RuntimeError, match="QA_EINVAL"):
        ops.linear_fwd(x, w, None, torch.empty(64, 32, device=DEV), None)
    xa, wa = torch.randn(64, 104, device=DEV)[:, :101], torch.randn(32, 104, device=DEV)[:, :101]
    with pytest.raises(RuntimeError, match="QA_EINVAL"):                       # output pitch 30 floats: not TMA-storable
        ops.linear_fwd(xa, wa, None, torch.empty(64, 30, device=DEV), None)


def test_actor_critic_tc_mode_matches_fp32_mode_and_grads():
    w = synthetic.make_weights(3)
    ac = ActorCritic(101, 671, 12, 57, 10, 4, 29, 11, **bbc_train_cfg()["policy"]).to(DEV)
    ac.load_state_dict(w["ac"])
    flat = ac.flatten_parameters()
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(2048, 672, generator=g).to(DEV)[:, :671]
    outs = {}
    for mode in ("fp32", "tc"):
        linear.set_mode(mode)
        flat.zero_grad()
        mean = ac.act_inference(obs, hist_encoding=False)
        val = ac.evaluate(obs)
        (mean.square().mean() + val.square().mean()).backward()
        outs[mode] = (mean.detach().clone(), val.detach().clone(), flat.grad.clone())
    linear.set_mode("fp32")
    for a, b, name in zip(outs["tc"], outs["fp32"], ("mean", "value", "grad")):
        denom = b.abs().max().item()
        assert (a - b).abs().max().item() < 5e-3 * denom, name


BWD_SHAPES = [  # (M, N, K, x_pitch, w_pitch)
    (24576, 512, 671, 672, 672), (24576, 256, 512, 512, 512), (24576, 128, 256, 256, 256), (24576, 12, 128, 128, 128),
    (4096, 512, 101, 104, 104), (1000, 64, 57, 672, 60), (300, 4, 64, 64, 64), (129, 128, 32, 32, 32), (4096, 64, 128, 128, 128),
    # CTA pairs: ragged M (dX: last pair partly out of range; dW: the reduction's last K block partly out of range), N = 300
    # output features (dW: second CTA of the second pair partly out of range), K not a multiple of the n tile
    (10000, 512, 671, 672, 672), (8200, 300, 200, 200, 300), (9001, 256, 100, 100, 256), (24576, 512, 101, 104, 104),
]


@pytest.mark.parametrize("shape", BWD_SHAPES)
def test_linear_bwd_matches_fp32_reference(shape):
    """dx = gz W (A K-major, B MN-major) and dw += gz^T x (both MN-major, split-K + atomics) on tcgen05."""
    M, N, K, xp, wp = shape
    g = torch.Generator().manual_seed(3 * M + N + K)
    x = torch.randn(M, xp, generator=g).to(DEV)[:, :K]
    w = (torch.randn(N, wp, generator=g) / K ** 0.5).to(DEV)[:, :K]
    gz = (torch.randn(M, N, generator=g) / N ** 0.5).to(DEV)
    assert ops.linear_bwd_ok(gz, x, w)
    dx = torch.full((M, (K + 3) // 4 * 4), float("nan"), device=DEV)[:, :K]
    dw_buf = torch.zeros(N, wp, device=DEV)
    dw = dw_buf[:, :K]
    dw.fill_(0.25)                                          # accumulate semantics: += on top of existing content
    ops.linear_bwd(gz, x, w, dx=dx, dw=dw)
    torch.cuda.synchronize()
    want_dx = gz.double() @ w.double()
    want_dw = gz.double().t() @ x.double() + 0.25
    sx = (gz.double().abs() @ w.double().abs()).clamp(min=1e-6)
    sw = (gz.double().abs().t() @ x.double().abs()).clamp(min=1e-6)
    assert torch.isfinite(dx).all() and torch.isfinite(dw).all()
    ex = ((dx.double() - want_dx).abs() / sx).max().item()
    ew = ((dw.double() - want_dw).abs() / sw).max().item()
    assert ex < 2e-3, f"{shape}: dx scaled err {ex:.3e}"
    assert ew < 2e-3, f"{shape}: dw scaled err {ew:.3e}"
    assert float(dw_buf[:, K:].abs().sum()) == 0.0          # the row padding of the flat layout is never touched
    # each half alone
    dx2 = torch.empty(M, (K + 3) // 4 * 4, device=DEV)[:, :K]
    ops.linear_bwd(gz, None, w, dx=dx2)
    assert torch.equal(dx2, dx)


@pytest.mark.parametrize("act", ["elu", "relu"])
def test_linear_bwd_fused_activation_backward(act):
    """dx epilogue fused with the previous layer's activation derivative and bias gradient (EPI_ACTBWD)."""
    g = torch.Generator().manual_seed(11)
    for M, N, K in ((24576, 256, 512), (1000, 128, 256), (4096, 12, 128), (130, 64, 100), (10001, 300, 200), (9000, 64, 128)):
        kp = (K + 3) // 4 * 4
        z_prev = torch.randn(M, kp, generator=g).to(DEV)[:, :K]
        y_prev = F.elu(z_prev) if act == "elu" else F.relu(z_prev)
        yp = torch.empty(M, kp, device=DEV)[:, :K]
        yp.copy_(y_prev)
        w = (torch.randn(N, kp, generator=g) / K ** 0.5).to(DEV)[:, :K]
        gz = (torch.randn(M, N, generator=g) / N ** 0.5).to(DEV)
        out = torch.full((M, kp), float("nan"), device=DEV)[:, :K]
        db = torch.full((K,), 3.0, device=DEV)
        ops.linear_bwd(gz, None, w, dx=out, act_prev=act, y_prev=yp, db_prev=db)
        torch.cuda.synchronize()
        dact = torch.where(z_prev > 0, 1.0, torch.exp(z_prev)) if act == "elu" else (z_prev > 0).float()
        want = (gz.double() @ w.double()) * dact.double()
        scale = (gz.double().abs() @ w.double().abs()).clamp(min=1e-6)
        assert torch.isfinite(out).all()
        assert ((out.double() - want).abs() / scale).max().item() < 2e-3, (M, N, K)
        sdb = (want.abs().sum(0)).clamp(min=1e-3)
        assert ((db.double() - want.sum(0)).abs() / sdb).max().item() < 2e-3, (M, N, K)


def test_mlp_chain_gradients_match_fp32_autograd():
    """The fused GEMM chain (forward + backward with fused epilogues) against plain fp32 autograd."""
    from qa_b200.rsl_rl.linear import mlp_chain
    g = torch.Generator().manual_seed(2)
    M = 4096
    dims = [671, 512, 256, 128, 12]
    x = torch.randn(M, 672, generator=g).to(DEV)[:, :671]
    params = []
    for i in range(4):
        kp = (dims[i] + 3) // 4 * 4
        w = (torch.randn(dims[i + 1], kp, generator=g) / dims[i] ** 0.5).to(DEV)[:, :dims[i]].requires_grad_(True)
        b = (0.1 * torch.randn(dims[i + 1], generator=g)).to(DEV).requires_grad_(True)
        params.append((w, b))
    acts = ["elu", "elu", "elu", None]
    tgt = torch.randn(M, 12, generator=g).to(DEV)
    outs = {}
    for mode in ("fp32", "tc"):
        linear.set_mode(mode)
        for w, b in params:
            w.grad = None
            b.grad = None
        y = mlp_chain(x, params, acts)
        ((y - tgt) ** 2).mean().backward()
        outs[mode] = [y.detach().clone()] + [t.grad.clone() for wb in params for t in wb]
    linear.set_mode("fp32")
    errs = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(outs["tc"], outs["fp32"])]
    assert max(errs) < 1e-2, errs

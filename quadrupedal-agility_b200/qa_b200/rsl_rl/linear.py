"""`linear_act(x, W, b, act)` = act(x @ W^T + b): the single dense-contraction site of the hot path.

Forward: the hand-written tcgen05 kernel `qa_linear_fwd` (K7: TMA-fed, TF32 operands, fp32 accumulate in TMEM, bias +
ELU/ReLU fused in the epilogue) whenever the operands satisfy TMA's 16-byte pitch/alignment rules -- `FlatParams`
lays the weights out so that they do.  Backward: K9 turns the upstream gradient into the pre-activation gradient from
the saved OUTPUT (ELU'(z) = y + 1 for z <= 0) and reduces the bias gradient in the same pass; `qa_linear_bwd` runs
dX = dZ W and dW += dZ^T X on the same tcgen05 kernel with MN-major operands (dW: split-K over the batch, fp32 atomics
straight into the flat gradient buffer).  Layers whose widths break the 16-byte rule (N = 1, 29) fall back to cuBLAS.

The tcgen05 path is the DEFAULT.  `set_mode("fp32")` (or QA_LINEAR_MODE=fp32) routes everything through `F.linear` in full
fp32 -- the parity-test path, which `tests/conftest.py` selects for the whole test session (the TF32-vs-fp32 tests switch
modes themselves).
"""
import os

import torch
import torch.nn.functional as F

from .. import ops

# "tc": tcgen05 TF32 operands / fp32 accumulate, forward and both backward contractions | "fp32": cuBLAS fp32 (parity tests)
_MODE = os.environ.get("QA_LINEAR_MODE", "tc")
if _MODE not in ("fp32", "tc"):
    raise ValueError(f"QA_LINEAR_MODE must be 'tc' or 'fp32', not {_MODE!r}")


def set_mode(mode: str) -> None:
    global _MODE
    if mode not in ("fp32", "tc"):
        raise ValueError(mode)
    _MODE = mode


def get_mode() -> str:
    return _MODE


def _padded(rows: int, cols: int, device) -> torch.Tensor:
    """(rows, cols) fp32 view whose row pitch is a multiple of 4 floats (16 B): a legal TMA operand."""
    return torch.empty(rows, (cols + 3) // 4 * 4, device=device, dtype=torch.float32)[:, :cols]


def _tma_operand(x: torch.Tensor) -> torch.Tensor:
    """x itself when it already satisfies TMA's base/pitch alignment, else an aligned copy (odd-width inputs such as the
    29 latent lanes sliced out of the observation row, or the 98-wide discriminator input)."""
    if x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0:
        return x
    v = _padded(x.shape[0], x.shape[1], x.device)
    v.copy_(x)
    return v


class _LinearActTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        y = _padded(x.shape[0], weight.shape[0], x.device)
        ops.linear_fwd(x, weight, bias, y, act)
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        if gy.stride(1) != 1:
            gy = gy.contiguous()
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        gb = torch.empty(weight.shape[0], device=gy.device, dtype=torch.float32) if want_b else None
        aligned = gy.stride(0) % 4 == 0 and gy.data_ptr() % 16 == 0
        if ctx.act in ("elu", "relu") or not aligned:
            gz = _padded(gy.shape[0], gy.shape[1], gy.device)   # K9: act'(y) and the bias gradient in one pass
            ops.act_bwd(gy, y if ctx.act else None, ctx.act, gz=gz, db=gb)
        else:
            gz = gy
            if want_b:
                ops.act_bwd(gy, None, None, gz=None, db=gb)
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if ops.linear_bwd_ok(gz, x, weight):
            # both contractions on the tcgen05 kernel (MN-major operands); dW is accumulated by the kernel's split-K
            # atomics straight into the flat gradient buffer when the weight lives in one
            gx = _padded(x.shape[0], x.shape[1], x.device) if need_x else None
            gw = None
            dw_target = None
            if need_w:
                if weight.grad is not None and weight.grad.stride(1) == 1 and weight.grad.stride(0) % 4 == 0:
                    dw_target = weight.grad
                else:
                    kp = (weight.shape[1] + 3) // 4 * 4
                    gw = torch.zeros(weight.shape[0], kp, device=x.device, dtype=torch.float32)[:, :weight.shape[1]]
                    dw_target = gw
            if gx is not None or dw_target is not None:
                ops.linear_bwd(gz, x if dw_target is not None else None, weight if gx is not None else None,
                               dx=gx, dw=dw_target)
            return gx, gw, gb, None
        gx = gz @ weight if need_x else None
        gw = gz.t() @ x if need_w else None
        return gx, gw, gb, None


class _MlpChainTC(torch.autograd.Function):
    """A stack of Linear(+ELU/ReLU) layers as ONE autograd node (K7 GEMM chain).

    forward : one `qa_linear_fwd` per layer (bias + activation in the epilogue); every layer output is saved.
    backward: per layer one split-K `dW += gz^T h` GEMM and one `dx` GEMM whose epilogue already applies the PREVIOUS
              layer's activation derivative (from its saved output) and reduces that layer's bias gradient -- so the
              hidden layers need no separate element-wise pass; only the last layer's upstream gradient goes through K9.
    Weight gradients are accumulated by the kernels straight into the flat gradient buffer (`weight.grad`)."""

    @staticmethod
    def forward(ctx, x, acts, *params):
        n = len(acts)
        hs = [x]
        for i in range(n):
            w, b = params[2 * i], params[2 * i + 1]
            y = _padded(x.shape[0], w.shape[0], x.device)
            ops.linear_fwd(hs[-1], w, b, y, acts[i])
            hs.append(y)
        ctx.acts = acts
        ctx.save_for_backward(*hs, *params)
        return hs[-1]

    @staticmethod
    def backward(ctx, gout):
        acts, n = ctx.acts, len(ctx.acts)
        saved = ctx.saved_tensors
        hs, params = saved[:n + 1], saved[n + 1:]
        dev, M = gout.device, gout.shape[0]
        if gout.stride(1) != 1:
            gout = gout.contiguous()
        grads = [None] * (2 * n)
        # last layer: upstream gradient -> pre-activation gradient (+ bias gradient), K9
        gb = torch.empty(params[2 * n - 2].shape[0], device=dev, dtype=torch.float32)
        aligned = gout.stride(0) % 4 == 0 and gout.data_ptr() % 16 == 0
        if acts[-1] is not None or not aligned:
            gz = _padded(M, gout.shape[1], dev)
            ops.act_bwd(gout, hs[n] if acts[-1] is not None else None, acts[-1], gz=gz, db=gb)
        else:
            gz = gout
            ops.act_bwd(gout, None, None, gz=None, db=gb)
        grads[2 * n - 1] = gb
        gx = None
        for i in range(n, 0, -1):                          # layer i: input hs[i-1], weight params[2(i-1)]
            w = params[2 * (i - 1)]
            h_in = hs[i - 1]
            dw_target = None
            if ctx.needs_input_grad[2 + 2 * (i - 1)]:
                if w.grad is not None and w.grad.stride(1) == 1 and w.grad.stride(0) % 4 == 0:
                    dw_target = w.grad
                else:
                    dw_target = _padded(w.shape[0], w.shape[1], dev).zero_()
                    grads[2 * (i - 1)] = dw_target
            if i > 1:
                act_prev = acts[i - 2]
                gz_prev = _padded(M, h_in.shape[1], dev)
                db_prev = torch.empty(h_in.shape[1], device=dev, dtype=torch.float32)
                if act_prev is not None:
                    ops.linear_bwd(gz, h_in, w, dx=gz_prev, dw=dw_target, act_prev=act_prev, y_prev=h_in, db_prev=db_prev)
                else:
                    ops.linear_bwd(gz, h_in, w, dx=gz_prev, dw=dw_target)
                    ops.act_bwd(gz_prev, None, None, gz=None, db=db_prev)
                grads[2 * (i - 2) + 1] = db_prev
                gz = gz_prev
            else:
                if ctx.needs_input_grad[0]:
                    gx = _padded(M, h_in.shape[1], dev)
                if gx is not None or dw_target is not None:
                    ops.linear_bwd(gz, h_in if dw_target is not None else None, w if gx is not None else None,
                                   dx=gx, dw=dw_target)
        return (gx, None, *grads)


def mlp_chain(x: torch.Tensor, layers, acts):
    """layers: [(weight, bias), ...]; acts: ["elu" | "relu" | None, ...].  tcgen05 chain in "tc" mode, cuBLAS otherwise."""
    if _MODE == "tc" and x.is_cuda and x.dim() == 2 and x.shape[0] > 0:
        xa = _tma_operand(x)
        if all(ops.linear_tc_ok(xa if i == 0 else w, w) and b is not None for i, (w, b) in enumerate(layers)):
            flat = [t for wb in layers for t in wb]
            return _MlpChainTC.apply(xa, tuple(acts), *flat)
    for (w, b), a in zip(layers, acts):
        x = linear_act(x, w, b, a)
    return x


def linear_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, act):
    if _MODE == "tc" and x.is_cuda and x.dim() == 2 and x.shape[0] > 0:
        xa = _tma_operand(x)
        if ops.linear_tc_ok(xa, weight):
            return _LinearActTC.apply(xa, weight, bias, act)
    y = F.linear(x, weight, bias)
    if act == "elu":
        return F.elu(y)
    if act == "relu":
        return F.relu(y)
    return y

"""`linear_act(x, W, b, act)` = act(x @ W^T + b): the single dense-contraction site of the hot path.

Forward: the hand-written tcgen05 kernel `qa_linear_fwd` (K7: TMA-fed, TF32 operands, fp32 accumulate in TMEM, bias +
ELU/ReLU fused in the epilogue) whenever the operands satisfy TMA's 16-byte pitch/alignment rules -- `FlatParams`
lays the weights out so that they do.  Backward: K9 turns the upstream gradient into the pre-activation gradient from
the saved OUTPUT (ELU'(z) = y + 1 for z <= 0) and reduces the bias gradient in the same pass; `qa_linear_bwd` runs
dX = dZ W and dW += dZ^T X on the same tcgen05 kernel with MN-major operands (dW: split-K over the batch, fp32 atomics
straight into the flat gradient buffer).  Layers whose widths break the 16-byte rule (N = 1, 29) fall back to cuBLAS.

`set_mode("fp32")` routes everything through `F.linear` in full fp32 -- the parity-test path.
"""
import torch
import torch.nn.functional as F

from .. import ops

_MODE = "fp32"          # "fp32": cuBLAS fp32 everywhere (parity tests) | "tc": tcgen05 TF32 forward


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

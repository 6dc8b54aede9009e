"""`linear_act(x, W, b, act)` = act(x @ W^T + b): the single dense-contraction site of the hot path.

Forward: the hand-written tcgen05 kernel `qa_linear_fwd` (K7: TMA-fed, TF32 operands, fp32 accumulate in TMEM, bias +
ELU/ReLU fused in the epilogue) whenever the operands satisfy TMA's 16-byte pitch/alignment rules -- `FlatParams`
lays the weights out so that they do.  Backward (training) reuses the saved OUTPUT for the activation derivative
(ELU'(z) = y + 1 for z <= 0) and contracts through cuBLAS (dX = dZ W, dW = dZ^T X); porting the two MN-major backward
contractions to tcgen05 is the next step.

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


class _LinearActTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        y = torch.empty(x.shape[0], weight.shape[0], device=x.device, dtype=torch.float32)
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
        if ctx.act in ("elu", "relu"):
            gz = torch.empty_like(gy)
            ops.act_bwd(gy, y, ctx.act, gz=gz, db=gb)            # K9: act'(y) and the bias gradient in one pass
        else:
            gz = gy
            if want_b:
                ops.act_bwd(gy, None, None, gz=None, db=gb)
        gx = gz @ weight if ctx.needs_input_grad[0] else None
        gw = gz.t() @ x if ctx.needs_input_grad[1] else None
        return gx, gw, gb, None


def linear_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, act):
    if _MODE == "tc" and ops.linear_tc_ok(x, weight):
        return _LinearActTC.apply(x, weight, bias, act)
    y = F.linear(x, weight, bias)
    if act == "elu":
        return F.elu(y)
    if act == "relu":
        return F.relu(y)
    return y

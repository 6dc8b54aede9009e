"""`linear_act(x, W, b, act)` = act(x @ W^T + b): the single dense-contraction site of the hot path.

Round-1 state: the contraction goes to cuBLAS through `F.linear` (library GEMM) and the bias + ELU/ReLU
epilogue is a separate elementwise op; the hand-written tcgen05 kernel with the fused epilogue replaces
this function body ("K7").  Kept as one function so that the swap is local.
"""
import torch
import torch.nn.functional as F


def linear_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, act):
    y = F.linear(x, weight, bias)
    if act == "elu":
        return F.elu(y)
    if act == "relu":
        return F.relu(y)
    return y

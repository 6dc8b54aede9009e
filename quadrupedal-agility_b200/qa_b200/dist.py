"""Env-sharded data parallelism: the host-side rules of SURVEY.md 8(e).

Envs are independent, so rank r of W owns a contiguous shard of the env ids, its own physics instance, its own
RolloutStorage and RNG stream (seed + r); the only coupling is the policy.  Per optimiser step there is ONE in-place
all-reduce(SUM) of each flat gradient buffer (actor-critic 2.95 MB, estimator 64 KB) and one of the scalar KL that
drives the adaptive learning rate; the 1/W scaling is folded into the fused clip+Adam kernel (K8 `grad_scale`).
Equal shards => mean of per-shard mean-loss gradients == gradient of the mean loss over the union batch.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_envs(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first env id, count) of rank's shard; shards must be equal-sized (gradient averaging assumes it)."""
    if total_envs % world_size != 0:
        raise ValueError(f"{total_envs} envs do not split evenly over {world_size} ranks")
    n = total_envs // world_size
    return rank * n, n


def rank_seed(seed: int, rank: int) -> int:
    return seed + rank


def allreduce_flat_(flat_grad: torch.Tensor) -> float:
    """In-place SUM over ranks of a flat gradient buffer; returns the scale (1/W) the optimiser pass must apply."""
    _, w = world()
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / w


def allreduce_mean_scalar_(x: torch.Tensor) -> torch.Tensor:
    """In-place mean over ranks of a (0-d or 1-element) tensor, e.g. kl_mean (gail.py:368-379)."""
    _, w = world()
    if w > 1:
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
        x /= w
    return x


def allreduce_mean_grads_(params) -> None:
    """In-place mean over ranks of the gradients of `params` (those that have one), as ONE all-reduce of a flat staging buffer.
    For the torch-optimiser modules of the TSC depth student (conv / GRU on cuDNN), whose parameters do not live in a
    `FlatParams` buffer; with equal env shards the mean of the per-shard mean-loss gradients is the union batch's gradient."""
    _, w = world()
    if w == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= w
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n

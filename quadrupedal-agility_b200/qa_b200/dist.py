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


class _CudaBlob:
    """A cudaMalloc'ed buffer exposed to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, nbytes: int, typestr: str, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
                                         "strides": None}
        self.nbytes = nbytes


class PeerArena:
    """The PPO gradient arena of every rank of ONE node, mapped into every process (cudaIpc over NVLink peers), for K31
    (`ops.peer_allreduce`): `tensor` = this rank's arena (fp32, n floats, zero-filled), `arena_ptrs[p]` / `ctrl_ptrs[p]` = rank
    p's arena / control block as mapped here.  Handles travel through `dist.all_gather_object`.  `available()` says whether the
    process group qualifies (NCCL backend, all ranks on this node, <= 8 ranks, QA_PEER_ALLREDUCE != 0)."""

    @staticmethod
    def available() -> bool:
        import os
        if os.environ.get("QA_PEER_ALLREDUCE", "1") != "1" or not (dist.is_available() and dist.is_initialized()):
            return False
        w = dist.get_world_size()
        if w < 2 or w > 8 or dist.get_backend() != "nccl":
            return False
        local = int(os.environ.get("LOCAL_WORLD_SIZE", w))
        return local == w

    def __init__(self, n: int, device):
        import ctypes as C
        from . import _abi
        lib = _abi.load()
        self._lib, self.n = lib, int(n)
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
        dev = torch.device(device)
        torch.cuda.set_device(dev)
        ctrl_bytes = int(lib.qa_peer_ctrl_bytes(self.n))
        self._own = []
        handles = []
        for nbytes in (self.n * 4, ctrl_bytes):
            p = C.c_void_p()
            _abi.check(lib.qa_ipc_alloc(C.byref(p), nbytes), "qa_ipc_alloc")
            h = C.create_string_buffer(64)
            _abi.check(lib.qa_ipc_get_handle(p, h), "qa_ipc_get_handle")
            self._own.append(int(p.value))
            handles.append(bytes(h.raw))
        gathered = [None] * self.world_size
        dist.all_gather_object(gathered, handles)
        self.arena_ptrs, self.ctrl_ptrs, self._opened = [], [], []
        for r, (ha, hc) in enumerate(gathered):
            if r == self.rank:
                self.arena_ptrs.append(self._own[0])
                self.ctrl_ptrs.append(self._own[1])
                continue
            ptrs = []
            for h in (ha, hc):
                q = C.c_void_p()
                _abi.check(lib.qa_ipc_open_handle(C.c_char_p(h), C.byref(q)), "qa_ipc_open_handle")
                ptrs.append(int(q.value))
                self._opened.append(int(q.value))
            self.arena_ptrs.append(ptrs[0])
            self.ctrl_ptrs.append(ptrs[1])
        self.tensor = torch.as_tensor(_CudaBlob(self._own[0], self.n * 4, "<f4", (self.n,)), device=dev)
        dist.barrier()                                   # every rank has mapped every arena before the first kernel touches one

    def close(self):
        """Unmap the peers' buffers and free this rank's (after a barrier: nobody may still be reading them)."""
        import ctypes as C
        if getattr(self, "_lib", None) is None:
            return
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        for q in self._opened:
            self._lib.qa_ipc_close_handle(C.c_void_p(q))
        self.tensor = None
        for p in self._own:
            self._lib.qa_ipc_free(C.c_void_p(p))
        self._lib = None

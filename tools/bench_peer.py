#!/usr/bin/env python
"""K31 (peer-memory all-reduce + gradient norms) against NCCL on the PPO gradient arena (2.96 MB), N ranks of one node.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/bench_peer.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
from qa_b200 import _abi, ops  # noqa: E402
TRACE = os.environ.get("QA_PEER_TRACE") == "1"
if TRACE:                                                   # the -DQA_PEER_TRACE build of tools/k2_trace.py --build
    _abi.LIB_PATH = os.path.join(ROOT, "tools", "_k2trace", "libqa_b200.so")
from qa_b200.dist import PeerArena  # noqa: E402

N_AC, N_EST = 722_000, 16_000
n = (N_AC + N_EST + 4 + 3) // 4 * 4
pa = PeerArena(n, dev)
ws = [torch.zeros(2, device=dev, dtype=torch.float64) for _ in range(2)]
steps = [torch.zeros(1, device=dev, dtype=torch.int32) for _ in range(2)]
x = torch.randn(n, device=dev)
REPS = 50


def peer():
    ops.peer_allreduce(pa.world_size, pa.rank, pa.n, pa.arena_ptrs, pa.ctrl_ptrs, seg_split=N_AC, norm_end=N_AC + N_EST,
                       sumsq_out=(ws[0], ws[1]), grad_scale=1.0 / world, step_inc=(steps[0], steps[1]), scale_index=N_AC + N_EST)


def nccl():
    dist.all_reduce(x)


def timed(fn, name):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (5 * REPS) * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:28s} {float(t):7.2f} us per call (max over {world} ranks, {REPS} calls per graph replay)")
    del g


timed(peer, "K31 peer all-reduce + norms")
if TRACE:
    import ctypes
    import numpy as np
    lib = _abi.load()
    lib.qa_peer_trace_dump.restype = ctypes.c_int
    lib.qa_peer_trace_dump.argtypes = [ctypes.c_void_p]
    acc = []
    for _ in range(20):
        dist.barrier()
        peer()
        torch.cuda.synchronize()
        h = np.zeros(8, dtype=np.int64)
        assert lib.qa_peer_trace_dump(h.ctypes.data) == 0
        acc.append(np.diff(h[:5]) / 1e3)
    m = np.median(np.array(acc), axis=0)
    print(f"rank {rank} CTA 0 phases [us]: barrier {m[0]:.2f} | reduce own slice + push flagged words {m[1]:.2f} | poll + unpack the "
          f"peers' slices {m[2]:.2f} | ticket + norms {m[3]:.2f}", flush=True)
x.mul_(0)
timed(nccl, "ncclAllReduce (same bytes)")
dist.barrier()
torch.cuda.synchronize()
dist.destroy_process_group()

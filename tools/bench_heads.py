#!/usr/bin/env python
"""Micro-benchmark of K20 / K21 (qa_head_fwd / qa_head_bwd) on the head shapes of the hot path.
usage (on the GPU box): python tools/bench_heads.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))
import torch  # noqa: E402
from qa_b200 import ops  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_time(body, reps=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


NCALL = 8
T_FLUSH = graph_time(lambda: [flush.fill_(0) for _ in range(NCALL)])


def timeit(fn):
    def body():
        for _ in range(NCALL):
            flush.fill_(0)
            fn()
    return (graph_time(body) - T_FLUSH) / NCALL


print(f"{'M':>6} {'N':>3} {'Kh':>4} | {'fwd us':>7} {'bwd us':>7}")
for M, N, Kh in [(24576, 12, 128), (24576, 1, 128), (24576, 4, 64), (4096, 12, 128), (4096, 7, 256), (4096, 1, 128), (4096, 4, 64), (3684, 7, 256)]:
    h = torch.randn(M, Kh, device=dev)
    w = torch.randn(N, Kh, device=dev) / Kh ** 0.5
    b = torch.randn(N, device=dev)
    y = torch.empty(M, (N + 3) // 4 * 4, device=dev)[:, :N]
    t_f = timeit(lambda: ops.head_fwd(h, w, b, y))
    t_b = float("nan")
    if Kh <= 128:
        gz = torch.randn(M, N, device=dev)
        gp = torch.empty(M, Kh, device=dev)
        dw, db, dbp = torch.zeros(N, Kh, device=dev), torch.zeros(N, device=dev), torch.zeros(Kh, device=dev)
        t_b = timeit(lambda: ops.head_bwd(gz, h, w, "elu", gz_prev=gp, dw=dw, db=db, db_prev=dbp))
    print(f"{M:6d} {N:3d} {Kh:4d} | {t_f:7.1f} {t_b:7.1f}")

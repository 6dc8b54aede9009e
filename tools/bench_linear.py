#!/usr/bin/env python
"""Micro-benchmark of K7 (qa_linear_fwd) against cuBLAS TF32 for the layer shapes of the hot path.
usage (on the GPU box): python tools/bench_linear.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from qa_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda:0"
SHAPES = [(24576, 512, 671), (24576, 512, 101), (24576, 256, 512), (24576, 128, 256), (24576, 12, 128), (24576, 1, 128),
          (24576, 128, 57), (24576, 64, 128), (4096, 512, 671), (4096, 512, 101), (4096, 256, 512), (4096, 128, 256),
          (4096, 12, 128), (4096, 512, 98)]
if len(sys.argv) > 1 and sys.argv[1] == "--only":            # profiling mode: one shape, plain launches, no timing
    M, N, K = (int(v) for v in sys.argv[2].split(","))
    kp = (K + 3) // 4 * 4
    x = torch.randn(M, kp, device=dev)[:, :K]
    w = (torch.randn(N, kp, device=dev) / K ** 0.5)[:, :K]
    b = torch.randn(N, device=dev)
    y = torch.empty(M, (N + 3) // 4 * 4, device=dev)[:, :N]
    gz = torch.randn(M, N, device=dev)
    dx = torch.empty(M, kp, device=dev)[:, :K]
    dw = torch.zeros(N, kp, device=dev)[:, :K]
    db = torch.zeros(K, device=dev)
    xin = torch.nn.functional.elu(x.clone()) if kp == K else None       # previous layer's output for the fused activation backward
    for _ in range(3):
        ops.linear_fwd(x, w, b, y, "elu")
        if xin is not None:
            ops.linear_bwd(gz, None, w, dx=dx, act_prev="elu", y_prev=xin, db_prev=db, db_accumulate=True)
        else:
            ops.linear_bwd(gz, None, w, dx=dx)
        ops.linear_bwd(gz, x, None, dw=dw)
    torch.cuda.synchronize()
    sys.exit(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_time(body, reps=10):
    """us per replay of a CUDA graph holding `body` (no host launch gaps inside the measurement)."""
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
    """us per call, L2 flushed before every call (flush time measured separately and subtracted)."""
    def body():
        for _ in range(NCALL):
            flush.fill_(0)
            fn()
    return (graph_time(body) - T_FLUSH) / NCALL


print(f"{'M':>6} {'N':>4} {'K':>4} | {'tcgen05 us':>10} {'TF/s':>7} | {'cuBLAS+elu us':>13} {'TF/s':>7} | {'GB/s(tc)':>9}")
for M, N, K in SHAPES:
    kp = (K + 3) // 4 * 4
    x = torch.randn(M, kp, device=dev)[:, :K]
    w = (torch.randn(N, kp, device=dev) / K ** 0.5)[:, :K]
    b = torch.randn(N, device=dev)
    y = torch.empty(M, (N + 3) // 4 * 4, device=dev)[:, :N]
    t_tc = timeit(lambda: ops.linear_fwd(x, w, b, y, "elu"))
    t_cb = timeit(lambda: F.elu(F.linear(x, w, b)))
    fl = 2.0 * M * N * K
    by = 4.0 * (M * K + N * K + M * N)
    print(f"{M:6d} {N:4d} {K:4d} | {t_tc:10.1f} {fl / t_tc / 1e6:7.1f} | {t_cb:13.1f} {fl / t_cb / 1e6:7.1f} | {by / t_tc / 1e3:9.0f}")

print()
print(f"{'M':>6} {'N':>4} {'K':>4} | {'tc dx us':>9} {'tc dw us':>9} | {'cuBLAS dx':>9} {'cuBLAS dw':>9}")
for M, N, K in [(24576, 512, 671), (24576, 512, 101), (24576, 256, 512), (24576, 128, 256), (24576, 12, 128), (24576, 64, 128)]:
    kp = (K + 3) // 4 * 4
    x = torch.randn(M, kp, device=dev)[:, :K]
    w = (torch.randn(N, kp, device=dev) / K ** 0.5)[:, :K]
    gz = torch.randn(M, N, device=dev)
    dx = torch.empty(M, kp, device=dev)[:, :K]
    dw = torch.zeros(N, kp, device=dev)[:, :K]
    if kp == K:                                      # hidden layer: dx with the fused activation backward + bias gradient
        yp, db = torch.nn.functional.elu(x.clone()), torch.zeros(K, device=dev)
        t1 = timeit(lambda: ops.linear_bwd(gz, None, w, dx=dx, act_prev="elu", y_prev=yp, db_prev=db, db_accumulate=True))
    else:
        t1 = timeit(lambda: ops.linear_bwd(gz, None, w, dx=dx))
    t2 = timeit(lambda: ops.linear_bwd(gz, x, None, dw=dw))
    t3 = timeit(lambda: torch.matmul(gz, w))
    t4 = timeit(lambda: torch.matmul(gz.t(), x))
    print(f"{M:6d} {N:4d} {K:4d} | {t1:9.1f} {t2:9.1f} | {t3:9.1f} {t4:9.1f}")

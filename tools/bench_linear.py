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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        flush.fill_(0)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


print(f"{'M':>6} {'N':>4} {'K':>4} | {'tcgen05 us':>10} {'TF/s':>7} | {'cuBLAS+elu us':>13} {'TF/s':>7} | {'GB/s(tc)':>9}")
for M, N, K in SHAPES:
    kp = (K + 3) // 4 * 4
    x = torch.randn(M, kp, device=dev)[:, :K]
    w = (torch.randn(N, kp, device=dev) / K ** 0.5)[:, :K]
    b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev)
    t_tc = timeit(lambda: ops.linear_fwd(x, w, b, y, "elu"))
    t_cb = timeit(lambda: F.elu(F.linear(x, w, b)))
    fl = 2.0 * M * N * K
    by = 4.0 * (M * K + N * K + M * N)
    print(f"{M:6d} {N:4d} {K:4d} | {t_tc:10.1f} {fl / t_tc / 1e6:7.1f} | {t_cb:13.1f} {fl / t_cb / 1e6:7.1f} | {by / t_tc / 1e3:9.0f}")

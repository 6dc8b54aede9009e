"""Two launches each of the trainer-side streaming kernels for an ncu capture (K5 GAE, K6 gather, K8 clip+Adam):
   ncu --set full --clock-control none -k regex:"k_gae|k_gather|k_clip_adam|k_grad_sumsq" -o out python tools/prof_trainer_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200")); sys.path.insert(0, ROOT)
import torch
from qa_b200 import ops, synthetic
dev = "cuda:0"
T, N = 24, 4096
g = torch.Generator().manual_seed(0)
r, v = torch.rand(T, N, 1, generator=g).to(dev), torch.randn(T, N, 1, generator=g).to(dev)
d, lv = (torch.rand(T, N, 1, generator=g) < 0.02).to(torch.uint8).to(dev), torch.randn(N, 1, generator=g).to(dev)
ret, adv, ws = torch.empty(T, N, 1, device=dev), torch.empty(T, N, 1, device=dev), torch.zeros(8, dtype=torch.float64, device=dev)
for _ in range(2):
    ops.gae(r, v, d, lv, ret, adv, ws, 0.99, 0.95)
R = T * N
srcs = [torch.randn(R, w, device=dev) for w in (671, 671, 12, 1, 1, 1, 1, 12, 12, 29)]
idx = torch.randperm(R, device=dev)[:R // 4]
dsts = [torch.empty(R // 4, (s.shape[1] + 3) // 4 * 4, device=dev)[:, :s.shape[1]] if s.shape[1] == 671 else torch.empty(R // 4, s.shape[1], device=dev) for s in srcs]
for _ in range(2):
    ops.gather_minibatch(idx, srcs, dsts)
n = 738100
p, gr, m, vv = (torch.randn(n, device=dev) for _ in range(4))
vv.abs_()
lr, step, ws2, gn = torch.full((1,), 1e-3, device=dev), torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(2, dtype=torch.float64, device=dev), torch.zeros(1, device=dev)
for _ in range(2):
    ops.clip_adam(p, gr, m, vv, lr, step, ws2, grad_norm_out=gn)
torch.cuda.synchronize()

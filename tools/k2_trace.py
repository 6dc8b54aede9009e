#!/usr/bin/env python
"""Phase trace + stand-alone timing of the fused post-physics kernel (K2, tiled variant).

  python tools/k2_trace.py --build          (build container) compile a -DQA_K2_TRACE variant of libqa_b200.so into
                                            tools/_k2trace/ (git-ignored, travels to the GPU box)
  python tools/k2_trace.py [--envs N ...]   (GPU box) per-CTA clock64 stamps at the phase boundaries -> median / p90
                                            of every phase, CTA start skew and kernel span from %globaltimer; then the
                                            PRODUCT library is NOT touched: timing of the product build is bench.py's job.
"""
import argparse
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quadrupedal-agility_b200")
OUT = os.path.join(ROOT, "tools", "_k2trace")
sys.path.insert(0, PKG)
sys.path.insert(0, ROOT)

SEGMENTS = [  # (name, stamp_from, stamp_to); stamps 0-11 env warp 0, 12-15/18 scalar warp A, 19 scalar warp B
    ("env: entry -> loads issued, noise drawn", 0, 1),
    ("env: wait small tiles", 1, 2),
    ("env: P1 (contact norms)", 2, 3),
    ("env: P3a row lanes", 3, 4),
    ("env: wait history tile", 4, 5),
    ("env: history shift", 5, 6),
    ("env: barrier 2 wait (P2b)", 6, 7),
    ("env: P3b", 7, 8),
    ("env: P3b end -> barrier 3 stamp (the clock read is hoisted ABOVE the barrier)", 8, 9),
    ("barrier 3 wait (slowest env warp of the CTA, e.g. a reset) + store issue", 9, 10),
    ("reset env: P1 start -> clip chosen (contact ballot, 3 Philox draws, lane-parallel CDF search)", 2, 21),
    ("reset env: clip chosen -> frame operands requested (fp64 index math)", 21, 22),
    ("reset env: operands landed + slerp of the root quaternion", 22, 23),
    ("reset env: lerps, velocity rotation, key-body positions -> end of the precompute", 23, 3),
    ("P4 wait_group.read", 10, 11),
    ("scalar A: entry -> own loads landed", 0, 12),
    ("scalar A: P2a", 12, 13),
    ("scalar A: wait P1 (small tiles + barrier 1)", 13, 14),
    ("scalar A: P2b", 14, 15),
    ("scalar A: barrier 2 + small-output stores", 15, 18),
    ("scalar B: entry -> key positions done", 0, 19),
    ("scalar B: key positions -> DOF sums done (incl. wait for small tiles)", 19, 20),
    ("env: entry -> barrier 2 passed", 0, 7),
    ("whole CTA (entry -> stores issued and read)", 0, 11),
]


def build():
    csrc = os.path.join(PKG, "csrc")
    os.makedirs(OUT, exist_ok=True)
    objs, procs = [], []
    for f in sorted(x for x in os.listdir(csrc) if x.endswith(".cu")):
        exact = f.startswith(("qa_env_kernels", "qa_post_physics", "qa_gae", "qa_depth", "qa_tsc"))
        o = os.path.join(OUT, f[:-3] + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(
            ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
             "-DQA_K2_TRACE", "-DQA_PEER_TRACE", "-I" + os.path.join(ROOT, "include"), "-I" + csrc] + (["-fmad=false"] if exact else []) +
            ["-c", os.path.join(csrc, f), "-o", o]))
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", *objs, "-o",
                    os.path.join(OUT, "libqa_b200.so")], check=True)
    for o in objs:
        os.unlink(o)
    print("built", os.path.join(OUT, "libqa_b200.so"))


def run(n_envs, reps, mode=0):
    import numpy as np
    import torch
    from qa_b200 import _abi
    _abi.LIB_PATH = os.path.join(OUT, "libqa_b200.so")
    lib = _abi.load()
    lib.qa_k2_trace_dump.restype = ctypes.c_int
    lib.qa_k2_trace_dump.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.qa_k2_trace_set_mode.restype = ctypes.c_int
    lib.qa_k2_trace_set_mode.argtypes = [ctypes.c_int]
    import bench
    from qa_b200.legged_robot import LeggedRobot, RecordedPhysics
    dev = torch.device("cuda:0")
    T = 4
    cfg, static, snaps, table = bench.build_workload(0, dev, n_envs=n_envs, steps=T)
    dev_snaps = [{k: s[k].to(dev) for k in bench.SIM_KEYS} for s in snaps]
    env = LeggedRobot(cfg, RecordedPhysics(dev_snaps), static, table, device=dev, seed=1234, bulk_store=True, tiled=True)
    from qa_b200.pipeline import CARRIED
    env.load_state({k: v.to(dev) for k, v in snaps[0].items() if k in CARRIED})
    env.global_counter = 1
    env.use_device_step_counter(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    assert lib.qa_k2_trace_set_mode(mode) == 0
    if mode:
        print(f"#### attribution mode {mode}: bit 0 = the four big output tiles are not stored")
    n_cta = min(n_envs // 8, 1024)
    host = np.zeros((n_cta, 24), dtype=np.int64)
    seg = {name: [] for name, _, _ in SEGMENTS}
    spans, skews, times = [], [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(reps + 2):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0.record()
        env._k2_only_step()
        e1.record()
        torch.cuda.synchronize()
        if r < 2:
            continue
        times.append(e0.elapsed_time(e1) * 1e3)
        rc = lib.qa_k2_trace_dump(host.ctypes.data, n_cta)
        assert rc == 0, rc
        g0, g1 = host[:, 16], host[:, 17]
        spans.append(float(g1.max() - g0.min()) / 1e3)
        skews.append(float(g0.max() - g0.min()) / 1e3)
        cyc_per_ns = np.median((host[:, 11] - host[:, 0]) / np.maximum(g1 - g0, 1))
        for name, a, b in SEGMENTS:
            d = (host[:, b] - host[:, a]) / cyc_per_ns / 1e3
            if name.startswith("reset env"):                    # only the CTAs whose env 0 reset in THIS launch
                d = d[(host[:, 21] > host[:, 2]) & (host[:, 21] < host[:, 3])]
            seg[name].append(d)
    print(f"== K2 tiled phase trace: {n_envs} envs, {n_envs // 8} CTAs (first {n_cta} traced), {reps} launches, L2 flushed; "
          f"clock {cyc_per_ns:.3f} cycles/ns")
    print(f"event time per launch (eager, incl. finalize kernel): median {np.median(times):.2f} us")
    print(f"kernel span, first CTA entry -> last CTA exit: median {np.median(spans):.2f} us;  CTA start skew {np.median(skews):.2f} us")
    for name, _, _ in SEGMENTS:
        v = np.concatenate(seg[name])
        if v.size == 0:
            continue
        print(f"  {name:55s} median {np.median(v):6.2f} us   p90 {np.percentile(v, 90):6.2f}   max {v.max():6.2f}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--envs", type=int, nargs="*", default=[4096, 32768])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--mode", type=int, nargs="*", default=[0], help="attribution modes to run (trace build only), e.g. 0 1")
    args = ap.parse_args()
    if args.build:
        build()
    else:
        for n in args.envs:
            for m in args.mode:
                run(n, args.reps, m)

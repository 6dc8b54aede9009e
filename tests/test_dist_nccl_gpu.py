"""2 x B200, NCCL: device-side results parity of the env-sharded path.  Two ranks run the scheduled PPO minibatch
step (K6-shaped inputs, K7 / K20 / K21 / K10 / K12, ONE all-reduce of the gradient arena, K13, K8) each on its half of a union
minibatch; a single process runs the same step on the union.  The all-reduced gradient (x 1/W) must equal the union gradient
(only the fp32 summation order differs: per-rank partial sums vs one split-K order), and the parameters after the step agree.
Skipped when fewer than two GPUs are visible (run it with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
M_RANK = 4096


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _union_sample(M, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)                                     # noqa: E731
    obs = 0.6 * r(M, 671)
    return (obs, obs.clone(), r(M, 12), r(M, 1), r(M, 1), r(M, 1), -12 + r(M, 1), 0.3 * r(M, 12),
            0.8 + 0.2 * torch.rand(M, 12, generator=g), (None, None), None)


def _run(rank, world, port, q, arena, peer=True):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.join(os.path.dirname(here), "quadrupedal-agility_b200"), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    dev = f"cuda:{rank}"
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), QA_SINGLE_ALLREDUCE="1" if arena else "0",
                          QA_PEER_ALLREDUCE="1" if peer else "0")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    from qa_b200 import synthetic
    from qa_b200.rsl_rl import linear
    import test_trainer_gpu as T
    T.DEV = dev
    linear.set_mode("tc")
    alg, env, norm = T.build(synthetic.make_weights(3), n_envs=64)
    full = _union_sample(M_RANK * 2)
    lo, hi = (rank * M_RANK, (rank + 1) * M_RANK) if world > 1 else (0, 2 * M_RANK)
    sample = tuple(t[lo:hi].to(dev) if torch.is_tensor(t) else t for t in full)
    alg.priv_reg_counter = 1500
    out = alg.update_actor_critic(sample)
    torch.cuda.synchronize()
    assert alg._plan is not None
    scale = 1.0 / world
    # numpy arrays through the queue (tensors would be passed as shared-memory file descriptors of a process that exits)
    res = dict(grad_ac=(alg.ac_flat.grad * scale).cpu().numpy(), grad_est=(alg.est_flat.grad * scale).cpu().numpy(),
               ac=alg.ac_flat.data.cpu().numpy(), est=alg.est_flat.data.cpu().numpy(), lr=alg.lr_ac,
               stats=[float(v) for v in out], arena=alg._grad_arena is not None, peer=getattr(alg, "_peer", None) is not None)
    if world > 1 and peer and arena:
        # K31 alone, many calls back to back (eager, then as a replayed CUDA graph): equals NCCL's all-reduce of the same data;
        # its norms equal the sums of squares of the reduced segments
        from qa_b200 import ops
        pa = alg._peer
        n_ac, n_est = alg.ac_flat.numel, alg.est_flat.numel
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        ws = [torch.zeros(2, device=dev, dtype=torch.float64) for _ in range(2)]
        steps = [torch.zeros(1, device=dev, dtype=torch.int32) for _ in range(2)]

        def call():
            ops.peer_allreduce(pa.world_size, pa.rank, pa.n, pa.arena_ptrs, pa.ctrl_ptrs, seg_split=n_ac, norm_end=n_ac + n_est,
                               sumsq_out=(ws[0], ws[1]), grad_scale=0.5, step_inc=(steps[0], steps[1]), scale_index=n_ac + n_est)

        worst, norm_err = 0.0, 0.0
        for it in range(12):
            x = torch.randn(pa.n, device=dev, generator=gen)
            ref = x.clone()
            dist.all_reduce(ref)
            pa.tensor.copy_(x)
            torch.cuda.synchronize()
            dist.barrier()
            if it < 6:
                call()
            else:
                if it == 6:
                    g = torch.cuda.CUDAGraph()
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        call()                                   # warm-up call on both ranks
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    pa.tensor.copy_(x)
                    torch.cuda.synchronize()
                    dist.barrier()
                    with torch.cuda.graph(g):
                        call()
                g.replay()
            torch.cuda.synchronize()
            got = pa.tensor.clone()
            kl = ref[n_ac + n_est] * 0.5
            worst = max(worst, float((got[:n_ac + n_est] - ref[:n_ac + n_est]).abs().max()), float((got[n_ac + n_est] - kl).abs()))
            for k, (lo, hi) in enumerate(((0, n_ac), (n_ac, n_ac + n_est))):
                want = float((ref[lo:hi].double() ** 2).sum()) * 0.25
                norm_err = max(norm_err, abs(float(ws[k][0]) - want) / want)
            dist.barrier()
        res["peer_worst"], res["peer_norm_err"], res["peer_steps"] = worst, norm_err, int(steps[0][0])
    q.put((rank, res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _spawn(world, arena, peer=True):
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    ps = [ctx.Process(target=_run, args=(r, world, port, q, arena, peer)) for r in range(world)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=600) for _ in range(world))
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in out.values():
        for k in ("grad_ac", "grad_est", "ac", "est"):
            r[k] = torch.from_numpy(r[k])
    return out


@pytest.mark.parametrize("arena,peer", [(True, True), (True, False), (False, False)])
def test_two_rank_step_equals_the_union_batch_step(arena, peer):
    """arena + peer: ONE K31 launch pair (all-reduce over NVLink peer memory fused with the gradient norms) per optimiser step;
    arena without peer: ONE NCCL all-reduce; neither: three NCCL all-reduces."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    union = _spawn(1, arena)[0]
    ranks = _spawn(2, arena, peer)
    assert ranks[0]["arena"] == arena and ranks[0]["peer"] == (arena and peer)
    if arena and peer:
        for r in (0, 1):
            assert ranks[r]["peer_worst"] == 0.0, ranks[r]["peer_worst"]        # two ranks: a + b in either order is the same float
            assert ranks[r]["peer_norm_err"] < 1e-6 and ranks[r]["peer_steps"] == 13
    for r in (0, 1):
        for k in ("grad_ac", "grad_est"):
            a, b = ranks[r][k].double(), union[k].double()
            rel = float((a - b).norm() / b.norm())
            assert rel < 2e-5, f"rank {r} {k}: relative L2 error {rel:.3e}"         # same TF32 products, different fp32 sum order
        assert ranks[r]["lr"] == union["lr"]
        for k in ("ac", "est"):
            d = (ranks[r][k] - union[k]).abs()
            assert float(d.max()) <= 2.5e-3 and float((d < 1e-5).float().mean()) > 0.99, (k, float(d.max()))
    assert torch.equal(ranks[0]["ac"], ranks[1]["ac"]) and torch.equal(ranks[0]["grad_ac"], ranks[1]["grad_ac"])   # replicas stay in sync
    # the logged statistics are per-rank means; their rank-mean is the union's mean
    for i in range(6):
        m = 0.5 * (ranks[0]["stats"][i] + ranks[1]["stats"][i])
        if i != 3:                                   # entropy is a function of std only
            assert abs(m - union["stats"][i]) <= 1e-4 * abs(union["stats"][i]) + 1e-6, (i, m, union["stats"][i])

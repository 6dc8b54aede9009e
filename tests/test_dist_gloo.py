"""CPU, world_size 2, gloo: the host-side rules of the env-sharded data-parallel path (qa_b200.dist) --
shard partition, flat-gradient all-reduce + 1/W scaling == gradient of the union batch, scalar KL mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qa_b200 import dist as qdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    total = 64
    g = torch.Generator().manual_seed(3)
    X, Y = torch.randn(total, 7, generator=g), torch.randn(total, 1, generator=g)
    W = torch.randn(1, 7, generator=g)
    start, n = qdist.shard_envs(total, rank, world)

    def flat_grad(x, y):
        w = W.clone().requires_grad_(True)
        ((x @ w.t() - y) ** 2).mean().backward()
        return w.grad.reshape(-1).clone()

    fg = flat_grad(X[start:start + n], Y[start:start + n])
    scale = qdist.allreduce_flat_(fg)
    kl = torch.tensor(float(rank + 1))
    qdist.allreduce_mean_scalar_(kl)
    want = flat_grad(X, Y)
    q.put((rank, start, n, scale, torch.allclose(fg * scale, want, atol=1e-6), float(kl), qdist.rank_seed(1234, rank)))
    dist.destroy_process_group()


def test_env_sharded_gradient_allreduce_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [o[1] for o in out] == [0, 32] and all(o[2] == 32 for o in out)          # disjoint, covering shards
    assert all(o[3] == 0.5 and o[4] for o in out)                                   # mean of shard grads == union grad
    assert all(abs(o[5] - 1.5) < 1e-6 for o in out)                                  # KL mean identical on all ranks
    assert [o[6] for o in out] == [1234, 1235]


def test_shard_envs_rejects_uneven_split():
    import pytest
    with pytest.raises(ValueError):
        qdist.shard_envs(10, 0, 4)
    assert qdist.shard_envs(32768, 7, 8) == (28672, 4096)
    assert qdist.world() == (0, 1)

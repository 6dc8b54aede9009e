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


# ---- TSC depth student (config "TSC-student ... 2xB200"): one gradient all-reduce per distillation step --------------------
_STUDENT_KEYS = ("actor_trunk.0.weight", "actor_c.bias")
_ENCODER_KEYS = ("base_backbone.image_compression.0.weight", "combination_mlp.2.weight", "rnn.weight_hh_l0", "output_mlp.0.bias")


def _student_grads(alg, inputs):
    """Distillation loss of update_depth_actor (ppo.py:329-337) on `inputs`, backward, rank-mean of the gradients."""
    from student_case import student_forward
    for m in alg.depth_encoder.byol_learner.augment1:
        m.p = -1.0                                                  # augmentation off: ranks would draw different noise
    cat = student_forward(alg, inputs)
    a, y, o = alg.depth_actor_losses(cat["student"], inputs["actions_teacher"], cat["yaw_s"], cat["yaw_t"], cat["obst_s"], cat["obst_t"])
    params = [*alg.depth_actor.parameters(), *alg.depth_encoder.parameters()]
    for p in params:
        p.grad = None
    (a + y + o).backward()
    qdist.allreduce_mean_grads_(params)
    an, en = dict(alg.depth_actor.named_parameters()), dict(alg.depth_encoder.named_parameters())
    return {**{k: an[k].grad.clone() for k in _STUDENT_KEYS}, **{k: en[k].grad.clone() for k in _ENCODER_KEYS}}


def _student_worker(rank, world, port, q):
    import test_tsc_student as TS
    from qa_b200 import synthetic
    from student_case import shard_inputs
    torch.set_num_threads(2)
    alg = TS.build("cpu")                                           # before init_process_group: BatchNorm1d, not SyncBatchNorm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inputs = synthetic.make_student_inputs(4, 2, 5)
    start, n = qdist.shard_envs(4, rank, world)
    grads = _student_grads(alg, shard_inputs(inputs, start, start + n))
    q.put((rank, {k: v.numpy() for k, v in grads.items()}))
    dist.destroy_process_group()


def test_student_distillation_gradients_average_to_the_union_batch_gloo():
    import numpy as np
    import test_tsc_student as TS
    from qa_b200 import synthetic
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_student_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=300) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _student_grads(TS.build("cpu"), synthetic.make_student_inputs(4, 2, 5))       # one process, all four envs
    for k, w in want.items():
        w = w.numpy()
        assert np.array_equal(out[0][k], out[1][k]), k                                   # ranks agree bit for bit
        assert float(np.abs(w).max()) > 0
        assert np.allclose(out[0][k], w, rtol=1e-4, atol=1e-6 * float(np.abs(w).max())), (k, float(np.abs(out[0][k] - w).max()))


# ---- discriminator update under env sharding: gradient all-reduce, pooled normaliser moments, prior mean ---------
def _disc_result(alg, env, norm, batches, scale):
    stats = alg.update_ss_info_gail(*batches)
    norm.sync_host()
    return dict(stats=torch.stack(stats).numpy(), grad=(alg.disc_flat.grad * scale).numpy().copy(), prior=env.prior_parameters.numpy().copy(),
                mean=norm.mean.copy(), var=norm.var.copy(), count=float(norm.count))


def _disc_shard(batches, rank, world):
    (s, e, c), (sl, lab), su = batches
    cut = lambda t: t[rank * (len(t) // world):(rank + 1) * (len(t) // world)]      # noqa: E731
    return (cut(s), cut(e), cut(c)), (cut(sl), cut(lab)), cut(su)


def _disc_worker(rank, world, port, q):
    import test_disc_batched as TD
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)        # before the algorithm: it reads the world size
    torch.set_num_threads(2)
    out = {}
    for batched in (False, True):
        alg, env, norm = TD._alg(batched, "MSELoss")
        assert alg.world_size == world
        out[batched] = _disc_result(alg, env, norm, _disc_shard(TD._batches(10), rank, world), 1.0 / world)
    q.put((rank, out))
    dist.destroy_process_group()


def test_discriminator_step_under_env_sharding_equals_the_union_batch_gloo():
    import numpy as np
    import test_disc_batched as TD
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_disc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=300) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for batched in (False, True):
        alg, env, norm = TD._alg(batched, "MSELoss")                    # one process, the union batches
        want = _disc_result(alg, env, norm, TD._batches(10), 1.0)
        a, b = out[0][batched], out[1][batched]
        for k in ("grad", "prior", "mean", "var"):
            assert np.array_equal(a[k], b[k]), (batched, k)             # ranks hold identical state
        assert a["count"] == b["count"] == want["count"]
        assert np.allclose(a["grad"], want["grad"], rtol=1e-4, atol=1e-6 * float(np.abs(want["grad"]).max())), batched
        assert np.allclose(a["prior"], want["prior"], rtol=1e-6, atol=1e-8)
        assert np.allclose(a["mean"], want["mean"], rtol=1e-5, atol=1e-6) and np.allclose(a["var"], want["var"], rtol=1e-5, atol=1e-6)
        # loss statistics are means over the rank's own rows: their rank-mean is the union's value
        assert np.allclose(0.5 * (a["stats"][:7] + b["stats"][:7]), want["stats"][:7], rtol=1e-4, atol=1e-6)


# ---- north_star: "a single NCCL allreduce of PPO gradients" -- the opt-in gradient arena ---------------------------------------
def _arena_worker(rank, world, port, q):
    import test_disc_batched as TD
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    for arena in (False, True):
        os.environ["QA_SINGLE_ALLREDUCE"] = "1" if arena else "0"          # the arena is the default with > 1 rank
        alg, env, norm = TD._alg(False, "MSELoss")
        assert (alg._grad_arena is not None) == arena
        # the peer-memory all-reduce (K31) is for NCCL groups on CUDA devices of one node; here the arena falls back to the
        # process group's all-reduce
        assert not qdist.PeerArena.available() and getattr(alg, "_peer", None) is None
        g = torch.Generator().manual_seed(100 + rank)
        alg.ac_flat.grad.copy_(torch.randn(alg.ac_flat.numel, generator=g))
        alg.est_flat.grad.copy_(torch.randn(alg.est_flat.numel, generator=g))
        if not arena:
            alg._kl = torch.zeros(())
        alg._kl.fill_(0.01 * (rank + 1))
        # every Parameter's .grad must alias the (possibly re-homed) flat gradient
        w = alg.actor_critic.actor_trunk[0].weight
        lo, n = alg.ac_flat.slices["actor_trunk.0.weight"]
        assert w.grad.data_ptr() == alg.ac_flat.grad[lo:lo + n].data_ptr()
        calls = [0]
        orig = dist.all_reduce

        def counting(*a, **k):
            calls[0] += 1
            return orig(*a, **k)
        dist.all_reduce = counting
        scale = alg._allreduce_grads()
        dist.all_reduce = orig
        out[arena] = (scale, calls[0], alg.ac_flat.grad.clone().numpy(), alg.est_flat.grad.clone().numpy(), float(alg._kl),
                      float(w.grad.abs().sum()))
    q.put((rank, out))
    dist.destroy_process_group()


def test_gradient_arena_reduces_everything_in_one_collective_gloo():
    import numpy as np
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_arena_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=300) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        (s3, c3, ga3, ge3, kl3, _), (s1, c1, ga1, ge1, kl1, wsum) = out[r][False], out[r][True]
        assert (c3, c1) == (3, 1) and s3 == s1 == 0.5                      # three collectives -> one
        assert np.array_equal(ga3, ga1) and np.array_equal(ge3, ge1) and abs(kl3 - kl1) < 1e-9 and abs(kl1 - 0.015) < 1e-8
        assert wsum > 0
    assert np.array_equal(out[0][True][2], out[1][True][2])

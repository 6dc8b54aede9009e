"""Checkpoint codec: flat Adam state <-> the reference's `torch.optim.Adam.state_dict()` layout
(bbc/rsl_rl/runners/on_policy_runner.py:306-339).  Host logic only -- plain tensor slicing, runs on CPU.

* synthetic round trip: flat moments -> torch dict -> a real `torch.optim.Adam.load_state_dict` -> flat again;
* when the reference tree is present (build container only): the six optimiser dicts of the SHIPPED checkpoint
  `tsc/weights/bbc/model.pt` load into the flat layout and come back out entry for entry."""
import copy
import ctypes
import os
import types

import pytest
import torch

from qa_b200.config import bbc_train_cfg
from qa_b200.rsl_rl import ActorCritic, Discriminator, Estimator, Normalizer, SSInfoGAIL
from qa_b200.rsl_rl import checkpoint as ckpt
from qa_b200.rsl_rl.algorithm import FlatAdam
from qa_b200.rsl_rl.utils import install_pickle_alias

SHIPPED = "/root/reference/tsc/weights/bbc/model.pt"


def _alg():
    cfg = bbc_train_cfg()
    ac = ActorCritic(101, 671, 12, 57, 10, 4, 29, 11, **cfg["policy"])
    est = Estimator(57, 4, hidden_dims=[128, 64])
    env = types.SimpleNamespace(task_obs_weight_decay=True, task_obs_weight=0.7, dim_c=5, num_obs_disc=49,
                                cfg=types.SimpleNamespace(), latent_eps=None, latent_c=None,
                                abi_args=ctypes.pointer(ctypes.c_int(0)))      # like LeggedRobot: not deep-copyable / picklable
    disc = Discriminator(env, 98, 49, 5, 0.02, "MSELoss", None, 1.0, 0.01, 0.2, 0.2, 2, 2, 0.0, [512, 256], "cpu")
    alg_cfg = dict(cfg["algorithm"], disc_replay_buffer_size=64, use_cuda_graph=False, fused_loss=False)
    return SSInfoGAIL(env, ac, disc, est, cfg["estimator"], None, Normalizer(98), 2, 2, 49, 0.0, device="cpu", **alg_cfg)


def test_flat_adam_round_trips_through_a_real_torch_optimizer():
    torch.manual_seed(0)
    alg = _alg()
    g = torch.Generator().manual_seed(1)
    a = alg.optim_ac
    a.exp_avg.copy_(torch.randn(a.exp_avg.shape, generator=g))
    a.exp_avg_sq.copy_(torch.rand(a.exp_avg_sq.shape, generator=g))
    a.step_count.fill_(37)
    a.lr.fill_(2.5e-4)
    groups = alg._optim_groups()["optim_ac"]
    sd = ckpt.to_torch_state_dict(groups)
    params = list(alg.actor_critic.parameters())
    assert sd["param_groups"][0]["params"] == list(range(len(params))) and sd["param_groups"][0]["name"] == "actor_critic"
    # a real torch optimiser over same-shaped parameters accepts the dict as is
    twin = [torch.nn.Parameter(torch.zeros(p.shape)) for p in params]
    opt = torch.optim.Adam([{"params": twin, "name": "actor_critic"}], lr=1e-3)
    opt.load_state_dict(copy.deepcopy(sd))
    assert opt.param_groups[0]["lr"] == pytest.approx(2.5e-4)
    for i, (p, (off, cnt, shape)) in enumerate(zip(twin, groups[0]["params"])):
        st = opt.state[p]
        assert float(st["step"]) == 37 and tuple(st["exp_avg"].shape) == tuple(p.shape) == shape
        if len(shape) == 2:
            want = a.exp_avg[off:off + cnt].view(shape[0], -1)[:, :shape[1]]
        else:
            want = a.exp_avg[off:off + p.numel()].view(shape)
        assert torch.equal(st["exp_avg"], want), i
    # ... and what the torch optimiser writes loads back into the flat layout (row padding stays zero)
    b = FlatAdam(alg.ac_flat, 1e-3, 1.0)
    ckpt.from_torch_state_dict(opt.state_dict(), [dict(groups[0], adam=b)])
    pad = torch.ones(alg.ac_flat.numel, dtype=torch.bool)
    for off, cnt, shape in groups[0]["params"]:
        if len(shape) == 2:
            pad[off:off + cnt].view(shape[0], -1)[:, :shape[1]] = False
        else:
            pad[off:off + int(torch.tensor(shape).prod())] = False
    live = ~pad
    assert torch.equal(b.exp_avg[live], a.exp_avg[live]) and torch.equal(b.exp_avg_sq[live], a.exp_avg_sq[live])
    assert float(b.exp_avg[pad].abs().sum()) == 0.0 and int(b.step_count) == 37 and float(b.lr) == pytest.approx(2.5e-4)


def test_six_optimizer_dicts_have_the_reference_layout_before_any_update():
    alg = _alg()
    d = alg.optimizer_state_dicts()
    assert list(d) == ["optim_ac", "optim_hist_encoder", "optim_estimator", "optim_d", "optim_q_eps", "optim_q_c"]
    assert [len(g["params"]) for g in d["optim_ac"]["param_groups"]] == [29]
    assert [len(g["params"]) for g in d["optim_hist_encoder"]["param_groups"]] == [8]
    assert [len(g["params"]) for g in d["optim_estimator"]["param_groups"]] == [6]
    for k, names in (("optim_d", ["trunk", "head"]), ("optim_q_eps", ["trunk", "encoder_eps"]), ("optim_q_c", ["trunk", "classifier"])):
        pg = d[k]["param_groups"]
        assert [g["name"] for g in pg] == names and [len(g["params"]) for g in pg] == [4, 2]
        assert all(g["weight_decay"] == 1e-3 and g["momentum"] == 0.9 for g in pg)
        assert tuple(d[k]["state"][0]["exp_avg"].shape) == (512, 98)
    # a dict loaded before the discriminator's first update is kept and handed to the optimisers when they are built
    d["optim_q_c"]["state"][5]["exp_avg"].fill_(0.25)
    alg.load_optimizer_state_dicts(d)
    assert alg.optimizer_state_dicts()["optim_q_c"]["state"][5]["exp_avg"].eq(0.25).all()
    alg._init_disc_update()
    assert alg._pending_disc_optim is None
    assert alg.optim_q_c[1].exp_avg[-alg.disc_flat.slices["classifier.bias"][1]:][:5].eq(0.25).all()
    assert alg.optimizer_state_dicts()["optim_q_c"]["state"][5]["exp_avg"].eq(0.25).all()


@pytest.mark.skipif(not os.path.exists(SHIPPED), reason="the reference tree only exists in the build container")
def test_shipped_reference_checkpoint_optimizers_load_and_re_export_entry_for_entry():
    install_pickle_alias()
    ref = torch.load(SHIPPED, map_location="cpu", weights_only=False)
    alg = _alg()
    alg.actor_critic.load_state_dict(ref["actor_critic"])
    alg.estimator.load_state_dict(ref["estimator"])
    alg.disc.load_state_dict(ref["disc"])
    alg._init_disc_update()
    alg.load_optimizer_state_dicts(ref)
    out = alg.optimizer_state_dicts()
    for k in ("optim_ac", "optim_hist_encoder", "optim_estimator", "optim_d", "optim_q_eps", "optim_q_c"):
        want, got = ref[k], out[k]
        assert [g["params"] for g in got["param_groups"]] == [g["params"] for g in want["param_groups"]], k
        for gw, gg in zip(want["param_groups"], got["param_groups"]):
            for key in ("lr", "betas", "eps", "weight_decay", "amsgrad", "name", "momentum"):
                if key in gw:
                    assert gg[key] == pytest.approx(gw[key]) if isinstance(gw[key], float) else tuple(gg[key]) == tuple(gw[key]) \
                        if isinstance(gw[key], (tuple, list)) else gg[key] == gw[key], (k, key)
        steps = {float(s["step"]) for s in want["state"].values()}
        for pid, st in got["state"].items():
            if pid in want["state"]:
                assert torch.equal(st["exp_avg"], want["state"][pid]["exp_avg"]), (k, pid)
                assert torch.equal(st["exp_avg_sq"], want["state"][pid]["exp_avg_sq"]), (k, pid)
                assert float(st["step"]) == max(steps)
            else:                                                  # never stepped in the reference (history encoder in optim_ac)
                assert float(st["exp_avg"].abs().sum()) == 0.0, (k, pid)


def test_runner_checkpoint_round_trip_keeps_the_reference_keys(tmp_path):
    """OnPolicyRunner.save / load (on_policy_runner.py:306-339) on the host path: same top-level keys, normaliser pickled under
    the reference's class path, optimiser moments restored."""
    import pickle
    import pickletools
    from qa_b200.rsl_rl.runner import OnPolicyRunner
    r = OnPolicyRunner.__new__(OnPolicyRunner)
    r.alg, r.device, r.current_learning_iteration = _alg(), "cpu", 7
    r.alg.optim_estimator.exp_avg.fill_(0.5)
    r.alg.optim_estimator.step_count.fill_(3)
    r.alg.disc_normalizer.mean[:] = 1.25
    path = str(tmp_path / "model.pt")
    r.save(path, infos={"note": 1})
    d = torch.load(path, map_location="cpu", weights_only=False)
    assert list(d) == ['actor_critic', 'estimator', 'disc', 'optim_ac', 'optim_hist_encoder', 'optim_estimator', 'optim_d',
                       'optim_q_eps', 'optim_q_c', 'disc_normalizer', 'reward_i_normalizer', 'iter', 'infos']
    ops_ = [(op.name, arg) for op, arg, _ in pickletools.genops(pickle.dumps(d["disc_normalizer"], 2))]
    assert ("GLOBAL", "rsl_rl.utils.utils Normalizer") in ops_
    r2 = OnPolicyRunner.__new__(OnPolicyRunner)
    r2.alg, r2.device, r2.current_learning_iteration = _alg(), "cpu", 0
    assert r2.load(path) == {"note": 1} and r2.current_learning_iteration == 7
    assert r2.alg.optim_estimator.exp_avg.max() == 0.5 and int(r2.alg.optim_estimator.step_count) == 3
    assert float(r2.alg.disc_normalizer.mean[0]) == 1.25
    for a, b in zip(r.alg.actor_critic.state_dict().values(), r2.alg.actor_critic.state_dict().values()):
        assert torch.equal(a, b)

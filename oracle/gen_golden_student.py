"""TEST INFRASTRUCTURE (never imported by the product).  Pins the TSC depth-student pieces against the
UNMODIFIED reference (`/root/reference/tsc`, build container only) and writes `tests/golden/tsc_student_seed5.npz`:

  * `RecurrentDepthBackbone` + `DepthOnlyFCBackbone58x87` + the BYOL learner (tsc/rsl_rl/modules/depth_backbone.py,
    modules/byol.py): state_dict keys / shapes / parameter order equal; with the same (formula-generated) weights and the same
    python-`random` / torch seeds, three recurrent forward steps with the augmentation ON agree bit-for-bit on CPU;
  * `PPO.update_depth_actor` (tsc/rsl_rl/algorithms/ppo.py:327-358) on the graph those steps built: the four returned
    statistics and the post-step parameters of the student actor, the depth encoder and the BYOL target encoder.

Run in its own process (bbc/ and tsc/ fork the same package names):   python oracle/gen_golden_student.py
"""
import copy
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ref_harness import import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N, T, SEED = 4, 3, 5
P, A, Y, L, ND, NC = 65, 8, 2, 32, 3, 6


def build_reference(ref):
    import importlib
    db = importlib.import_module("rsl_rl.modules.depth_backbone")
    policy = dict(scan_encoder_dims=[128, 64, 32], actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128],
                  priv_encoder_dims=[64], activation="elu", tanh_encoder_output=False, init_noise_std=1.0)
    ac = ref.actor_critic.ActorCriticTSC(P, A, 132, 800, 29, 4, 10, ND, NC, device="cpu", **policy)
    env_cfg = types.SimpleNamespace(env=types.SimpleNamespace(n_delta_yaw=Y, n_obst_type=A - Y, n_proprio=P))
    backbone = db.DepthOnlyFCBackbone58x87(P, L, 512)
    enc = db.RecurrentDepthBackbone(backbone, L, env_cfg)
    actor = copy.deepcopy(ac.actor)
    backbone.augment = enc.byol_learner.augment1
    est = ref.estimator.Estimator(input_dim=P - A, output_dim=4, hidden_dims=[128, 64])
    est_paras = dict(priv_states_dim=4, num_prop=P - A, num_auxiliary=A, num_scan=132, learning_rate=1e-4,
                     train_with_estimated_states=True)
    dcfg = dict(if_depth=True, learning_rate=1e-3, learning_rate_byol=3e-4, learning_rate_min=1e-5, num_steps_per_env=24, hidden_dims=512)
    ac_bbc = torch.nn.Linear(1, 1)
    alg = ref.ppo.PPO(ac, ac_bbc, est, est_paras, enc, dcfg, actor, device="cpu", max_grad_norm=1.0, learning_rate=5e-4)
    return alg


def build_ours():
    from qa_b200.config import tsc_train_cfg
    from qa_b200.rsl_rl.depth_backbone import DepthOnlyFCBackbone58x87, RecurrentDepthBackbone
    from qa_b200.rsl_rl.modules import Estimator
    from qa_b200.rsl_rl.tsc import ActorCriticTSC, PPO
    cfg = tsc_train_cfg(use_camera=True)
    ac = ActorCriticTSC(P, A, 132, 800, 29, 4, 10, ND, NC, device="cpu", **cfg["policy"])
    backbone = DepthOnlyFCBackbone58x87(P, L, 512)
    enc = RecurrentDepthBackbone(backbone, L, types.SimpleNamespace(n_delta_yaw=Y, n_obst_type=A - Y, n_proprio=P))
    actor = copy.deepcopy(ac.actor)
    backbone.augment = enc.byol_learner.augment1
    est = Estimator(input_dim=P - A, output_dim=4, hidden_dims=[128, 64])
    est_paras = dict(priv_states_dim=4, num_prop=P - A, num_auxiliary=A, num_scan=132, learning_rate=1e-4,
                     train_with_estimated_states=True)
    return PPO(ac, None, est, est_paras, enc, cfg["depth_encoder"], actor, device="cpu", max_grad_norm=1.0, learning_rate=5e-4,
               use_cuda_graph=False, fused_loss=False)


def run_case(alg, inputs, aug_p):
    """Three student steps (depth encoder with augmentation, student actor) + one update_depth_actor; same code for both
    implementations -- only the module objects differ."""
    from student_case import student_rollout_and_update
    return student_rollout_and_update(alg, inputs, aug_p, SEED)


def dagger_case(ref, theirs, n=64):
    """`PPO.update_dagger` (tsc/rsl_rl/algorithms/ppo.py:284-314): two epochs over one 64-row minibatch; the oracle restatement
    (`oracle/tsc_trainer.py update_dagger`) is pinned against the reference here, the CUDA path against this fixture on the GPU."""
    import tsc_trainer as OT
    from qa_b200 import synthetic
    g = torch.Generator().manual_seed(SEED + 7)
    obs = 0.5 * torch.randn(n, 800, generator=g)
    alg = theirs
    synthetic.load_student_weights(alg.actor_critic, SEED + 2)
    with torch.no_grad():
        alg.actor_critic.std.fill_(1.0)
    sd0 = {k: v.clone() for k, v in alg.actor_critic.state_dict().items()}
    alg.num_learning_epochs, alg.num_mini_batches = 2, 1
    alg.init_storage(n, 1, [800], [None], [19])
    alg.storage.observations[0].copy_(obs)
    torch.manual_seed(SEED)
    lw = alg.update_dagger()
    sd = alg.actor_critic.state_dict()
    lg, enc = OT.update_dagger(sd0, obs, lr=5e-4, epochs=2)
    assert abs(lw - lg) <= 1e-6 * abs(lw), (lw, lg)
    err = max(float((sd[k] - v).abs().max()) for k, v in enc.items())
    assert err <= 1e-6, err
    assert all(torch.equal(sd[k], sd0[k]) for k in sd if k not in enc) and alg.counter == 1
    ew = torch.cat([sd[k].reshape(-1) for k in sd if k.startswith("actor.history_encoder.")]).clone()
    print(f"  update_dagger x2: oracle == reference (mean loss {lw:.6f}, max param err {err:.1e})")
    return {"dagger.obs_seed_n": np.array([SEED + 7, n]), "dagger.mean_loss": np.array(lw), "dagger.encoder_params": ew.numpy()}


def main():
    from qa_b200 import synthetic
    ref = import_reference("tsc")
    torch.manual_seed(0)
    random.seed(0)
    theirs = build_reference(ref)
    mine = build_ours()
    # same keys, shapes and parameter order
    for name in ("depth_encoder", "depth_actor"):
        a, b = getattr(theirs, name), getattr(mine, name)
        ka, kb = list(a.state_dict().keys()), list(b.state_dict().keys())
        assert ka == kb, (name, [k for k in ka if k not in kb], [k for k in kb if k not in ka])
        assert [tuple(v.shape) for v in a.state_dict().values()] == [tuple(v.shape) for v in b.state_dict().values()]
        assert [tuple(p.shape) for p in a.parameters()] == [tuple(p.shape) for p in b.parameters()], name
        assert [p.requires_grad for p in a.parameters()] == [p.requires_grad for p in b.parameters()], name
        print(f"  {name}: {len(ka)} state_dict keys, {sum(1 for _ in a.parameters())} parameters: names / shapes / order equal")
    for alg in (theirs, mine):
        synthetic.load_student_weights(alg.depth_encoder, SEED)
        synthetic.load_student_weights(alg.depth_actor, SEED + 1)
    inputs = synthetic.make_student_inputs(N, T, SEED)
    aug_p = 0.6
    from student_case import byol_forward_backward
    bw, bg = byol_forward_backward(theirs, inputs, aug_p, SEED), byol_forward_backward(mine, inputs, aug_p, SEED)
    for k in bw:
        err = float((bw[k] - bg[k]).abs().max())
        assert err <= 1e-7 * max(1.0, float(bw[k].abs().max())), (k, err)
        print(f"  {k:62s} max|ref - ours| = {err:.1e}  (max |ref| {float(bw[k].abs().max()):.2e})")
    want = run_case(theirs, inputs, aug_p)
    got = run_case(mine, inputs, aug_p)
    assert want["n_aug_applied"] == got["n_aug_applied"] and want["n_aug_applied"] >= 6, want["n_aug_applied"]
    for k in want:
        if k == "n_aug_applied":
            continue
        w, g = torch.as_tensor(want[k]), torch.as_tensor(got[k])
        assert w.shape == g.shape, k
        err = float((w.double() - g.double()).abs().max())
        assert err <= 1e-6 * max(1.0, float(w.abs().max())), (k, err)
        print(f"  {k:28s} max|ref - ours| = {err:.2e}")
    out = {f"want.{k}": np.asarray(v) for k, v in want.items()}
    out.update(dagger_case(ref, theirs))
    out.update({k: v.numpy() for k, v in bw.items()})
    out["meta.N_T_seed_augp"] = np.array([N, T, SEED, aug_p], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "tsc_student_seed5.npz"), **out)
    print("wrote tests/golden/tsc_student_seed5.npz", os.path.getsize(os.path.join(GOLD, "tsc_student_seed5.npz")), "bytes")


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE (build container only): the drop-in boundary of SURVEY 8(b), checked from the REFERENCE's side.

The unmodified `SSInfoGAIL.act` / `update_actor_critic` (bbc/rsl_rl/algorithms/gail.py:176-197, 328-413) are run twice on the
same inputs -- once over the reference's own `ActorCritic` / `Estimator`, once over THIS package's classes constructed with the
same arguments and loaded from the same `state_dict` -- and must produce the same actions, values, log-probs, the same six loss
statistics, the same adapted learning rate and the same post-step parameters.  I.e. a maintainer can hand
`qa_b200.rsl_rl.ActorCritic` to the reference trainer unchanged.  (CPU, fp32 mode; exits non-zero on any mismatch.)

The same for the TSC fork: `PPO.act` / `PPO.update` (tsc/rsl_rl/algorithms/ppo.py:101-262) over `qa_b200.rsl_rl.ActorCriticTSC`.

Beside the trainer runs, class by class on the same inputs: the rollout storages (Transition / add_transitions, the mini-batch
generators under one seed, get_statistics), the replay-buffer ring across its wrap-around, `StateHistoryEncoder` for tsteps
10 / 20 / 50, the fork's `ActorCriticBBC` and `Discriminator` (all three style-reward mappings), and the env's
`set_commands` against the reference method on injected state -- all bit-equal or within 1e-6.

  python oracle/check_interop.py          # bbc
  python oracle/check_interop.py tsc      # tsc (own process: the two forks share package names)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

from gen_golden_policy import build_reference_nets, make_alg  # noqa: E402
from qa_b200 import synthetic  # noqa: E402
from qa_b200.config import bbc_train_cfg  # noqa: E402
from ref_harness import import_reference  # noqa: E402


def post_step_equal(want, got, lr_ac, lr_est):
    """Post-Adam parameters: equal entry for entry except where a ~0 gradient lets a last-bit difference of the two module
    implementations flip Adam's sign-like first step (bounded by the learning rate; at most 0.5 % of the entries)."""
    for part, lr in (("ac", lr_ac), ("est", lr_est)):
        a = torch.cat([v.reshape(-1) for v in want[part].values()])
        b = torch.cat([v.reshape(-1) for v in got[part].values()])
        d = (a - b).abs()
        tol = 1e-6 + 1e-5 * a.abs()
        assert float(d.max()) <= 2.5 * lr and float((d > tol).float().mean()) < 5e-3, (part, float(d.max()), float((d > tol).float().mean()))


def run(ref, ac, est, obs, draw, batch_noise):
    alg = make_alg(ref, ac, est)
    Normal = torch.distributions.Normal
    orig, orig_randn_like = Normal.sample, torch.randn_like
    out = {}
    try:
        # the same N(0,1) draw for both implementations: the reference samples through Normal.sample, this package through
        # mean + std * randn_like(mean)
        Normal.sample = lambda self, sample_shape=torch.Size(): (self.loc + self.scale * draw).detach()
        torch.randn_like = lambda t, **k: draw.clone() if t.shape == draw.shape else orig_randn_like(t, **k)
        for he in (False, True):
            with torch.inference_mode():
                a = alg.act(obs.clone(), obs.clone(), hist_encoding=he)
            tr = alg.transition
            out[f"act{int(he)}"] = [t.clone() for t in (a, tr.values, tr.actions_log_prob, tr.action_mean, tr.action_sigma)]
        a0, v0, lp0, mu0, sg0 = out["act0"]
        n1, n2, n3, n4 = batch_noise
        sample = (obs, obs, a0, v0, n1, v0 + 0.3 * n2, (lp0 + 0.05 * n3).unsqueeze(1), mu0 + 0.05 * n4, sg0 * 1.05, (None, None), None)
        Normal.sample = lambda self, sample_shape=torch.Size(): self.loc.detach()
        losses = alg.update_actor_critic(sample)
    finally:
        Normal.sample, torch.randn_like = orig, orig_randn_like
    out["losses"] = [torch.as_tensor(x).float().mean().detach() for x in losses]
    out["lr"] = alg.lr_ac
    out["ac"] = {k: v.detach().clone() for k, v in ac.state_dict().items()}
    out["est"] = {k: v.detach().clone() for k, v in est.state_dict().items()}
    return out


def main():
    ref = import_reference("bbc")
    from qa_b200.rsl_rl import ActorCritic, Estimator
    from qa_b200.rsl_rl import linear
    linear.set_mode("fp32")
    torch.set_num_threads(1)
    w = synthetic.make_weights(3)
    z = np.load(os.path.join(ROOT, "tests", "golden", "bbc_env_n64a.npz"))
    obs = torch.from_numpy(z["ref.obs_buf"]).clone()
    N = obs.shape[0]
    g = torch.Generator().manual_seed(5)
    draw = torch.randn(N, 12, generator=g)
    noise = (torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g), torch.randn(N, generator=g), torch.randn(N, 12, generator=g))
    ac_r, est_r, _, _, _ = build_reference_nets(ref, w)
    cfg = bbc_train_cfg()
    ac_o = ActorCritic(101, 671, 12, 57, 10, 4, 29, 11, **cfg["policy"])
    ac_o.load_state_dict(w["ac"])
    est_o = Estimator(input_dim=57, output_dim=4, hidden_dims=[128, 64])
    est_o.load_state_dict(w["est"])
    assert list(ac_o.state_dict()) == list(ac_r.state_dict()) and list(est_o.state_dict()) == list(est_r.state_dict())
    assert [tuple(p.shape) for p in ac_o.parameters()] == [tuple(p.shape) for p in ac_r.parameters()]
    want, got = run(ref, ac_r, est_r, obs, draw, noise), run(ref, ac_o, est_o, obs, draw, noise)
    worst = 0.0
    for he in ("act0", "act1"):
        for a, b in zip(want[he], got[he]):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (he, float((a - b).abs().max()))
            worst = max(worst, float((a - b).abs().max()))
    for a, b in zip(want["losses"], got["losses"]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (float(a), float(b))
    assert abs(want["lr"] - got["lr"]) < 1e-12
    post_step_equal(want, got, lr_ac=1e-3, lr_est=1e-4)
    # storage: the reference's 11-tuple generator over this package's RolloutStorage (same permutation under the same seed)
    from qa_b200.rsl_rl import RolloutStorage
    gs = torch.Generator().manual_seed(11)
    sr, so = ref.RolloutStorage(8, 6, [671], [671], [12], device="cpu"), RolloutStorage(8, 6, [671], [671], [12], device="cpu")
    for name in ("observations", "privileged_observations", "actions", "values", "advantages", "returns", "actions_log_prob", "mu", "sigma"):
        v = torch.randn(getattr(sr, name).shape, generator=gs)
        getattr(sr, name).copy_(v)
        getattr(so, name).copy_(v)
    torch.manual_seed(3)
    a = list(sr.mini_batch_generator(3, 2))
    torch.manual_seed(3)
    b = list(so.mini_batch_generator(3, 2))
    assert len(a) == len(b) == 6
    for ta, tb in zip(a, b):
        assert len(ta) == len(tb) == 11 and ta[9] == tb[9] == (None, None) and ta[10] is None and tb[10] is None
        assert all(torch.equal(x, y) for x, y in zip(ta[:9], tb[:9]))
    sr2, so2 = ref.RolloutStorage(8, 3, [671], [671], [12], device="cpu"), RolloutStorage(8, 3, [671], [671], [12], device="cpu")
    for step in range(3):
        fields = dict(observations=torch.randn(8, 671, generator=gs), critic_observations=torch.randn(8, 671, generator=gs),
                      actions=torch.randn(8, 12, generator=gs), rewards=torch.randn(8, generator=gs),
                      dones=(torch.rand(8, generator=gs) < 0.3), values=torch.randn(8, 1, generator=gs),
                      actions_log_prob=torch.randn(8, generator=gs), action_mean=torch.randn(8, 12, generator=gs),
                      action_sigma=torch.rand(8, 12, generator=gs))
        for st_, T_ in ((sr2, ref.RolloutStorage.Transition), (so2, RolloutStorage.Transition)):
            tr_ = T_(8, [671], [671], [12], "cpu")                   # the BBC fork's Transition pre-allocates (rollout_storage.py:8-22)
            for k_, v_ in fields.items():
                setattr(tr_, k_, v_)
            st_.add_transitions(tr_)
    for name in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "actions_log_prob", "mu", "sigma"):
        assert torch.equal(getattr(sr2, name).float(), getattr(so2, name).float()), name
    assert sr2.step == so2.step == 3
    # StateHistoryEncoder for every history length the reference defines (actor_critic.py:9-59: tsteps 10 / 20 / 50)
    from qa_b200.rsl_rl import StateHistoryEncoder
    for tsteps in (10, 20, 50):
        er = ref.actor_critic.StateHistoryEncoder(torch.nn.ELU(), 57, tsteps, 29)
        eo = StateHistoryEncoder(torch.nn.ELU(), 57, tsteps, 29)
        assert [(k, tuple(v.shape)) for k, v in er.state_dict().items()] == [(k, tuple(v.shape)) for k, v in eo.state_dict().items()], tsteps
        eo.load_state_dict(er.state_dict())
        xh = torch.randn(5, tsteps, 57, generator=gs)
        with torch.no_grad():
            assert torch.allclose(er(xh), eo(xh), rtol=1e-6, atol=1e-7), tsteps
    # replay buffer ring (storage/replay_buffer.py:23-42): inserts across the wrap-around
    import importlib
    from qa_b200.rsl_rl.algorithm import ReplayBuffer
    rr_, ro_ = importlib.import_module("rsl_rl.storage.replay_buffer").ReplayBuffer(49, 5, 2, 50, "cpu"), ReplayBuffer(49, 5, 2, 50, "cpu")
    for n_ins in (20, 20, 25, 50, 7):
        st_, e_, c_ = torch.randn(n_ins, 98, generator=gs), torch.rand(n_ins, 1, generator=gs), torch.rand(n_ins, 5, generator=gs)
        rr_.insert(st_, e_, c_)
        ro_.insert(st_, e_, c_)
        assert (rr_.step, rr_.num_samples) == (ro_.step, ro_.num_samples)
        assert torch.equal(rr_.states, ro_.states) and torch.equal(rr_.latent_eps, ro_.latent_eps) and torch.equal(rr_.latent_c, ro_.latent_c)
    print(f"interop OK: reference SSInfoGAIL.act / update_actor_critic over qa_b200 ActorCritic + Estimator == over the reference's "
          f"(max |diff| of act outputs {worst:.1e}; losses {[round(float(x), 6) for x in got['losses']]}; lr {got['lr']:.6g})")


def run_tsc(ref, OT, ac, est, obs, draw, mode_u, noise):
    """tsc/rsl_rl/algorithms/ppo.py: PPO.act (:101-125) and one PPO.update (:159-262) over the given modules."""
    N = obs.shape[0]
    P = ref.ppo.PPO
    alg = P.__new__(P)
    alg.device, alg.actor_critic, alg.estimator = "cpu", ac, est
    alg.train_with_estimated_states, alg.num_prop, alg.num_auxiliary, alg.num_scan, alg.priv_states_dim = True, 57, 8, 132, 4
    alg.transition = ref.RolloutStorage.Transition()
    alg.optimizer = torch.optim.Adam(ac.parameters(), lr=5e-4)
    alg.estimator_optimizer = torch.optim.Adam(est.parameters(), lr=1e-4)
    alg.learning_rate, alg.desired_kl, alg.schedule = 5e-4, 0.01, "adaptive"
    alg.clip_param, alg.use_clipped_value_loss, alg.value_loss_coef, alg.entropy_coef = 0.2, True, 1.0, 0.01
    alg.max_grad_norm, alg.num_learning_epochs, alg.num_mini_batches = 1.0, 1, 1
    alg.priv_reg_coef_schedual, alg.counter = [0, 0.1, 500, 1000], 800
    alg.gamma, alg.lam = 0.99, 0.95
    alg.update_counter = lambda: None
    Normal, Categorical = torch.distributions.Normal, torch.distributions.Categorical
    saved = (Normal.sample, Categorical.sample, torch.randn_like, torch.rand)
    out = {}
    try:
        Normal.sample = lambda self, sample_shape=torch.Size(): (self.loc + self.scale * draw).detach()
        Categorical.sample = lambda self, sample_shape=torch.Size(): OT.sample_mode(self.probs, mode_u)
        torch.randn_like = lambda t, **k: draw.clone() if t.shape == draw.shape else saved[2](t, **k)
        torch.rand = lambda *a, **k: mode_u.clone() if a == (N,) else saved[3](*a, **k)
        for he in (False, True):
            with torch.no_grad():
                a = alg.act(obs.clone(), obs.clone(), None, hist_encoding=he)
            tr = alg.transition
            out[f"act{int(he)}"] = [t.clone() for t in (a, tr.values, tr.actions_log_prob_d, tr.actions_log_prob_c, tr.action_mean,
                                                         tr.action_sigma)]
        a0, v0, lpd, lpc, mu0, sg0 = out["act0"]
        n1, n2, n3, n4, n5 = noise
        st = ref.RolloutStorage(N, 1, [800], [None], [19], device="cpu")
        st.observations[0], st.actions[0], st.values[0] = obs, a0, v0
        st.advantages[0], st.returns[0] = n1, v0 + 0.3 * n2
        st.actions_log_prob_d[0], st.actions_log_prob_c[0] = (lpd.reshape(N) + 0.05 * n3).unsqueeze(1), (lpc.reshape(N) + 0.05 * n4).unsqueeze(1)
        st.mu[0], st.sigma[0] = mu0 + 0.05 * n5, sg0 * 1.05
        alg.storage = st
        out["update"] = [float(x) for x in alg.update()]
    finally:
        Normal.sample, Categorical.sample, torch.randn_like, torch.rand = saved
    out["lr"] = alg.learning_rate
    out["ac"] = {k: v.detach().clone() for k, v in ac.state_dict().items()}
    out["est"] = {k: v.detach().clone() for k, v in est.state_dict().items()}
    return out


def main_tsc():
    ref = import_reference("tsc")
    import tsc_trainer as OT
    from qa_b200.rsl_rl import ActorCriticTSC, Estimator
    from qa_b200.rsl_rl import linear
    linear.set_mode("fp32")
    torch.set_num_threads(1)
    w = synthetic.make_tsc_weights(3)
    policy = dict(scan_encoder_dims=[128, 64, 32], actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128],
                  priv_encoder_dims=[64], activation="elu", init_noise_std=1.0, tanh_encoder_output=False)
    N = 64
    g = torch.Generator().manual_seed(3)
    obs = torch.randn(N, 800, generator=g) * 0.5
    obs[:, 65:197] = torch.clip(obs[:, 65:197], -1, 1)
    draw, mode_u = torch.randn(N, 18, generator=g), torch.rand(N, generator=g)
    noise = (torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g), torch.randn(N, generator=g), torch.randn(N, generator=g),
             torch.randn(N, 18, generator=g))
    res = []
    for AC, E in ((ref.actor_critic.ActorCriticTSC, ref.estimator.Estimator), (ActorCriticTSC, Estimator)):
        ac = AC(65, 8, 132, 800, 29, 4, 10, 3, 6, device="cpu", **policy)
        ac.load_state_dict(w["ac"])
        est = E(input_dim=57, output_dim=4, hidden_dims=[128, 64])
        est.load_state_dict(w["est"])
        res.append(run_tsc(ref, OT, ac, est, obs, draw, mode_u, noise))
    want, got = res
    worst = 0.0
    for he in ("act0", "act1"):
        for a, b in zip(want[he], got[he]):
            assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-5, atol=1e-6), (he, float((a - b).abs().max()))
            worst = max(worst, float((a - b).abs().max()))
    for a, b in zip(want["update"], got["update"]):
        assert abs(a - b) <= 1e-5 * abs(a) + 1e-7, (want["update"], got["update"])
    assert abs(want["lr"] - got["lr"]) < 1e-12
    post_step_equal(want, got, lr_ac=5e-4, lr_est=1e-4)
    # storage: the reference's 12-tuple generator and get_statistics over this package's RolloutStorageTSC
    from qa_b200.rsl_rl import RolloutStorageTSC
    gs = torch.Generator().manual_seed(11)
    sr, so = ref.RolloutStorage(8, 6, [800], [None], [19], device="cpu"), RolloutStorageTSC(8, 6, [800], [None], [19], device="cpu")
    for name in ("observations", "actions", "values", "advantages", "returns", "actions_log_prob_d", "actions_log_prob_c", "mu", "sigma",
                 "rewards"):
        v = torch.randn(getattr(sr, name).shape, generator=gs)
        getattr(sr, name).copy_(v)
        getattr(so, name).copy_(v)
    d = (torch.rand(6, 8, 1, generator=gs) < 0.2).to(torch.uint8)
    sr.dones.copy_(d)
    so.dones.copy_(d)
    torch.manual_seed(3)
    a = list(sr.mini_batch_generator(3, 2))
    torch.manual_seed(3)
    b = list(so.mini_batch_generator(3, 2))
    assert len(a) == len(b) == 6
    for ta, tb in zip(a, b):
        assert len(ta) == len(tb) == 12 and ta[10] == tb[10] == (None, None) and ta[11] is None and tb[11] is None
        assert all(torch.equal(x, y) for x, y in zip(ta[:10], tb[:10]))
    # add_transitions: the same Transition fields land in the same buffers
    sr2, so2 = ref.RolloutStorage(8, 3, [800], [None], [19], device="cpu"), RolloutStorageTSC(8, 3, [800], [None], [19], device="cpu")
    for step in range(3):
        fields = dict(observations=torch.randn(8, 800, generator=gs), critic_observations=None, actions=torch.randn(8, 19, generator=gs),
                      rewards=torch.randn(8, generator=gs), dones=(torch.rand(8, generator=gs) < 0.3), values=torch.randn(8, 1, generator=gs),
                      actions_log_prob_d=torch.randn(8, generator=gs), actions_log_prob_c=torch.randn(8, generator=gs),
                      action_mean=torch.randn(8, 18, generator=gs), action_sigma=torch.rand(8, 18, generator=gs))
        for st_, T_ in ((sr2, ref.RolloutStorage.Transition), (so2, RolloutStorageTSC.Transition)):
            tr_ = T_()
            for k_, v_ in fields.items():
                setattr(tr_, k_, v_)
            st_.add_transitions(tr_)
    for name in ("observations", "actions", "rewards", "dones", "values", "actions_log_prob_d", "actions_log_prob_c", "mu", "sigma"):
        assert torch.equal(getattr(sr2, name).float(), getattr(so2, name).float()), name
    assert sr2.step == so2.step == 3
    (la, ra), (lb, rb) = sr.get_statistics(), so.get_statistics()
    assert torch.equal(la, lb) and torch.equal(ra, rb)
    # env.set_commands (tsc/legged_gym/envs/base/legged_robot.py:699-760): the reference method and this package's, both
    # driven unbound over the same injected state and the same multiplicative-noise draw
    import types
    from qa_b200.legged_robot_tsc import LeggedRobotTSC, TscEnvConfig
    ep = torch.randint(0, 3, (N,), generator=g)
    acts = torch.cat([torch.randint(0, 3, (N, 1), generator=g).float(), 1.4 * torch.randn(N, 18, generator=g)], dim=1)
    u = torch.rand(N, 5, generator=g)
    for resampling_time in (0.02, 0.04):                             # every step (the shipped value) / every other step (masked)
        cfg_o = TscEnvConfig(num_envs=N, resampling_time=resampling_time)
        state = lambda: dict(episode_length_buf=ep.clone(), latent_c=torch.zeros(N, 5), latent_eps=torch.zeros(N, 1),   # noqa: E731
                             commands=0.3 * torch.ones(N, 5), dim_c=5, device="cpu", mocap_indices=torch.tensor([2, 3, 4]))
        so = types.SimpleNamespace(cfg=cfg_o, **state())
        cmd_o = LeggedRobotTSC.set_commands(so, acts, action_noise_u=u)
        rcfg = types.SimpleNamespace(commands=types.SimpleNamespace(resampling_time=cfg_o.resampling_time),
                                     domain_rand=types.SimpleNamespace(randomize_action=True, action_noise=list(cfg_o.action_noise)))
        sr_ = types.SimpleNamespace(cfg=rcfg, dt=cfg_o.dt, num_actions_c=6, command_ranges=cfg_o.command_ranges, **state())
        saved_rand = ref.legged_robot.torch_rand_float
        ref.legged_robot.torch_rand_float = lambda lo, hi, shape, device=None: (hi - lo) * u + lo
        try:
            want_cmd = ref.LeggedRobot.set_commands(sr_, acts)
        finally:
            ref.legged_robot.torch_rand_float = saved_rand
        assert torch.equal(cmd_o, want_cmd) and torch.equal(so.commands, sr_.commands) and torch.equal(so.latent_c, sr_.latent_c)
        assert torch.equal(so.latent_eps, sr_.latent_eps) and float(cmd_o.abs().sum()) > 0
        due = (ep % int(resampling_time / cfg_o.dt) == 0)
        assert bool(due.any()) and (resampling_time == 0.02 or not bool(due.all()))
    # the frozen low-level controller: the fork's ActorCriticBBC over a BBC checkpoint's weights
    from qa_b200.rsl_rl import ActorCriticBBC
    wb = synthetic.make_weights(3)
    bbc_policy = dict(actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128], priv_encoder_dims=[64], activation="elu")
    br = ref.actor_critic.ActorCriticBBC(101, 671, 12, 65, 8, 10, 4, 29, 11, **bbc_policy)
    bo = ActorCriticBBC(101, 671, 12, 65, 8, 10, 4, 29, 11, **bbc_policy)
    br.load_state_dict(wb["ac"])
    bo.load_state_dict(wb["ac"])
    assert list(br.state_dict()) == list(bo.state_dict()) and br.train_with_estimated_latent and bo.train_with_estimated_latent
    obs_bbc = 0.5 * torch.randn(N, 671, generator=g)
    with torch.no_grad():
        for he in (True, False):
            a, b = br.act_inference(obs_bbc, hist_encoding=he), bo.act_inference(obs_bbc, hist_encoding=he)
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-6), (he, float((a - b).abs().max()))
        assert not torch.allclose(bo.act_inference(obs_bbc, hist_encoding=True), bo.act_inference(obs_bbc, hist_encoding=False))
    # discriminator: the TSC fork's constructor / predict_disc_reward over the same weights and normaliser
    from qa_b200.rsl_rl import DiscriminatorTSC, Normalizer
    wd = synthetic.make_weights(3)
    nr, no = ref.utils.Normalizer(98), Normalizer(98)
    for nn_ in (nr, no):
        nn_.mean, nn_.var = wd["norm_mean"].numpy().copy(), wd["norm_var"].numpy().copy()
    args = (98, 49, 5, 0.02, "MSELoss", None, 0.05, 0.3, 0.2, 2.0, 2, [512, 256])
    dr_, do_ = ref.discriminator.Discriminator(*args, nr, "cpu"), DiscriminatorTSC(*args, no, "cpu")
    dr_.load_state_dict(wd["disc"])
    do_.load_state_dict(wd["disc"])
    obs_b = torch.randn(N, 671, generator=g)
    hist = 0.5 * torch.randn(N, 2, 49, generator=g)
    rew_t = torch.rand(N, 1, generator=g)
    for a, b in zip(dr_.predict_disc_reward(rew_t, obs_b, hist), do_.predict_disc_reward(rew_t, obs_b, hist)):
        assert a.dtype == b.dtype and torch.allclose(a, b, rtol=1e-6, atol=1e-7), float((a - b).abs().max())
    for loss_fn in ("BCEWithLogitsLoss", "WassersteinLoss"):                       # the other two style-reward mappings
        ri_r, ri_o = ref.utils.Normalizer(1), Normalizer(1)
        args2 = (98, 49, 5, 0.02, loss_fn, None, 0.05, 0.3, 0.2, 2.0, 2, [512, 256])
        d1, d2 = ref.discriminator.Discriminator(*args2, nr, "cpu"), DiscriminatorTSC(*args2, no, "cpu")
        d1.reward_i_normalizer, d2.reward_i_normalizer = ri_r, ri_o
        d1.load_state_dict(wd["disc"])
        d2.load_state_dict(wd["disc"])
        for _ in range(2):                                                          # second call sees the updated normaliser
            for a, b in zip(d1.predict_disc_reward(rew_t, obs_b, hist), d2.predict_disc_reward(rew_t, obs_b, hist)):
                assert a.dtype == b.dtype and torch.allclose(a, b, rtol=1e-5, atol=1e-6), (loss_fn, float((a - b).abs().max()))
    print(f"interop OK (tsc): reference PPO.act / PPO.update over qa_b200 ActorCriticTSC + Estimator == over the reference's "
          f"(max |diff| of act outputs {worst:.1e}; update() = {[round(x, 6) for x in got['update']]}; lr {got['lr']:.6g})")


if __name__ == "__main__":
    main_tsc() if sys.argv[1:] == ["tsc"] else main()

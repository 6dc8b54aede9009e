"""Host logic of `OnPolicyRunner.learn` (bbc/rsl_rl/runners/on_policy_runner.py:118-233) on CPU: the rollout loop over a stand-in
env (the reference-shaped non-fused step), device-staged episode bookkeeping, the TensorBoard tags, `save` every iteration and
`load` into a second runner.  The CUDA-only pieces (GAE K5, minibatch gather K6, clip+Adam K8) are stubbed out; the device run
of the same loop is tests/test_zz_runner_gpu.py."""
import ctypes
import types

import numpy as np
import torch

from qa_b200 import config as K
from qa_b200.config import bbc_train_cfg
from qa_b200.rsl_rl.runner import OnPolicyRunner


class FakeEnv:
    """What the runner and SSInfoGAIL touch of `LeggedRobot`, with random dynamics."""

    def __init__(self, n=32, seed=0):
        self.num_envs, self.num_obs, self.num_privileged_obs, self.num_actions, self.num_obs_disc = n, 101, 101, 12, 49
        self.dt, self.dim_c, self.mocap_category = 0.02, 5, ["walk", "pace", "trot", "canter", "jump"]
        self.reward_names = list(K.REWARD_NAMES)
        self.reward_scales = {k: 0.5 for k in self.reward_names}
        self.task_obs_weight_decay, self.task_obs_weight, self.task_obs_weight_decay_steps = True, 1.0, 50000
        self.cfg = types.SimpleNamespace(send_timeouts=True)
        self.abi_args = ctypes.pointer(ctypes.c_int(0))            # like the real env: neither deep-copyable nor picklable
        self.g = torch.Generator().manual_seed(seed)
        self.dof_pos_limits = torch.stack([-torch.ones(12), torch.ones(12)], dim=1)
        self.episode_length_buf, self.max_episode_length = torch.zeros(n, dtype=torch.int64), 1000.0
        self.latent_eps, self.latent_c = torch.zeros(n, 1), torch.nn.functional.one_hot(torch.arange(n) % 5, 5).float()
        self._time_outs_latched = torch.zeros(n, dtype=torch.bool)
        self.default_dof_pos = torch.tensor([0.0, 0.9, -1.8] * 4).unsqueeze(0)
        self.obs_scales = types.SimpleNamespace(lin_vel=0.5, ang_vel=0.25, dof_pos=1.0, dof_vel=0.05, key_pos=1.0, foot_contact=1.0,
                                                lin_vel_dist=0.5, ang_vel_dist=0.25)
        self._episode_rew_means = torch.zeros(len(self.reward_names))
        self.prior_parameters = torch.full((5,), 0.2)
        self.prior_prob = self.prior_parameters.clone()
        self._draw()

    def _draw(self):
        n, g = self.num_envs, self.g
        self.obs = 0.3 * torch.randn(n, 671, generator=g)
        self.disc = 0.3 * torch.randn(n, 49, generator=g)
        self.rew = 0.05 * torch.rand(n, generator=g)
        self.reset_buf = torch.rand(n, generator=g) < 0.2

    def reset(self):
        return self.obs, self.obs

    def get_observations(self):
        return self.obs

    def get_privileged_observations(self):
        return self.obs

    def get_disc_observations(self):
        return self.disc

    def step_device(self, actions):
        assert actions.shape == (self.num_envs, 12)
        self._draw()
        if bool(self.reset_buf.any()):
            self._episode_rew_means = torch.rand(len(self.reward_names), generator=self.g)
            self._time_outs_latched = self.reset_buf & (torch.rand(self.num_envs, generator=self.g) < 0.5)
        return self.obs, self.obs, self.rew, self.reset_buf, None, None, None


def _runner(tmp, seed):
    torch.manual_seed(seed)
    cfg = bbc_train_cfg()
    cfg["runner"]["num_steps_per_env"], cfg["runner"]["save_interval"] = 6, 1
    cfg["algorithm"].update(disc_replay_buffer_size=512, use_cuda_graph=False, fused_loss=False)
    r = OnPolicyRunner(FakeEnv(seed=seed), cfg, log_dir=tmp, device="cpu")
    alg = r.alg

    def update(expert=None):                                        # K6 / K8 are CUDA-only: keep the bookkeeping side effects
        alg.storage.clear()
        alg.optim_ac.exp_avg.add_(0.25)
        alg.optim_ac.step_count.add_(20)
        return (0.1, 0.2, 0.0, 0.3, 0.4, 0.5)

    alg.update, alg.update_dagger, alg.compute_returns = update, (lambda: 0.75), (lambda critic_obs: None)
    return r


def test_learn_loop_books_logs_saves_and_loads_on_host(tmp_path):
    r = _runner(str(tmp_path / "a"), 1)
    r.learn(3)
    assert r.current_learning_iteration == 3 and r.alg.storage.step == 0
    sc = r.writer.scalars
    for tag in ("Loss/surrogate_loss", "Loss/estimator_loss", "Loss/hist_latent_loss", "Loss/mean_noise_std", "LR/lr_ac", "LR/lr_disc",
                "Perf/total_fps", "Episode/rew_torques", "Train/mean_reward", "Train/mean_reward_ss", "Train/mean_episode_length"):
        assert [s for s, _ in sc[tag]] == [0, 1, 2] and all(np.isfinite(v) for _, v in sc[tag]), tag
    assert "Loss/ss_loss" not in sc                                  # no expert set: the discriminator was not updated
    assert sc["Loss/hist_latent_loss"][-1][1] == 0.75 and sc["Loss/value_loss"][0][1] == 0.2
    assert r.book.len_buffer and max(r.book.len_buffer) <= 18
    # the bookkeeping saw exactly the rewards that went into the storage (before the time-out bootstrap), per finished episode
    assert abs(r.env.task_obs_weight - (1.0 - 3 / 50000)) < 1e-9
    d = torch.load(str(tmp_path / "a" / "model.pt"), map_location="cpu", weights_only=False)
    assert list(d) == ['actor_critic', 'estimator', 'disc', 'optim_ac', 'optim_hist_encoder', 'optim_estimator', 'optim_d',
                       'optim_q_eps', 'optim_q_c', 'disc_normalizer', 'reward_i_normalizer', 'iter', 'infos'] and d["iter"] == 3
    assert float(d["optim_ac"]["state"][0]["step"]) == 60 and float(d["optim_d"]["state"][0]["step"]) == 0
    r2 = _runner(str(tmp_path / "b"), 2)
    r2.load(str(tmp_path / "a" / "model.pt"))
    assert r2.current_learning_iteration == 3
    for (k, v), w in zip(r.alg.actor_critic.state_dict().items(), r2.alg.actor_critic.state_dict().values()):
        assert torch.equal(v, w), k
    sa, sb = r.alg.optimizer_state_dicts()["optim_ac"]["state"], r2.alg.optimizer_state_dicts()["optim_ac"]["state"]
    assert all(torch.equal(sa[i]["exp_avg"], sb[i]["exp_avg"]) and float(sa[i]["exp_avg"].min()) == 0.75 for i in sa)
    assert int(r2.alg.optim_ac.step_count) == 60          # (the stub also wrote the row padding, which a checkpoint does not carry)
    assert set(r2.alg._pending_disc_optim) == {"optim_d", "optim_q_eps", "optim_q_c"}
    r2.learn(1)                                                      # resumes at iteration 3
    assert [s for s, _ in r2.writer.scalars["Perf/total_fps"]] == [3]
    # the reference's log(locals()) entry point
    r2.log(dict(it=9, collection_time=0.5, learn_time=0.5, mean_surrogate_loss=1.0, mean_value_loss=2.0, mean_b_loss=0.0,
                mean_entropy_batch=3.0, mean_priv_reg_loss=4.0, mean_estimator_loss=5.0, mean_hist_latent_loss=6.0))
    assert r2.writer.scalars["Loss/value_loss"][-1] == (9, 2.0) and r2.writer.scalars["Loss/hist_latent_loss"][-1] == (9, 6.0)


def test_runner_builds_the_expert_sets_from_the_shipped_clips_like_the_reference_runner():
    """on_policy_runner.py:56-71: with `motion_files_lb / _ulb` in the runner cfg the expert transitions are preloaded at
    construction and `learn` hands them to the discriminator update.  Needs the reference's mocap clips (build container)."""
    import glob
    import os
    import pytest
    lb = glob.glob("/root/reference/bbc/mocap_data/mocap_all_lb/*")
    ulb = glob.glob("/root/reference/bbc/mocap_data/mocap_all_ulb/*")
    if not lb or not ulb:
        pytest.skip("the reference tree only exists in the build container")
    torch.manual_seed(0)
    cfg = bbc_train_cfg()
    cfg["runner"].update(num_steps_per_env=4, motion_files_lb=lb, motion_files_ulb=ulb[:3], num_preload_transitions=500)
    cfg["algorithm"].update(disc_replay_buffer_size=512, use_cuda_graph=False, fused_loss=False)
    r = OnPolicyRunner(FakeEnv(seed=3), cfg, log_dir=None, device="cpu")
    ml = r.alg.motion_loader
    assert tuple(ml.preloaded_s_lb.shape) == (500, 98) and tuple(ml.preloaded_s_ulb.shape) == (500, 98)
    assert ml.preloaded_label.shape == (500,) and int(ml.preloaded_label.min()) >= 0 and int(ml.preloaded_label.max()) <= 4
    assert torch.isfinite(ml.preloaded_s_lb).all() and torch.isfinite(ml.preloaded_s_ulb).all()
    s, lab = next(ml.feed_forward_generator_lb(1, 16))
    assert tuple(s.shape) == (16, 98) and lab.shape == (16,)


class FakeTscEnv:
    """What `OnPolicyRunnerTSC` touches of `LeggedRobotTSC`, with random dynamics."""

    def __init__(self, n=16, seed=0):
        from qa_b200.legged_robot_tsc import TscEnvConfig
        self.cfg = TscEnvConfig(num_envs=n)
        self.num_envs, self.num_obs, self.num_privileged_obs, self.num_actions = n, 800, None, 12
        self.num_obs_disc, self.num_obs_bbc, self.dim_c, self.dt = 49, 671, 5, 0.02
        self.g = torch.Generator().manual_seed(seed)
        self.episode_length_buf, self.max_episode_length = torch.zeros(n, dtype=torch.int64), 2000.0
        self._episode_rew_means = torch.zeros(len(self.cfg.reward_names))
        self.extras = {}
        self._draw()

    def _draw(self):
        n, g = self.num_envs, self.g
        self.obs_buf, self.obs_bbc_buf = 0.3 * torch.randn(n, 800, generator=g), 0.3 * torch.randn(n, 671, generator=g)
        self.obs_disc_buf = 0.3 * torch.randn(n, 49, generator=g)
        self.rew_buf, self.reset_buf = 0.05 * torch.rand(n, generator=g), torch.rand(n, generator=g) < 0.25

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return None

    def get_observations_bbc(self):
        return self.obs_bbc_buf

    def get_observations_disc(self):
        return self.obs_disc_buf

    def set_commands(self, actions, action_noise_u=None):
        assert actions.shape == (self.num_envs, 19)
        return torch.zeros(self.num_envs, 11)

    def step(self, actions_bbc, action_hl_history_buf=None):
        assert actions_bbc.shape == (self.num_envs, 12) and action_hl_history_buf.shape == (self.num_envs, 8, 19)
        self._draw()
        if bool(self.reset_buf.any()):
            self._episode_rew_means = torch.rand(len(self.cfg.reward_names), generator=self.g)
            self.extras["time_outs"] = self.reset_buf.clone()
        self.extras["reach_goal"] = torch.rand(self.num_envs, generator=self.g) < 0.5
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras, None, None


def test_tsc_teacher_learn_loop_logs_and_saves_on_host(tmp_path, monkeypatch):
    """`OnPolicyRunnerTSC.learn_RL` (tsc/rsl_rl/runners/on_policy_runner.py:164-276) host logic: rollout loop, high-level action
    history, bookkeeping incl. the success rate, the reference's tags, checkpoint keys.  K18 / K19 and the update are stubbed."""
    from qa_b200 import ops
    from qa_b200.config import tsc_train_cfg
    from qa_b200.rsl_rl.tsc_runner import OnPolicyRunnerTSC

    def disc_input(dones, prev_disc, next_disc, hist_prev, hist_new, hist_next, x_norm, mean, std, clip, *rest):
        hist_next.copy_(torch.stack([hist_prev[:, 1], next_disc], dim=1))
        x_norm.copy_(hist_next.reshape(len(dones), -1))

    def disc_reward(heads, obs, reward_t, dt, coefs, out, **kw):
        out.copy_(reward_t * coefs[3] + 0.01 * heads[:, 0])

    monkeypatch.setattr(ops, "disc_input", disc_input)
    monkeypatch.setattr(ops, "disc_reward", disc_reward)
    torch.manual_seed(0)
    cfg = tsc_train_cfg()
    cfg["runner"].update(num_steps_per_env=5, save_interval=1)
    cfg["algorithm"].update(use_cuda_graph=False, fused_loss=False)
    env = FakeTscEnv()
    r = OnPolicyRunnerTSC(env, cfg, log_dir=str(tmp_path), device="cpu")
    assert r.learn.__func__ is OnPolicyRunnerTSC.learn_RL
    alg = r.alg
    alg.compute_returns = lambda critic_obs: None
    alg.update = lambda: (alg.storage.clear(), (0.1, 0.2, 0.3, 0.0, 0.0, 0.4, 0.05))[1]
    alg.update_dagger = lambda: 0.6
    r.learn(2)
    sc = r.writer.scalars
    for tag in ("Loss/value_function", "Loss/surrogate", "Loss/estimator", "Loss/hist_latent_loss", "Loss/priv_reg_loss",
                "Loss/priv_ref_lambda", "Loss/learning_rate", "Policy/mean_noise_std", "Perf/total_fps", "Episode_rew/rew_reach_goal",
                "Train/mean_reward", "Train/mean_episode_length", "Train/mean_success_rate"):
        assert [s for s, _ in sc[tag]] == [0, 1], tag
    assert sc["Loss/hist_latent_loss"][0][1] == 0.6 and 0.0 <= env.success_rate <= 1.0
    assert r.action_history_buf.shape == (16, 8, 19) and float(r.action_history_buf.abs().sum()) > 0
    d = torch.load(str(tmp_path / "model.pt"), map_location="cpu", weights_only=False)
    assert list(d) == ["model_state_dict", "estimator_state_dict", "optimizer_state_dict", "iter", "infos"]
    assert len(d["optimizer_state_dict"]["param_groups"][0]["params"]) == sum(1 for _ in alg.actor_critic.parameters())
    r2 = OnPolicyRunnerTSC(FakeTscEnv(seed=1), cfg, log_dir=None, device="cpu")
    r2.load(str(tmp_path / "model.pt"))
    for (k, v), w in zip(alg.actor_critic.state_dict().items(), r2.alg.actor_critic.state_dict().values()):
        assert torch.equal(v, w), k


def test_tsc_runner_load_bbc_takes_the_shipped_checkpoint():
    """tsc/rsl_rl/runners/on_policy_runner.py:647-660 on the checkpoint the reference ships (`tsc/weights/bbc/model.pt`): the
    frozen controller, the estimator, the discriminator and its pickled normaliser."""
    import os
    import pytest
    from qa_b200.config import tsc_train_cfg
    from qa_b200.rsl_rl.tsc_runner import OnPolicyRunnerTSC
    from qa_b200.rsl_rl.utils import install_pickle_alias
    path = "/root/reference/tsc/weights/bbc/model.pt"
    if not os.path.exists(path):
        pytest.skip("the reference tree only exists in the build container")
    cfg = tsc_train_cfg()
    cfg["algorithm"].update(use_cuda_graph=False, fused_loss=False)
    r = OnPolicyRunnerTSC(FakeTscEnv(), cfg, log_dir=None, device="cpu")
    r.load_bbc(path)
    install_pickle_alias()
    ck = torch.load(path, map_location="cpu", weights_only=False)
    for k, v in ck["actor_critic"].items():
        assert torch.equal(r.actor_critic_bbc.state_dict()[k], v), k
    for k, v in ck["estimator"].items():
        assert torch.equal(r.estimator.state_dict()[k], v), k
    for k, v in ck["disc"].items():
        assert torch.equal(r.discriminator.state_dict()[k], v), k
    assert np.array_equal(r.disc_normalizer.mean, ck["disc_normalizer"].mean) and r.disc_normalizer.count == ck["disc_normalizer"].count
    assert not r.actor_critic_bbc.training and r.actor_critic_bbc.train_with_estimated_latent
    obs_bbc = 0.3 * torch.randn(4, 671, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a = r.get_inference_policy_bbc()(obs_bbc, hist_encoding=True)
    assert a.shape == (4, 12) and torch.isfinite(a).all()

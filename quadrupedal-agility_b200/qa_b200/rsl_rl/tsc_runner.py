"""`OnPolicyRunnerTSC` -- the teacher loop of tsc/rsl_rl/runners/on_policy_runner.py (`learn_RL`, :164-300) over
`LeggedRobotTSC`, the TSC `PPO`, a frozen BBC `ActorCritic` as the low-level controller and the style discriminator.

Per env step (:201-228): high-level act -> high-level action history -> `env.set_commands` -> the command lanes of the BBC
observation -> BBC `act_inference` -> `env.step` -> disc-history bookkeeping + style reward (K18 -> GEMMs -> K19, the same
fused path as the BBC runner; the TSC discriminator applies no task-obs weighting) -> `process_env_step`.
The depth-student loop (`learn_vision`, :302-460) is SURVEY 8(f)-3 and not part of this class.
"""
import time
import types

import torch

from .. import ops
from .modules import ActorCritic, Discriminator, Estimator
from .tsc import ActorCriticTSC, PPO
from .utils import Normalizer


class OnPolicyRunnerTSC:
    def __init__(self, env, train_cfg, log_dir=None, device='cpu'):
        self.cfg, self.alg_cfg = train_cfg["runner"], dict(train_cfg["algorithm"])
        self.policy_cfg, self.estimator_cfg = train_cfg["policy"], dict(train_cfg["estimator"])
        self.device, self.env = device, env
        ec = env.cfg
        self.n_proprio, self.n_auxiliary, self.n_scan, self.n_priv, self.n_priv_latent = 65, 8, 132, 4, 29
        self.history_len = ec.history_len
        self.num_actions_d, self.num_actions_c = len(ec.mocap_category), ec.num_actions_c
        self.num_actions = 1 + self.num_actions_d * self.num_actions_c
        self.num_command = self.num_actions_c + len(ec.mocap_category_all)
        self.num_obs_bbc = self.n_proprio - self.n_auxiliary + self.n_priv_latent + self.n_priv + self.num_command
        self.num_critic_obs = self.num_obs_bbc + self.history_len * (self.n_proprio - self.n_auxiliary)
        self.disc_obs_len, self.num_disc_obs = 2, env.num_obs_disc
        r = self.cfg
        self.actor_critic = ActorCriticTSC(self.n_proprio, self.n_auxiliary, self.n_scan, env.num_obs, self.n_priv_latent,
                                           self.n_priv, self.history_len, self.num_actions_d, self.num_actions_c,
                                           device=device, **self.policy_cfg).to(device)
        bbc_cfg = {k: v for k, v in self.policy_cfg.items() if k not in ("scan_encoder_dims", "tanh_encoder_output",
                                                                         "continue_from_last_std", "rnn_type",
                                                                         "rnn_hidden_size", "rnn_num_layers")}
        self.actor_critic_bbc = ActorCritic(self.num_obs_bbc, self.num_critic_obs, env.num_actions,
                                            self.n_proprio - self.n_auxiliary, self.history_len, self.n_priv,
                                            self.n_priv_latent, self.num_command, **bbc_cfg).to(device)
        self.estimator = Estimator(input_dim=self.n_proprio - self.n_auxiliary, output_dim=self.n_priv,
                                   hidden_dims=self.estimator_cfg["hidden_dims"]).to(device)
        est_paras = dict(priv_states_dim=self.n_priv, num_prop=self.n_proprio - self.n_auxiliary, num_auxiliary=self.n_auxiliary,
                         num_scan=self.n_scan, learning_rate=self.estimator_cfg.get("learning_rate", 1e-4),
                         train_with_estimated_states=self.estimator_cfg.get("train_with_estimated_states", True))
        self.alg = PPO(self.actor_critic, self.actor_critic_bbc, self.estimator, est_paras, device=device, **self.alg_cfg)
        self.num_steps_per_env, self.save_interval = r["num_steps_per_env"], r["save_interval"]
        self.dagger_update_freq = self.alg_cfg.get("dagger_update_freq", 20)
        self.alg.init_storage(env.num_envs, self.num_steps_per_env, [env.num_obs], [env.num_privileged_obs], [self.num_actions])
        self.disc_normalizer = Normalizer(self.num_disc_obs * self.disc_obs_len)
        denv = types.SimpleNamespace(task_obs_weight_decay=False, task_obs_weight=1.0, dim_c=env.dim_c)
        self.discriminator = Discriminator(denv, self.num_disc_obs * self.disc_obs_len, self.num_disc_obs, env.dim_c, env.dt,
                                           r["disc_loss_function"], None, r["reward_i_coef"], r["reward_us_coef"],
                                           r["reward_ss_coef"], r["reward_t_coef"], 2, self.disc_obs_len, 0.0,
                                           r["disc_hidden_units"], device).to(device)
        self.log_dir = log_dir
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0.0, 0
        self.perf = {}
        N, dev = env.num_envs, device
        self.action_history_buf = torch.zeros(N, ec.action_buf_len, self.num_actions, device=dev)
        self._disc_hist = None
        self._hist_pp = [torch.zeros(N, self.disc_obs_len, self.num_disc_obs, device=dev) for _ in range(2)]
        self._hist_new = torch.zeros(N, self.disc_obs_len * self.num_disc_obs, device=dev)
        w = self.disc_obs_len * self.num_disc_obs
        self._x_norm = torch.zeros(N, (w + 3) // 4 * 4, device=dev)[:, :w]
        self._rew = torch.zeros(N, device=dev)

    def load_bbc(self, state_dicts):
        """:647-660: the frozen low-level controller, its estimator and the discriminator come from a BBC checkpoint."""
        self.actor_critic_bbc.load_state_dict(state_dicts["actor_critic"])
        if "estimator" in state_dicts:
            self.estimator.load_state_dict(state_dicts["estimator"])
        if "disc" in state_dicts:
            self.discriminator.load_state_dict(state_dicts["disc"])

    @torch.no_grad()
    def rollout_step(self, obs, obs_bbc, critic_obs, infos, hist_encoding=False, normal_draw=None, mode_u=None,
                     action_noise_u=None):
        """One teacher rollout step (:201-228).  Returns (obs, obs_bbc, critic_obs, infos)."""
        env, alg = self.env, self.alg
        actions = alg.act(obs, critic_obs, infos, hist_encoding=hist_encoding, normal_draw=normal_draw, mode_u=mode_u)
        self.action_history_buf = torch.cat([self.action_history_buf[:, 1:], actions[:, None, :]], dim=1)
        next_commands = env.set_commands(actions, action_noise_u=action_noise_u)
        obs_bbc[:, -next_commands.shape[1]:] = next_commands
        actions_bbc = self.actor_critic_bbc.act_inference(obs_bbc, hist_encoding=True)
        prev_disc = env.get_observations_disc().clone()            # terminal disc state of the envs that reset (:264)
        obs, priv, rewards, dones, infos, _ids, _term = env.step(actions_bbc, self.action_history_buf)
        critic_obs = priv if priv is not None else obs
        next_obs_bbc, disc_obs = env.get_observations_bbc(), env.get_observations_disc()
        dst = self._hist_pp[0] if self._disc_hist is not self._hist_pp[0] else self._hist_pp[1]
        mean, std = self.disc_normalizer.device_moments(self.device)
        ops.disc_input(dones, prev_disc, disc_obs, self._disc_hist.contiguous(), self._hist_new, dst, self._x_norm, mean, std,
                       self.disc_normalizer.clip_obs, False, 1.0, 0.0)
        heads = self.discriminator.heads_forward(self._x_norm)
        d = self.discriminator
        ops.disc_reward(heads, obs_bbc, rewards, d.dt, (d.reward_i_coef, d.reward_us_coef, d.reward_ss_coef, d.reward_t_coef),
                        self._rew)
        alg.process_env_step(self._rew, dones, infos)
        self._disc_hist = dst
        return obs, next_obs_bbc.clone(), critic_obs, infos

    def learn_RL(self, num_learning_iterations, init_at_random_ep_len=False):
        env, alg = self.env, self.alg
        if init_at_random_ep_len:
            env.episode_length_buf.copy_(torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length)))
        obs, obs_bbc = env.get_observations(), env.get_observations_bbc().clone()
        priv = env.get_privileged_observations()
        critic_obs = priv if priv is not None else obs
        self._disc_hist = torch.stack([env.get_observations_disc()] * self.disc_obs_len, dim=1)
        infos = {}
        for it in range(self.current_learning_iteration, self.current_learning_iteration + num_learning_iterations):
            start = time.time()
            hist_encoding = it % self.dagger_update_freq == 0
            for _ in range(self.num_steps_per_env):
                obs, obs_bbc, critic_obs, infos = self.rollout_step(obs, obs_bbc, critic_obs, infos, hist_encoding)
            torch.cuda.synchronize() if torch.device(self.device).type == "cuda" else None
            stop = time.time()
            collection_time, start = stop - start, stop
            alg.compute_returns(critic_obs)
            stats = alg.update()
            learn_time = time.time() - start
            self.tot_timesteps += self.num_steps_per_env * env.num_envs
            self.tot_time += collection_time + learn_time
            self.perf = {"total_fps": int(self.num_steps_per_env * env.num_envs / (collection_time + learn_time)),
                         "collection_time": collection_time, "learning_time": learn_time, "stats": stats}
        self.current_learning_iteration += num_learning_iterations

    learn = learn_RL

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        return self.alg.actor_critic.act_inference

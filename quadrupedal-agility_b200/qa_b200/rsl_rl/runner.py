"""`OnPolicyRunner` -- rollout / learn loop with the reference's constructor, `learn`, `save`, `load` and
`get_inference_policy` (bbc/rsl_rl/runners/on_policy_runner.py:18-351).

Loop order per step is the reference's (:156-181): act -> env.step -> disc-history update with the terminal
states patched in -> predict_disc_reward -> process_env_step -> reset rows of the disc history.  The
variable-length `reset_env_ids` indexing of the reference is replaced by dense masked selects on the reset
mask (same values, no host sync).  The episode bookkeeping and TensorBoard scalars (:183-206, :238-304) are staged on
the device and flushed once per iteration (`train_log.py`); they are active when `log_dir` is given, as in the reference.
"""
import os
import time

import torch

from .. import config as K
from .. import ops
from .algorithm import SSInfoGAIL
from .modules import ActorCritic, Estimator, Discriminator
from .train_log import EpisodeBook, ScalarLog, log_bbc
from .utils import Normalizer, install_pickle_alias, reference_picklable


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device='cpu', motion_loader=None):
        self.device, self.env = device, env
        self.cfg, self.alg_cfg = train_cfg["runner"], dict(train_cfg["algorithm"])
        self.policy_cfg, self.estimator_cfg = train_cfg["policy"], train_cfg["estimator"]
        self.disc_loss_function = self.alg_cfg["disc_loss_function"]
        r = self.cfg
        self.disc_history_len, self.disc_obs_len, self.obs_disc_weight_step = 2, K.DISC_OBS_LEN, 0.0
        num_prop, num_hist, num_explicit = K.NUM_PROP, K.HISTORY_LEN, K.NUM_EXPLICIT
        num_latent, num_command = K.NUM_LATENT, K.NUM_COMMAND
        num_actor_obs = env.num_obs
        num_critic_obs = env.num_obs + num_hist * num_prop
        num_disc_obs = env.num_obs_disc
        actor_critic = ActorCritic(num_actor_obs, num_critic_obs, env.num_actions, num_prop, num_hist, num_explicit,
                                   num_latent, num_command, **self.policy_cfg).to(device)
        estimator = Estimator(input_dim=num_prop, output_dim=num_explicit,
                              hidden_dims=self.estimator_cfg["hidden_dims"]).to(device)
        disc_normalizer = Normalizer(num_disc_obs * self.disc_obs_len)
        discriminator = Discriminator(env, num_disc_obs * self.disc_obs_len, num_disc_obs, len(env.mocap_category),
                                      env.dt, self.disc_loss_function, None, r["reward_i_coef"], r["reward_us_coef"],
                                      r["reward_ss_coef"], r["reward_t_coef"], self.disc_history_len, self.disc_obs_len,
                                      self.obs_disc_weight_step, r["disc_hidden_units"], device).to(device)
        if motion_loader is None and r.get("motion_files_lb"):
            motion_loader = self.build_motion_loader(env, r, device)
        min_std = (torch.tensor(r["min_normalized_std"], device=device) *
                   torch.abs(env.dof_pos_limits[:, 1] - env.dof_pos_limits[:, 0]))
        self.alg = SSInfoGAIL(env, actor_critic, discriminator, estimator, self.estimator_cfg, motion_loader,
                              disc_normalizer, self.disc_history_len, self.disc_obs_len, num_disc_obs,
                              self.obs_disc_weight_step, device=device, min_std=min_std, **self.alg_cfg)
        self.num_steps_per_env = r["num_steps_per_env"]
        self.save_interval = r["save_interval"]
        self.dagger_update_freq = r["dagger_update_freq"]
        self.alg.init_storage(env.num_envs, self.num_steps_per_env, [num_actor_obs + num_hist * num_prop],
                              [env.num_privileged_obs + num_hist * num_prop], [env.num_actions])
        self.log_dir = log_dir
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0.0, 0
        self.perf = {}
        self._disc_hist = None
        self._hist_pp = None
        self.book, self.writer = None, None
        self.fused_rollout = bool(train_cfg["runner"].get("fused_rollout", True))
        env.reset()

    def build_motion_loader(self, env, r, device):
        """What the reference's runner does at :56-71 (`MotionLoader(..., preload_transitions=True)`): the labelled clips as a
        `MocapTable`, the unlabelled ones as one trajectory, `num_preload_transitions` expert windows of each blended on the
        device (`ExpertData.build`, bit-exact against `MotionLoader.pre_load_data` under injected draws)."""
        from ..expert import ExpertData, load_unlabelled_clips
        from ..mocap import MocapTable
        fds = getattr(getattr(env.cfg, "env", None), "frame_duration_scale", 1.0)
        table = MocapTable.from_json_files(sorted(r["motion_files_lb"]), mocap_category=list(env.mocap_category), frame_duration_scale=fds)
        ulb = load_unlabelled_clips(sorted(r["motion_files_ulb"]), frame_duration_scale=fds)
        return ExpertData.build(table, ulb, int(r["num_preload_transitions"]), env.dt, env.default_dof_pos, env.obs_scales,
                                disc_obs_len=self.disc_obs_len, device=device, seed=int(r.get("seed", 0)))

    def _ensure_rollout_plan(self):
        """The static rollout schedule (rollout_plan) when it applies: CUDA, tcgen05 layers, MSE style reward, QA_ROLLOUT_PLAN != 0."""
        from . import linear
        from .rollout_plan import RolloutPlan
        if getattr(self, "_rollout_plan_off", None) is None:
            self._rollout_plan_off = os.environ.get("QA_ROLLOUT_PLAN", "1") != "1"
        if self._rollout_plan_off or linear.get_mode() != "tc" or not hasattr(self.env, "_time_outs_latched"):
            return None
        plan = getattr(self, "_rollout_plan", None)
        if plan is None:
            if RolloutPlan.supported(self.alg) is not None:
                self._rollout_plan_off = True
                return None
            plan = self._rollout_plan = RolloutPlan(self.alg, self.env, self.disc_obs_len, self.obs_disc_weight_step)
            self.alg._disc_listeners = list(getattr(self.alg, "_disc_listeners", ())) + [plan.refresh_disc_heads]
        # whole-rollout callers (learn(), pipeline.BbcIteration) let the reward tail of step t run next to step t+1
        plan.defer_reward_tail = bool(getattr(self, "full_rollouts", False)) and os.environ.get("QA_DEFER_REWARD_TAIL", "1") == "1"
        return plan

    def finish_rollout(self):
        """After the last step of a rollout: a deferred reward tail (rollout_plan) is joined into the current stream."""
        plan = getattr(self, "_rollout_plan", None)
        if plan is not None:
            plan.finish_rewards()

    # ---- one rollout step (:156-181) ----------------------------------------------------------------------
    def _rollout_step_fused(self, obs, critic_obs, hist_encoding):
        """Same step with the disc-history bookkeeping, the normalisation, the reward tail and the time-out bootstrap in two
        kernels (K18, K19) around the discriminator GEMMs."""
        env, alg = self.env, self.alg
        plan = self._ensure_rollout_plan()
        if plan is not None:
            actions = plan.act(obs, critic_obs, hist_encoding)
            prev_disc = env.get_disc_observations()
            next_obs, next_priv, rewards, dones, _ids, _cnt, _term = env.step_device(actions)
            book = self.book
            self._disc_hist = plan.reward(obs, rewards, dones, prev_disc, env.get_disc_observations(), self._disc_hist.contiguous(),
                                          reward_terms=None if book is None else self._terms4)
            if book is not None:
                slot = book.term_slot()
                slot[:, 1:].copy_(self._terms4)
                slot[:, 0].copy_(self._terms4 @ self._reward_coefs)
                book.record(dones, episode_means=env._episode_rew_means, num_resets=getattr(env, '_num_resets', None))
            return next_obs, next_priv
        N, L, W = env.num_envs, self.disc_obs_len, env.num_obs_disc
        if self._hist_pp is None:
            dev = self.device
            self._hist_pp = [torch.zeros(N, L, W, device=dev) for _ in range(2)]
            self._hist_new = torch.zeros(N, L * W, device=dev)
            self._x_norm = torch.zeros(N, (L * W + 3) // 4 * 4, device=dev)[:, :L * W]
        actions = alg.act(obs, critic_obs, hist_encoding)
        prev_disc = env.get_disc_observations()
        next_obs, next_priv, rewards, dones, _ids, _cnt, _term = env.step_device(actions)
        next_disc = env.get_disc_observations()
        staged = alg._disc_stage is not None
        hist_new = alg._disc_stage[0][alg.storage.step] if staged else self._hist_new
        dst = self._hist_pp[0] if self._disc_hist is not self._hist_pp[0] else self._hist_pp[1]
        mean, std = alg.disc_normalizer.device_moments(self.device)
        ops.disc_input(dones, prev_disc, next_disc, self._disc_hist.contiguous(), hist_new, dst, self._x_norm, mean, std,
                       alg.disc_normalizer.clip_obs, env.task_obs_weight_decay, env.task_obs_weight, self.obs_disc_weight_step)
        heads = alg.disc.heads_forward(self._x_norm)
        infos = {"time_outs": env._time_outs_latched} if env.cfg.send_timeouts else {}
        book = self.book
        alg.process_env_step_fused(heads, obs, rewards, dones, infos, None if staged else hist_new,
                                   reward_terms=None if book is None else self._terms4)
        if book is not None:                                                # :183-206, staged on the device
            slot = book.term_slot()
            slot[:, 1:].copy_(self._terms4)
            slot[:, 0].copy_(self._terms4 @ self._reward_coefs)
            book.record(dones, episode_means=env._episode_rew_means, num_resets=getattr(env, '_num_resets', None))
        self._disc_hist = dst
        return next_obs, next_priv

    def rollout_step(self, obs, critic_obs, hist_encoding=False):
        env, alg = self.env, self.alg
        if self.fused_rollout and torch.device(self.device).type == "cuda" and self.disc_loss_function == "MSELoss":
            return self._rollout_step_fused(obs, critic_obs, hist_encoding)
        actions = alg.act(obs, critic_obs, hist_encoding)
        prev_disc = env.get_disc_observations()
        next_obs, next_priv, rewards, dones, _ids, _cnt, _term = env.step_device(actions)
        next_disc = env.get_disc_observations()
        done_col = dones.unsqueeze(1)
        # terminal disc state of a reset env == its disc obs of the previous step (:153-154, :166-167)
        with_term = torch.where(done_col, prev_disc, next_disc)
        hist = torch.stack([self._disc_hist[:, 1], with_term], dim=1)
        rew, r_i, r_us, r_ss, r_t = alg.disc.predict_disc_reward(rewards.unsqueeze(1), obs, hist,
                                                                 normalizer=alg.disc_normalizer)
        infos = {"time_outs": env._time_outs_latched} if env.cfg.send_timeouts else {}
        alg.process_env_step(rew, dones, infos, hist)
        if self.book is not None:
            self.book.record(dones, torch.stack([rew.float(), r_i, r_us, r_ss.float(), r_t], dim=1), env._episode_rew_means, getattr(env, '_num_resets', None))
        # fresh episodes restart their discriminator history (:180-181)
        self._disc_hist = torch.where(done_col.unsqueeze(2), next_disc.unsqueeze(1).expand(-1, self.disc_obs_len, -1), hist)
        return next_obs, next_priv

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        env, alg = self.env, self.alg
        if init_at_random_ep_len:
            env.episode_length_buf.copy_(torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length)))
        obs, critic_obs = env.get_observations(), env.get_privileged_observations()
        self._disc_hist = torch.stack([env.get_disc_observations()] * self.disc_obs_len, dim=1)
        alg.actor_critic.train()
        if self.log_dir is not None and self.writer is None:               # :120-121
            self.writer = ScalarLog(self.log_dir)
            self.book = EpisodeBook(env.num_envs, self.num_steps_per_env, ("total", "i", "us", "ss", "t"), self.device,
                                    num_episode_keys=len(env.reward_names))
            d = alg.disc
            self._terms4 = torch.zeros(env.num_envs, 4, device=self.device)
            self._reward_coefs = torch.tensor([d.reward_i_coef, d.reward_us_coef, d.reward_ss_coef, d.reward_t_coef],
                                              device=self.device)
        hist_latent_loss = None
        self.full_rollouts = True                                           # every rollout below runs all num_steps_per_env steps
        for it in range(self.current_learning_iteration, self.current_learning_iteration + num_learning_iterations):
            start = time.time()
            hist_encoding = it % self.dagger_update_freq == 0
            with torch.inference_mode(False), torch.no_grad():
                for _ in range(self.num_steps_per_env):
                    obs, critic_obs = self.rollout_step(obs, critic_obs, hist_encoding)
                self.finish_rollout()
                torch.cuda.synchronize() if torch.device(self.device).type == "cuda" else None
                stop = time.time()
                collection_time = stop - start
                start = stop
                alg.compute_returns(critic_obs)
            expert = alg.motion_loader if hasattr(alg.motion_loader, "preloaded_s_lb") else None
            stats = alg.update(expert=expert)
            if hist_encoding:                                                   # on_policy_runner.py:220-221
                hist_latent_loss = self.perf_hist_latent_loss = alg.update_dagger()
            if env.task_obs_weight_decay_steps:
                env.task_obs_weight = max(0, env.task_obs_weight - 1.0 / env.task_obs_weight_decay_steps)
            learn_time = time.time() - start
            self.tot_timesteps += self.num_steps_per_env * env.num_envs
            self.tot_time += collection_time + learn_time
            self.perf = {"total_fps": int(self.num_steps_per_env * env.num_envs / (collection_time + learn_time)),
                         "collection_time": collection_time, "learning_time": learn_time, "stats": stats}
            if self.book is not None:                                           # :229-230
                self.book.flush()
                self.train_means = log_bbc(self, self.writer, it, stats, hist_latent_loss, collection_time, learn_time)
            if self.log_dir is not None and (it + 1) % self.save_interval == 0:
                self.save(os.path.join(self.log_dir, 'model.pt'))
        self.current_learning_iteration += num_learning_iterations
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, 'model.pt'))

    # ---- checkpoints (:306-339): the reference's dict, key for key, optimiser states in torch.optim layout ------------
    def save(self, path, infos=None):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        d = {'actor_critic': self.alg.actor_critic.state_dict(), 'estimator': self.alg.estimator.state_dict(),
             'disc': self.alg.disc.state_dict()}
        d.update(self.alg.optimizer_state_dicts())
        d.update({'disc_normalizer': reference_picklable(self.alg.disc_normalizer),
                  'reward_i_normalizer': getattr(self.alg.disc, "reward_i_normalizer", None),
                  'iter': self.current_learning_iteration, 'infos': infos})
        torch.save(d, path)

    def load(self, path, load_optimizer=True):
        install_pickle_alias()
        d = torch.load(path, map_location=self.device, weights_only=False)
        self.alg.actor_critic.load_state_dict(d['actor_critic'])
        self.alg.estimator.load_state_dict(d['estimator'])
        self.alg.disc.load_state_dict(d['disc'])
        n = d['disc_normalizer']
        if n is not None:
            mine = self.alg.disc_normalizer
            if mine is None or mine.mean.shape != n.mean.shape:
                mine = self.alg.disc_normalizer = Normalizer(n.mean.shape[0], epsilon=n.epsilon, clip_obs=n.clip_obs)
            mine.load_moments(n.mean, n.var, n.count, epsilon=n.epsilon, clip_obs=n.clip_obs)
            self.alg._disc_graph_key = None            # a captured discriminator step is re-captured on the next update
        if d.get('reward_i_normalizer'):
            self.alg.disc.reward_i_normalizer = d['reward_i_normalizer']
        if load_optimizer:
            self.alg.load_optimizer_state_dicts(d)
        self.alg.notify_disc_changed()
        self.current_learning_iteration = d.get('iter', 0)
        return d.get('infos')

    def log(self, locs, pbar=None):
        """The reference's entry point (:238-304), fed with its `locals()`-style dict: `it`, `collection_time`, `learn_time` and the
        `mean_*` statistics (the six PPO means; the eleven discriminator means when present)."""
        names = ("mean_surrogate_loss", "mean_value_loss", "mean_b_loss", "mean_entropy_batch", "mean_priv_reg_loss",
                 "mean_estimator_loss", "mean_ss_loss", "mean_info_max_loss", "mean_disc_loss", "mean_us_loss", "mean_grad_pen_loss",
                 "mean_disc_logit_loss", "mean_disc_weight_decay", "mean_acc_lb", "mean_acc_pi", "mean_acc_exp", "mean_acc_ulb")
        stats = tuple(locs[k] for k in names if k in locs)
        if self.writer is None:
            self.writer = ScalarLog(self.log_dir)
        if self.book is None:
            self.book = EpisodeBook(self.env.num_envs, self.num_steps_per_env, ("total", "i", "us", "ss", "t"), self.device,
                                    num_episode_keys=len(self.env.reward_names))
        return log_bbc(self, self.writer, locs["it"], stats, locs.get("mean_hist_latent_loss"), locs["collection_time"],
                       locs["learn_time"])

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_inference

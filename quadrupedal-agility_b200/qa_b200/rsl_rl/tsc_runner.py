"""`OnPolicyRunnerTSC` -- the teacher loop of tsc/rsl_rl/runners/on_policy_runner.py (`learn_RL`, :164-300) over
`LeggedRobotTSC`, the TSC `PPO`, a frozen BBC `ActorCritic` as the low-level controller and the style discriminator.

Per env step (:201-228): high-level act -> high-level action history -> `env.set_commands` -> the command lanes of the BBC
observation -> BBC `act_inference` -> `env.step` -> disc-history bookkeeping + style reward (K18 -> GEMMs -> K19, the same
fused path as the BBC runner; the TSC discriminator applies no task-obs weighting) -> `process_env_step`.
`learn_vision` (:278-420) is the depth-student loop: per step the teacher's scan-dot latent and action are
computed without grad, the student (depth encoder on the K14-preprocessed depth image -> latent / yaw / obstacle type ->
copied actor) WITH grad; the student's action drives the frozen BBC controller; once per iteration
`PPO.update_depth_actor` back-propagates the distillation loss through the whole rollout.  The reference's per-step
boolean-mask writes and `nonzero` counts (device syncs) are dense `torch.where` / device-side means here.

`save` / `load` / `load_bbc` keep the reference's checkpoint keys (:611-660).
"""
import copy
import os
import time
import types
import warnings

import torch
import torch.nn.functional as F

from .. import ops
from . import checkpoint as ckpt
from .depth_backbone import DepthOnlyFCBackbone58x87, RecurrentDepthBackbone
from .modules import ActorCriticBBC, Discriminator, Estimator
from .train_log import EpisodeBook, ScalarLog, log_tsc
from .tsc import ActorCriticTSC, PPO
from .utils import Normalizer, install_pickle_alias


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
        # the fork's ActorCriticBBC (:72-81): acts on the history-encoder latent (train_with_estimated_latent defaults to True)
        self.actor_critic_bbc = ActorCriticBBC(self.num_obs_bbc, self.num_critic_obs, env.num_actions, self.n_proprio,
                                               self.n_auxiliary, self.history_len, self.n_priv, self.n_priv_latent,
                                               self.num_command, **bbc_cfg).to(device)
        self.estimator = Estimator(input_dim=self.n_proprio - self.n_auxiliary, output_dim=self.n_priv,
                                   hidden_dims=self.estimator_cfg["hidden_dims"]).to(device)
        est_paras = dict(priv_states_dim=self.n_priv, num_prop=self.n_proprio - self.n_auxiliary, num_auxiliary=self.n_auxiliary,
                         num_scan=self.n_scan, learning_rate=self.estimator_cfg.get("learning_rate", 1e-4),
                         train_with_estimated_states=self.estimator_cfg.get("train_with_estimated_states", True))
        # depth student (:85-99): the student actor starts as a copy of the teacher's actor, the backbone augments its input
        self.depth_encoder_cfg = dict(train_cfg.get("depth_encoder", {"if_depth": False}))
        self.if_depth = bool(self.depth_encoder_cfg.get("if_depth", False))
        self.n_delta_yaw, self.n_obst_type = 2, self.n_auxiliary - 2
        self.n_depth_latent = self.policy_cfg["scan_encoder_dims"][-1]
        if self.if_depth:
            # on the device from the start: the BYOL learner sizes its projector with a dummy forward in its constructor, and
            # with > 1 rank its batch norms are SyncBatchNorm, which refuses host tensors
            self.depth_backbone = DepthOnlyFCBackbone58x87(self.n_proprio, self.n_depth_latent,
                                                           self.depth_encoder_cfg["hidden_dims"]).to(device)
            env_ns = types.SimpleNamespace(n_delta_yaw=self.n_delta_yaw, n_obst_type=self.n_obst_type, n_proprio=self.n_proprio)
            self.depth_encoder = RecurrentDepthBackbone(self.depth_backbone, self.n_depth_latent, env_ns).to(device)
            self.depth_actor = copy.deepcopy(self.actor_critic.actor)
            self.depth_backbone.augment = self.depth_encoder.byol_learner.augment1
        else:
            self.depth_encoder = self.depth_actor = None
        self.alg = PPO(self.actor_critic, self.actor_critic_bbc, self.estimator, est_paras, self.depth_encoder,
                       self.depth_encoder_cfg, self.depth_actor, device=device, **self.alg_cfg)
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
        self.writer, self.book = None, None
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0.0, 0
        self.perf = {}
        self.learn = self.learn_vision if self.if_depth else self.learn_RL          # :147
        N, dev = env.num_envs, device
        self.action_history_buf = torch.zeros(N, ec.action_buf_len, self.num_actions, device=dev)
        self._disc_hist = None
        self._hist_pp = [torch.zeros(N, self.disc_obs_len, self.num_disc_obs, device=dev) for _ in range(2)]
        self._hist_new = torch.zeros(N, self.disc_obs_len * self.num_disc_obs, device=dev)
        w = self.disc_obs_len * self.num_disc_obs
        self._x_norm = torch.zeros(N, (w + 3) // 4 * 4, device=dev)[:, :w]
        self._rew = torch.zeros(N, device=dev)

    def load_bbc(self, path_or_dicts):
        """:647-660: the frozen low-level controller, (optionally) its estimator, the discriminator and its normaliser come
        from a BBC checkpoint -- a path to one (e.g. the shipped `tsc/weights/bbc/model.pt`) or the loaded dict."""
        d = path_or_dicts
        if not isinstance(d, dict):
            install_pickle_alias()
            d = torch.load(d, map_location=self.device, weights_only=False)
        self.actor_critic_bbc.load_state_dict(d["actor_critic"])
        if "estimator" in d and self.estimator_cfg.get("load_estimator_bbc", True):
            self.estimator.load_state_dict(d["estimator"])
        if "disc" in d:
            self.discriminator.load_state_dict(d["disc"])
        n = d.get("disc_normalizer")
        if n is not None:
            mine = self.disc_normalizer                 # in place: the discriminator (and anything captured) holds this object
            if mine is None or mine.mean.shape != n.mean.shape:
                mine = self.disc_normalizer = Normalizer(n.mean.shape[0], epsilon=n.epsilon, clip_obs=n.clip_obs)
                if hasattr(self.discriminator, "normalizer"):
                    self.discriminator.normalizer = mine
            mine.load_moments(n.mean, n.var, n.count, epsilon=n.epsilon, clip_obs=n.clip_obs)
        if d.get("reward_i_normalizer"):
            self.discriminator.reward_i_normalizer = d["reward_i_normalizer"]
        self.actor_critic_bbc.eval()
        self.discriminator.eval()
        return d.get("infos")

    # ---- checkpoints (:611-645): the reference's keys; `optimizer_state_dict` in torch.optim.Adam layout ---------------
    def save(self, path, infos=None):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        alg = self.alg
        groups = [dict(adam=alg.optimizer, params=ckpt.layout(alg.ac_flat, alg.actor_critic))]
        d = {"model_state_dict": alg.actor_critic.state_dict(), "estimator_state_dict": alg.estimator.state_dict(),
             "optimizer_state_dict": ckpt.to_torch_state_dict(groups), "iter": self.current_learning_iteration, "infos": infos}
        if self.if_depth:
            d["depth_encoder_state_dict"] = alg.depth_encoder.state_dict()
            d["depth_actor_state_dict"] = alg.depth_actor.state_dict()
        torch.save(d, path)

    def load(self, path, load_optimizer=True):
        d = torch.load(path, map_location=self.device, weights_only=False)
        alg = self.alg
        alg.actor_critic.load_state_dict(d["model_state_dict"])
        alg.estimator.load_state_dict(d["estimator_state_dict"])
        if self.if_depth:
            if "depth_encoder_state_dict" not in d:
                warnings.warn("'depth_encoder_state_dict' key does not exist, not loading depth encoder...")
            else:
                alg.depth_encoder.load_state_dict(d["depth_encoder_state_dict"])
            if "depth_actor_state_dict" in d:
                alg.depth_actor.load_state_dict(d["depth_actor_state_dict"])
            else:                                                               # :634-636: start the student from the teacher
                alg.depth_actor.load_state_dict(alg.actor_critic.actor.state_dict())
        if load_optimizer and ckpt.is_torch_state_dict(d.get("optimizer_state_dict")):
            ckpt.from_torch_state_dict(d["optimizer_state_dict"], [dict(adam=alg.optimizer, params=ckpt.layout(alg.ac_flat, alg.actor_critic))])
        return d.get("infos")                                                   # the reference does not restore `iter` (:641)

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
        if self.book is not None:                                               # :230-260, staged on the device
            slot = self.book.term_slot()
            slot[:, 0].copy_(self._rew)
            slot[:, 1].copy_(infos["reach_goal"])
            self.book.record(dones, episode_means=env._episode_rew_means, num_resets=getattr(env, '_num_resets', None))
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
        self._open_log(("total",), self.num_steps_per_env)
        hist_latent_loss = 0.0
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
            if hist_encoding:                                                   # :264-266
                hist_latent_loss = alg.update_dagger()
            learn_time = time.time() - start
            self.tot_timesteps += self.num_steps_per_env * env.num_envs
            self.tot_time += collection_time + learn_time
            self.perf = {"total_fps": int(self.num_steps_per_env * env.num_envs / (collection_time + learn_time)),
                         "collection_time": collection_time, "learning_time": learn_time, "stats": stats,
                         "hist_latent_loss": hist_latent_loss}
            if self.book is not None:
                self.book.flush()
                v, s_, e, dl, da, pr, coef = stats
                tags = {"Loss/value_function": v, "Loss/surrogate": s_, "Loss/estimator": e, "Loss/hist_latent_loss": hist_latent_loss,
                        "Loss/priv_reg_loss": pr, "Loss/priv_ref_lambda": coef, "Loss/learning_rate": alg.learning_rate,
                        "Loss/discriminator": dl, "Loss/discriminator_accuracy": da,
                        "Policy/mean_noise_std": alg.actor_critic.std.mean().item()}
                m = log_tsc(self, self.writer, it, tags, collection_time, learn_time)
                if "reach_goal" in m:
                    env.success_rate = m["reach_goal"]                         # :270-271
            if self.log_dir is not None and it % self.save_interval == 0:      # :274-275
                self.save(os.path.join(self.log_dir, "model.pt"))
        self.current_learning_iteration += num_learning_iterations

    def _open_log(self, term_names, num_steps, maxlen=100):
        if self.log_dir is not None and self.writer is None:
            self.writer = ScalarLog(self.log_dir)
            self.book = EpisodeBook(self.env.num_envs, num_steps, term_names, self.device, maxlen=maxlen,
                                    num_episode_keys=len(self.env.cfg.reward_names), snap_names=("reach_goal",))

    # ---- depth student (:278-420) ----------------------------------------------------------------------------------------
    def student_step(self, obs, obs_bbc, infos, action_student_history_buf, buf, use_teacher_actions=False):
        """One step of `learn_vision`'s rollout (:320-385).  Appends this step's distillation tensors to the lists in `buf`;
        returns (obs, obs_bbc, infos, action_student_history_buf, rewards, dones)."""
        env, alg = self.env, self.alg
        P, A, Y = self.n_proprio, self.n_auxiliary, self.n_delta_yaw
        L = self.n_depth_latent
        if infos["depth"] is not None:
            with torch.no_grad():
                o = alg.num_prop + alg.num_auxiliary + alg.num_scan
                obs[:, o:o + alg.priv_states_dim] = alg.estimator(obs[:, :alg.num_prop])
                scandots_latent = alg.actor_critic.actor.infer_scandots_latent(obs)
            buf["scandots_latent"].append(scandots_latent)
            prop = obs[:, :P].clone()
            prop[:, P - A:P] = 0                                                # mask delta_yaw, obstacle type
            out = alg.depth_encoder(infos["depth"].clone(), prop)
            depth_latent, delta_yaw, obst_type = out[:, :L], 1.5 * out[:, L:L + Y], out[:, L + Y:]
            buf["depth"].append(infos["depth"].clone())
            buf["depth_latent"].append(depth_latent)
            buf["yaw_student"].append(delta_yaw)
            buf["yaw_teacher"].append(obs[:, P - A:P - A + Y].clone())            # the env overwrites its observation buffer in
            buf["obst_student"].append(obst_type)                                 # place next step (the reference rebinds a new
            buf["obst_teacher"].append(obs[:, P - A + Y:P].clone())               # tensor, :515): keep copies, not views
            self._student_last = (depth_latent, delta_yaw, obst_type)
        depth_latent, delta_yaw, obst_type = self._student_last                 # :347-352 reuse the last depth inference
        with torch.no_grad():
            buf["actions_teacher"].append(alg.actor_critic.act_inference(obs, hist_encoding=True, scandots_latent=None))
        ok = infos["delta_yaw_ok"]
        obs_student = obs.clone()
        obs_student[:, P - A:P - A + Y] = torch.where(ok.unsqueeze(1), delta_yaw.detach(), obs_student[:, P - A:P - A + Y])
        obs_student[:, P - A + Y:P] = F.one_hot(torch.argmax(obst_type.detach(), dim=-1), num_classes=obst_type.shape[-1]).to(obs.dtype)
        buf["delta_yaw_ok"].append(ok.float().mean())
        emb = alg.depth_actor(obs_student, hist_encoding=True, scandots_latent=depth_latent)
        prob, mean = alg.depth_actor.actor_d(emb), alg.depth_actor.actor_c(emb)
        actions_student = torch.cat([torch.argmax(prob, dim=-1).unsqueeze(-1).to(mean.dtype), mean], dim=-1)
        buf["actions_student"].append(torch.cat([prob, mean], dim=-1))
        action_student_history_buf = torch.cat([action_student_history_buf[:, 1:], actions_student[:, None, :].detach()], dim=1)
        drive = buf["actions_teacher"][-1] if use_teacher_actions else action_student_history_buf[:, -1]
        with torch.no_grad():
            next_commands = env.set_commands(drive.detach())
            obs_bbc[:, -next_commands.shape[1]:] = next_commands
            actions_bbc = alg.actor_critic_bbc.act_inference(obs_bbc, hist_encoding=True).detach()
            obs, _priv, rewards, dones, infos, _ids, _term = env.step(actions_bbc)
            obs_bbc = env.get_observations_bbc().clone()
            action_student_history_buf = action_student_history_buf * (~dones.bool()).to(obs.dtype)[:, None, None]
        return obs, obs_bbc, infos, action_student_history_buf, rewards, dones

    def learn_vision(self, num_learning_iterations, init_at_random_ep_len=False, num_pretrain_iter=0):
        env, alg, cfg = self.env, self.alg, self.depth_encoder_cfg
        T = cfg["num_steps_per_env"]
        if hasattr(env, "reconfigure"):                                     # :281-286 (observation noise / curriculum are
            env.reconfigure(next_goal_threshold=0.45)                       # simulator-side switches, not on this path)
        self._open_log(("total",), T, maxlen=1000)
        hist = torch.zeros(env.num_envs, env.cfg.action_buf_len, self.num_actions, device=self.device)
        obs, obs_bbc = env.get_observations(), env.get_observations_bbc().clone()
        infos = {"depth": env.depth_buffer.clone()[:, -1] if self.if_depth else None,
                 "delta_yaw_ok": torch.ones(env.num_envs, dtype=torch.bool, device=self.device)}
        alg.depth_encoder.train()
        alg.depth_actor.train()
        keys = ("depth", "depth_latent", "scandots_latent", "actions_teacher", "actions_student", "yaw_student", "yaw_teacher",
                "obst_student", "obst_teacher", "delta_yaw_ok")
        tot_iter = self.current_learning_iteration + num_learning_iterations
        for it in range(self.current_learning_iteration, tot_iter):
            start = time.time()
            buf = {k: [] for k in keys}
            for _ in range(T):
                obs, obs_bbc, infos, hist, rewards, dones = self.student_step(obs, obs_bbc, infos, hist, buf, it < num_pretrain_iter)
                if self.book is not None:
                    slot = self.book.term_slot()
                    slot[:, 0].copy_(rewards)
                    slot[:, 1].copy_(infos["reach_goal"])
                    self.book.record(dones, episode_means=env._episode_rew_means, num_resets=getattr(env, '_num_resets', None))
            torch.cuda.synchronize() if torch.device(self.device).type == "cuda" else None
            stop = time.time()
            collection_time, start = stop - start, stop
            cat = {k: torch.cat(v, dim=0) for k, v in buf.items() if k != "delta_yaw_ok"}
            actor_loss, yaw_loss, obst_loss, byol_loss = alg.update_depth_actor(
                cat["actions_student"], cat["actions_teacher"], cat["yaw_student"], cat["yaw_teacher"], cat["obst_student"],
                cat["obst_teacher"], cat["depth"])
            learn_time = time.time() - start
            alg.depth_encoder.detach_hidden_states()
            self._student_last = tuple(t.detach() for t in self._student_last)
            lr0, lr_byol, lr_min = cfg["learning_rate"], cfg["learning_rate_byol"], cfg["learning_rate_min"]
            for opt, base in ((alg.depth_encoder_optimizer, lr0), (alg.depth_actor_optimizer, lr0), (alg.byol_optimizer, lr_byol)):
                for g in opt.param_groups:                                      # linear decay over 20 000 iterations (:405-416)
                    g["lr"] = max(base - (base - lr_min) * it / 20000, lr_min)
            self.tot_timesteps += T * env.num_envs
            self.tot_time += collection_time + learn_time
            ok_pct = float(torch.stack(buf["delta_yaw_ok"]).mean().item())
            self.perf = {"total_fps": int(T * env.num_envs / (collection_time + learn_time)), "collection_time": collection_time,
                         "learning_time": learn_time, "depth_actor_loss": actor_loss, "yaw_loss": yaw_loss,
                         "obst_type_loss": obst_loss, "byol_loss": byol_loss, "delta_yaw_ok_percentage": ok_pct}
            if self.book is not None:
                self.book.flush()
                tags = {"Loss_depth/delta_yaw_ok_percent": ok_pct, "Loss_depth/depth_actor": actor_loss, "Loss_depth/yaw": yaw_loss,
                        "Loss_depth/obst_type": obst_loss, "Loss_depth/byol": byol_loss,
                        "Policy/mean_noise_std": alg.actor_critic.std.mean().item()}
                log_tsc(self, self.writer, it, tags, collection_time, learn_time)
            if self.log_dir is not None and it % self.save_interval == 0:
                self.save(os.path.join(self.log_dir, "model.pt"))
        self.current_learning_iteration = tot_iter

    # ---- inference handles (:662-703) -----------------------------------------------------------------------------------
    def _eval(self, module, device):
        module.eval()
        if device is not None:
            module.to(device)
        return module

    @staticmethod
    def s_to_hms(seconds):                                                  # :603-608
        return seconds // 3600, (seconds % 3600) // 60, seconds % 60

    def get_disc_inference_policy(self, device=None):
        """The style discriminator in eval mode (the reference's method reads an attribute that does not exist, :699-703)."""
        return self._eval(self.discriminator, device)

    def get_inference_policy(self, device=None):
        return self._eval(self.alg.actor_critic, device).act_inference

    def get_inference_policy_bbc(self, device=None):
        return self._eval(self.alg.actor_critic_bbc, device).act_inference

    def get_depth_actor_inference_policy(self, device=None):
        return self._eval(self.alg.depth_actor, device)

    def get_actor_critic(self, device=None):
        return self._eval(self.alg.actor_critic, device)

    def get_estimator_inference_policy(self, device=None):
        est = self._eval(self.alg.estimator, device)
        return est.inference if hasattr(est, "inference") else est

    def get_depth_encoder_inference_policy(self, device=None):
        return self._eval(self.alg.depth_encoder, device)

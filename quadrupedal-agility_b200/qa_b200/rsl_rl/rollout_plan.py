"""`RolloutPlan` -- the trainer's half of one rollout step (bbc/rsl_rl/runners/on_policy_runner.py:156-181) as a static
schedule of libqa_b200 launches on preallocated buffers, the inference twin of `ppo_plan.PpoStepPlan`:

  act()     `SSInfoGAIL.act` (gail.py:176-197): estimator -> explicit lanes, privileged-latent (or history) encoder -> latent
            lanes, actor trunk + head, critic trunk + head, Normal sample + log-prob -- and the transition's storage writes
            (`RolloutStorage.add_transitions`, rollout_storage.py:58-76) folded into the producing kernels: the observation rows
            are copied into their storage slots by the same launch that assembles the actor's input row, the critic reads its
            input from the storage slot and writes the value into it, the sample kernel writes actions / log-prob / mu / sigma.
  reward()  the discriminator reward of the step (algorithms/discriminator.py:71-118) + `process_env_step` (gail.py:199-212):
            K18 history bookkeeping + normalisation, two tcgen05 trunk layers, ONE head launch over the three heads, K19.

The four independent chains (estimator | encoder | critic, then the actor) run on side streams; captured into the rollout's
CUDA graph they are parallel branches.  At 4096 rows every launch is latency bound, so what counts is the number of DEPENDENT
launches on the critical path (21 per env step, was ~45) and that no framework kernel sits between them.
"""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from .ppo_plan import PpoStepPlan, _Chain, _Cur, _linears, _padded


class RolloutPlan:
    @staticmethod
    def supported(alg) -> Optional[str]:
        why = PpoStepPlan.supported(alg)
        if why is not None:
            return why
        d = alg.disc
        if d is None or _linears(d.trunk) is None or not isinstance(d.trunk[1], nn.ReLU):
            return "discriminator trunk is not a Linear/ReLU stack"
        n_heads = d.linear.out_features + d.encoder_eps.out_features + d.classifier.out_features
        if n_heads > 16 or d.linear.in_features not in (32, 64, 128, 256):
            return "discriminator heads outside K20"
        if alg.disc_loss_function != "MSELoss" or alg.disc_normalizer is None:
            return "fused reward tail needs the MSE style reward and a normaliser"
        if alg.storage is None or alg.storage.privileged_observations is None:
            return "no storage"
        if alg.storage.observations.stride(1) % 4 != 0:
            return "storage rows are not TMA operands"
        return None

    def __init__(self, alg, env, disc_obs_len, obs_disc_weight_step):
        self.alg, self.env = alg, env
        dev = self.dev = torch.device(alg.device)
        ac, est, st, d = alg.actor_critic, alg.estimator, alg.storage, alg.disc
        if getattr(alg, "disc_flat", None) is None:
            alg._init_disc_update()          # flat, 16-byte-pitch discriminator parameters: its trunk is read through TMA
        N = self.N = st.num_envs
        self.act_name = ac.activation_name
        self.p, self.e, self.l = alg.num_prop, alg.num_explicit, alg.num_latent
        self.h = alg.num_hist * alg.num_prop
        self.W, self.Wc, self.A = st.observations.shape[-1], st.privileged_observations.shape[-1], st.actions.shape[-1]
        self.n_in = ac.num_actor_obs
        self.n_cmd = self.n_in - self.p - self.e - self.l
        est_mods = list(est.estimator)
        self.c_est = _Chain(_linears(est_mods[:-1]), est_mods[-1], self.act_name, N, dev)
        self.c_priv = _Chain(_linears(ac.priv_encoder), None, self.act_name, N, dev)
        self.c_actor = _Chain(_linears(ac.actor_trunk), ac.actor_head, self.act_name, N, dev)
        self.c_critic = _Chain(_linears(ac.critic_trunk), ac.critic_head, self.act_name, N, dev)
        self.c_disc = _Chain(_linears(d.trunk), None, "relu", N, dev)
        for c in (self.c_est, self.c_priv, self.c_actor, self.c_critic, self.c_disc):
            c.gz = None                                             # inference: no gradient buffers
        self.xa = _padded(N, self.n_in, dev)
        self.lat_in = _padded(N, self.l, dev)
        self.rows = torch.arange(N, device=dev, dtype=torch.int64)
        self.actions = torch.zeros(N, self.A, device=dev)
        self.L, self.Wd = disc_obs_len, env.num_obs_disc
        self.obs_disc_weight_step = obs_disc_weight_step
        self.x_norm = _padded(N, self.L * self.Wd, dev)
        self.hist_pp = [torch.zeros(N, self.L, self.Wd, device=dev) for _ in range(2)]
        self.hist_new = torch.zeros(N, self.L * self.Wd, device=dev)
        n_heads = d.linear.out_features + d.encoder_eps.out_features + d.classifier.out_features
        self.heads_w = torch.zeros((n_heads + 3) // 4 * 4, d.linear.in_features, device=dev)
        self.heads_b = torch.zeros((n_heads + 3) // 4 * 4, device=dev)
        self.heads_out = torch.zeros(N, self.heads_w.shape[0], device=dev)
        self.refresh_disc_heads()
        self.s_est, self.s_priv, self.s_critic = (torch.cuda.Stream(device=dev) for _ in range(3))
        # reward tail (discriminator trunk, heads, K19) of step t on its own stream, next to the policy chain and the env kernels
        # of step t+1 (`defer_reward_tail`): what the tail reads and the next env step overwrites is snapshotted by K18
        self.s_rew = torch.cuda.Stream(device=dev)
        self.defer_reward_tail = False                  # set by callers that run WHOLE rollouts (runner.learn, BbcIteration)
        self._tail_pending = False
        self.rew_snap = torch.zeros(N, device=dev)
        self.dones_snap = torch.zeros(N, device=dev, dtype=torch.uint8)
        self.tout_snap = torch.zeros(N, device=dev, dtype=torch.uint8)
        # action noise: in-kernel Philox by default; True = draw it from torch's CUDA generator like the reference's
        # Normal.sample() (one extra framework kernel per step; what the torch-path comparison tests use)
        self.use_torch_generator = False

    def refresh_disc_heads(self) -> None:
        """[d | eps | classifier logits | 0 pad] weights as one matrix (the layout K19 reads), refreshed IN PLACE after every
        discriminator update / checkpoint load (the only writers of these parameters)."""
        d = self.alg.disc
        with torch.no_grad():
            r = 0
            for m in (d.linear, d.encoder_eps, d.classifier):
                k = m.out_features
                self.heads_w[r:r + k].copy_(m.weight)
                self.heads_b[r:r + k].copy_(m.bias)
                r += k

    # ---- act -------------------------------------------------------------------------------------------------------------
    def _trunk(self, c: _Chain, x, x_col0=0, last_y=None, last_y_col0=0):
        for i, lin in enumerate(c.trunk):
            if i == len(c.trunk) - 1 and last_y is not None:
                ops.linear_fwd(x, lin.weight, lin.bias, last_y, c.act, x_col0=x_col0, y_col0=last_y_col0)
            else:
                ops.linear_fwd(x, lin.weight, lin.bias, c.h[i], c.act, x_col0=x_col0)
            x, x_col0 = c.h[i], 0

    @torch.no_grad()
    def act(self, obs, critic_obs, hist_encoding=False, normal_draw=None):
        alg, env, st = self.alg, self.env, self.alg.storage
        ac, tr = alg.actor_critic, alg.transition
        p, e, l, h = self.p, self.e, self.l, self.h
        t = st.step
        if t >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        xa = self.xa
        cur = _Cur(True)
        on = torch.cuda.stream
        # one launch: actor input row windows + the observation rows into their storage slots (add_transitions' two big copies)
        ent = [(obs, 0, xa, 0, p + e), (obs, p + e + l + h, xa, p + e + l, self.n_cmd),
               (obs, 0, st.observations[t], 0, self.W), (critic_obs, 0, st.privileged_observations[t], 0, self.Wc)]
        if not hist_encoding:
            ent.append((obs, p + e, self.lat_in, 0, l))
        ops.gather_minibatch_windows(self.rows, ent)
        cur.fork(self.s_est)
        cur.fork(self.s_priv)
        cur.fork(self.s_critic)
        with on(self.s_critic):                                           # critic reads its storage slot, writes the value slot
            self._trunk(self.c_critic, st.privileged_observations[t])
            ops.head_fwd(self.c_critic.h[-1], ac.critic_head.weight, ac.critic_head.bias, st.values[t])
        if alg.train_with_estimated_explicit:                             # gail.py:181-184: estimated explicit lanes
            with on(self.s_est):
                self._trunk(self.c_est, xa)
                ops.head_fwd(self.c_est.h[-1], self.c_est.head.weight, self.c_est.head.bias, xa[:, p:p + e])
        with on(self.s_priv):
            if hist_encoding:                                             # actor_critic.py:179-180 / :219-220, K11
                ops.hist_encoder_fwd(obs[:, p + e + l:p + e + l + h], ac.history_encoder, xa[:, p + e:p + e + l])
            else:
                self._trunk(self.c_priv, self.lat_in, last_y=xa, last_y_col0=p + e)
        cur.join(self.s_est)
        cur.join(self.s_priv)
        self._trunk(self.c_actor, xa)
        ops.head_fwd(self.c_actor.h[-1], ac.actor_head.weight, ac.actor_head.bias, self.c_actor.out)
        dev_counter = getattr(env, "device_step_counter", False)
        if normal_draw is None and self.use_torch_generator:
            normal_draw = torch.randn(self.N, self.A, device=self.dev)
        ops.policy_sample(self.c_actor.out, ac.std.detach(), self.actions, noise=normal_draw, rng_seed=getattr(env, "seed", 0) + 0x5EED,
                          rng_step=getattr(env, "common_step_counter", 0) + 1,
                          step_state=env._step_state[0:1] if dev_counter else None,
                          actions_st=st.actions[t], logp_st=st.actions_log_prob[t].view(-1), mu_st=st.mu[t], sigma_st=st.sigma[t])
        cur.join(self.s_critic)
        tr.actions, tr.values = self.actions, st.values[t]
        tr.actions_log_prob, tr.action_mean, tr.action_sigma = st.actions_log_prob[t].view(-1), st.mu[t], st.sigma[t]
        tr.observations, tr.critic_observations = obs, critic_obs
        return self.actions

    def finish_rewards(self):
        """Joins a reward tail that is still running on its side stream (end of a rollout that does not fill the storage)."""
        if self._tail_pending:
            _Cur(True).join(self.s_rew)
            self._tail_pending = False

    # ---- reward + process_env_step ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def reward(self, obs, rewards, dones, prev_disc, next_disc, disc_hist, reward_terms=None):
        """Returns the next step's discriminator history buffer.  `disc_hist`: (N, L, Wd) history BEFORE this step."""
        alg, env, st, d = self.alg, self.env, self.alg.storage, self.alg.disc
        t = st.step
        staged = alg._disc_stage is not None
        hist_new = alg._disc_stage[0][t] if staged else self.hist_new
        dst = self.hist_pp[0] if disc_hist is not self.hist_pp[0] else self.hist_pp[1]
        mean, std = alg.disc_normalizer.device_moments(self.dev)
        time_outs = env._time_outs_latched if env.cfg.send_timeouts else None
        # Deferred tail: the trunk / heads / K19 of this step run on `s_rew` while the main stream goes on with the next step's
        # policy and env kernels.  Hazards: (1) x_norm / heads_out / the snapshots are single buffers -> the previous tail is
        # joined first (it has had a whole env step to finish); (2) rew_buf, reset_buf and the latched time-outs are overwritten
        # by the next K2 -> K18, which runs here on the main stream, copies them; (3) the observation row the policy saw (the
        # labels K19 reads) lives in a ping-pong buffer the step after next overwrites -> read from its storage slot;
        # (4) latent_eps / latent_c may be resampled by the next K2 -> the replay-buffer staging stays on the main stream.
        defer = self.defer_reward_tail and reward_terms is None
        cur = _Cur(True)
        if self._tail_pending:
            cur.join(self.s_rew)
            self._tail_pending = False
        weight = alg._task_obs_weight_dev() if (env.task_obs_weight_decay and hasattr(alg, "_task_obs_weight_dev")) else env.task_obs_weight
        snaps = (rewards, self.rew_snap, self.dones_snap, time_outs, None if time_outs is None else self.tout_snap) if defer else None
        lat = (env.latent_eps, alg._disc_stage[1][t], env.latent_c, alg._disc_stage[2][t]) if staged else None
        ops.disc_input(dones, prev_disc, next_disc, disc_hist, hist_new, dst, self.x_norm, mean, std, alg.disc_normalizer.clip_obs,
                       env.task_obs_weight_decay, weight, self.obs_disc_weight_step, snapshots=snaps, latents=lat)
        if not staged:
            alg.disc_storage.insert(hist_new, env.latent_eps, env.latent_c)
        coefs = (d.reward_i_coef, d.reward_us_coef, d.reward_ss_coef, d.reward_t_coef)
        if defer:
            cur.fork(self.s_rew)
            with torch.cuda.stream(self.s_rew):
                self._trunk(self.c_disc, self.x_norm)
                ops.head_fwd(self.c_disc.h[-1], self.heads_w, self.heads_b, self.heads_out)
                ops.disc_reward(self.heads_out, st.observations[t], self.rew_snap, d.dt, coefs, st.rewards[t].view(-1),
                                values=st.values[t], time_outs=None if time_outs is None else self.tout_snap, gamma=alg.gamma,
                                dones=self.dones_snap, dones_out=st.dones[t].view(-1))
            self._tail_pending = True
            if t + 1 >= st.num_transitions_per_env:          # last step of the rollout: GAE reads rewards / dones next
                cur.join(self.s_rew)
                self._tail_pending = False
        else:
            self._trunk(self.c_disc, self.x_norm)
            ops.head_fwd(self.c_disc.h[-1], self.heads_w, self.heads_b, self.heads_out)
            ops.disc_reward(self.heads_out, obs, rewards, d.dt, coefs, st.rewards[t].view(-1), values=st.values[t],
                            time_outs=time_outs, gamma=alg.gamma, dones=dones, dones_out=st.dones[t].view(-1),
                            reward_terms=reward_terms)
        alg.transition.dones = dones
        st.step += 1
        alg.actor_critic.reset(dones)
        return dst

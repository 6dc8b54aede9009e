"""Runner bookkeeping of bbc/rsl_rl/runners/on_policy_runner.py:183-206 (per-step episode accounting) and :238-304 (`log`),
without the reference's per-step device->host traffic.

The reference, every env step, adds the step's rewards to per-env running sums, pulls the sums of the envs that finished
through `.cpu().numpy().tolist()` into six `deque(maxlen=100)`s and zeroes them (6 syncs per step, 144 per iteration).
Here the step's reward terms and done mask are copied (device->device, stream ordered) into a `(T, N, C)` / `(T, N)`
staging area; `flush()` does ONE device->host copy per iteration and replays the reference's accounting on the host in the
same order (steps in order, finished envs in ascending env index within a step), so the deques hold exactly what the
reference's hold.

`ScalarLog` is the writer: TensorBoard's `SummaryWriter` when `log_dir` is given and tensorboard imports, and always an
in-memory `{tag: [(step, value)]}` record (what the tests read).
"""
import statistics
from collections import deque

import numpy as np
import torch


class EpisodeBook:
    def __init__(self, num_envs, num_steps, term_names, device, maxlen=100, num_episode_keys=0, snap_names=()):
        """`term_names`: per-step quantities SUMMED over an episode (rewards); `snap_names`: per-step quantities whose value
        AT the finishing step is recorded (e.g. the TSC runner's `reach_goal`, tsc on_policy_runner.py:248)."""
        self.N, self.T, self.names, self.snap_names = num_envs, num_steps, tuple(term_names), tuple(snap_names)
        self.n_sum = len(self.names)
        C = len(self.names) + len(self.snap_names)
        self.terms = torch.zeros(num_steps, num_envs, C, device=device)
        self.dones = torch.zeros(num_steps, num_envs, device=device, dtype=torch.uint8)
        self.ep_means = torch.zeros(num_steps, max(num_episode_keys, 1), device=device)
        # infos['episode'] only exists once some env has been reset (legged_robot.py:188-189, 230-234), so the reference appends
        # nothing before that (on_policy_runner.py:185-186): per-step validity flag, latched on the device
        self.ep_valid = torch.zeros(num_steps, device=device)
        self._ever_reset = torch.zeros((), device=device)
        self.buffers = {n: deque(maxlen=maxlen) for n in self.names + self.snap_names}
        self.len_buffer = deque(maxlen=maxlen)
        self._cur = np.zeros((num_envs, C), dtype=np.float32)          # cur_reward_* (:141-146), fp32 like the reference
        self._len = np.zeros(num_envs, dtype=np.float32)
        self._t = 0
        pin = torch.device(device).type == "cuda"
        self._h_terms = torch.zeros(num_steps, num_envs, C, pin_memory=pin)
        self._h_dones = torch.zeros(num_steps, num_envs, dtype=torch.uint8, pin_memory=pin)
        self._h_ep = torch.zeros(num_steps, max(num_episode_keys, 1), pin_memory=pin)
        self._h_valid = torch.zeros(num_steps, pin_memory=pin)
        self.episode_rows = []                                          # ep_infos of the current iteration (one row per step)

    def term_slot(self, t=None):
        """(N, C) view a kernel may write the step's reward terms into directly."""
        return self.terms[self._t if t is None else t]

    def record(self, dones, terms=None, episode_means=None, num_resets=None):
        """One env step: `terms` (N, C) or None when already written through `term_slot()`.  `num_resets`: the env's device
        count of resets in this step (any dtype); without it the step's `dones` decide whether an episode info exists yet."""
        t = self._t
        if terms is not None:
            self.terms[t].copy_(terms)
        self.dones[t].copy_(dones)
        if episode_means is not None:
            k = min(self.ep_means.shape[1], episode_means.numel())
            self.ep_means[t, :k].copy_(episode_means.reshape(-1)[:k])
            any_reset = (num_resets.reshape(-1)[0] > 0) if num_resets is not None else dones.any()
            torch.maximum(self._ever_reset, any_reset.to(self._ever_reset.dtype), out=self._ever_reset)
            self.ep_valid[t].copy_(self._ever_reset)
        self._t = t + 1

    def flush(self):
        """One D2H of the iteration's staging area, then the reference's accounting (:190-206) step by step."""
        n = self._t
        if n == 0:
            return
        self._h_terms[:n].copy_(self.terms[:n], non_blocking=True)
        self._h_dones[:n].copy_(self.dones[:n], non_blocking=True)
        self._h_ep[:n].copy_(self.ep_means[:n], non_blocking=True)
        self._h_valid[:n].copy_(self.ep_valid[:n], non_blocking=True)
        if self.terms.is_cuda:
            torch.cuda.current_stream().synchronize()
        terms, dones = self._h_terms[:n].numpy(), self._h_dones[:n].numpy()
        self.episode_rows = [self._h_ep[i].numpy().copy() for i in range(n) if self._h_valid[i] > 0]
        k = self.n_sum
        for t in range(n):
            self._cur[:, :k] += terms[t][:, :k]
            self._cur[:, k:] = terms[t][:, k:]
            self._len += 1
            ids = np.nonzero(dones[t])[0]
            if ids.size:
                for c, name in enumerate(self.names + self.snap_names):
                    self.buffers[name].extend(self._cur[ids, c].tolist())
                self.len_buffer.extend(self._len[ids].tolist())
                self._cur[ids] = 0
                self._len[ids] = 0
        self._t = 0

    def means(self):
        """{name: statistics.mean(deque)} + 'episode_length'; empty when no episode has finished yet (:279)."""
        if len(self.len_buffer) == 0:
            return {}
        out = {n: statistics.mean(b) for n, b in self.buffers.items()}
        out["episode_length"] = statistics.mean(self.len_buffer)
        return out


class ScalarLog:
    def __init__(self, log_dir=None, writer=None):
        self.scalars = {}
        self.writer = writer
        if writer is None and log_dir is not None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=log_dir, flush_secs=10)
            except Exception:                                           # tensorboard not installed: keep the in-memory record
                self.writer = None

    def add_scalar(self, tag, value, step):
        value = float(value)
        self.scalars.setdefault(tag, []).append((int(step), value))
        if self.writer is not None:
            self.writer.add_scalar(tag, value, step)

    def last(self, tag):
        return self.scalars[tag][-1][1]


BBC_LOSS_TAGS = ("surrogate_loss", "value_loss", "b_loss", "entropy_batch", "priv_reg_loss", "estimator_loss")
BBC_DISC_TAGS = ("ss_loss", "info_max_loss", "disc_loss", "us_loss", "grad_pen_loss", "disc_logit_loss", "disc_weight_decay")
BBC_ACC_TAGS = ("acc_lb", "acc_pi", "acc_exp", "acc_ulb")            # order of the tuple `SSInfoGAIL.update` returns (:322-326)


def log_bbc(runner, log, it, stats, hist_latent_loss, collection_time, learn_time):
    """`OnPolicyRunner.log` (:238-304): same tags, same values.  `stats` = the tuple `alg.update()` returned (6 PPO means,
    then the 11 discriminator means when the discriminator was updated)."""
    env, alg, book = runner.env, runner.alg, runner.book
    rows = book.episode_rows
    if rows:
        mean_per_key = np.mean(np.stack(rows, axis=0), axis=0)           # torch.mean over the concatenated infos (:245-254)
        for i, name in enumerate(env.reward_names):
            log.add_scalar("Episode/rew_" + name, mean_per_key[i] / env.reward_scales[name], it)
    for tag, v in zip(BBC_LOSS_TAGS, stats[:6]):
        log.add_scalar("Loss/" + tag, v, it)
    if hist_latent_loss is not None:
        log.add_scalar("Loss/hist_latent_loss", hist_latent_loss, it)
    if len(stats) >= 17:
        for tag, v in zip(BBC_DISC_TAGS, stats[6:13]):
            log.add_scalar("Loss/" + tag, v, it)
        for tag, v in zip(BBC_ACC_TAGS, stats[13:17]):
            log.add_scalar("Acc/" + tag, v, it)
    log.add_scalar("Loss/mean_noise_std", alg.actor_critic.std.mean().item(), it)
    log.add_scalar("LR/lr_ac", alg.lr_ac, it)
    log.add_scalar("LR/lr_disc", alg.lr_disc, it)
    log.add_scalar("LR/lr_q", alg.lr_q, it)
    fps = int(runner.num_steps_per_env * env.num_envs / (collection_time + learn_time))
    log.add_scalar("Perf/total_fps", fps, it)
    log.add_scalar("Perf/collection time", collection_time, it)
    log.add_scalar("Perf/learning_time", learn_time, it)
    m = book.means()
    if m:
        log.add_scalar("Train/mean_reward", m["total"], it)
        for k in ("i", "us", "ss", "t"):
            log.add_scalar("Train/mean_reward_" + k, m[k], it)
        log.add_scalar("Train/mean_episode_length", m["episode_length"], it)
    return m


def log_tsc(runner, log, it, tags, collection_time, learn_time):
    """The scalar groups of tsc/rsl_rl/runners/on_policy_runner.py `log` / `log_vision` (:462-594): `tags` = {tag: value} of the
    iteration's loss statistics; episode means, throughput and the bookkeeping means are added here."""
    env, book = runner.env, runner.book
    rows = book.episode_rows
    if rows:
        mean_per_key = np.mean(np.stack(rows, axis=0), axis=0)
        for i, name in enumerate(env.cfg.reward_names):
            log.add_scalar("Episode_rew/rew_" + name, mean_per_key[i], it)
    for tag, v in tags.items():
        log.add_scalar(tag, v, it)
    fps = int(runner.num_steps_per_env * env.num_envs / (collection_time + learn_time))
    log.add_scalar("Perf/total_fps", fps, it)
    log.add_scalar("Perf/collection time", collection_time, it)
    log.add_scalar("Perf/learning_time", learn_time, it)
    m = book.means()
    if m:
        log.add_scalar("Train/mean_reward", m["total"], it)
        for k in ("i", "us", "ss", "t"):
            if k in m:
                log.add_scalar("Train/mean_reward_" + k, m[k], it)
        log.add_scalar("Train/mean_episode_length", m["episode_length"], it)
        if "reach_goal" in m:
            log.add_scalar("Train/mean_success_rate", m["reach_goal"], it)
    return m

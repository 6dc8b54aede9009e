"""`SSInfoGAIL` -- the PPO half of the reference algorithm (bbc/rsl_rl/algorithms/gail.py:12-413) re-designed
for one B200 per process:

  * `act` / `process_env_step` / `compute_returns` keep the reference's signatures (:176-217);
  * `update()` runs the 20 PPO minibatch steps of `update_actor_critic` (:328-413) with NO host round trip:
    minibatches are gathered by one kernel (K6) into static buffers, the forward/backward of the whole step
    is a replayed CUDA graph, the adaptive-KL learning rate (:368-379) lives in a device scalar, gradient
    clipping + Adam are one fused pass over the flat parameter buffer (K8), and the per-minibatch loss
    statistics are accumulated on the device and read back once per update (the reference calls `.item()`
    six times per minibatch, :277-282);
  * multi-GPU: envs are sharded one shard per rank; the only collectives are an in-place NCCL all-reduce of
    the flat gradient buffers and of the scalar KL, both before the fused optimiser pass.

The semi-supervised InfoGAIL discriminator update (`update_ss_info_gail`, :415-541) is provided
with the reference's signature: flat discriminator parameters, the reference's three weight-decayed Adam optimisers on
K8, the double-backward gradient penalty on torch ops, the running normaliser merged on the device instead of through
numpy; `update_disc` drives it over the replay buffer and the expert sets.  DAgger (`update_dagger`, :543-575) fits the
history encoder on its slice of the flat buffer.
"""
import os
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .. import ops
from .. import dist as qdist
from . import checkpoint as ckpt
from .storage import RolloutStorage


class ReplayBuffer:
    """Fixed-size ring of discriminator observations (storage/replay_buffer.py:6-48); filled every env step
    by `process_env_step` (gail.py:207-208), consumed by the discriminator update."""

    def __init__(self, obs_dim, dim_c, history_len, buffer_size, device):
        self.states = torch.zeros(buffer_size, history_len * obs_dim, device=device)
        self.latent_eps = torch.zeros(buffer_size, 1, device=device)
        self.latent_c = torch.zeros(buffer_size, dim_c, device=device)
        self.buffer_size, self.device = buffer_size, device
        self.step = 0
        self.num_samples = 0

    def insert(self, states, latent_eps, latent_c):
        n = states.shape[0]
        first = min(n, self.buffer_size - self.step)
        for dst, src in ((self.states, states), (self.latent_eps, latent_eps), (self.latent_c, latent_c)):
            dst[self.step:self.step + first].copy_(src[:first])
            if first < n:
                dst[:n - first].copy_(src[first:])
        self.num_samples = min(self.buffer_size, max(self.step + n, self.num_samples))
        self.step = (self.step + n) % self.buffer_size

    def feed_forward_generator(self, num_mini_batch, mini_batch_size):
        """storage/replay_buffer.py:44-50 with the index draw (np.random.choice with replacement there) on the device."""
        for _ in range(num_mini_batch):
            idx = torch.randint(self.num_samples, (mini_batch_size,), device=self.states.device)
            yield self.states[idx], self.latent_eps[idx], self.latent_c[idx]


class FlatAdam:
    """Adam state for one flat parameter buffer (or a contiguous slice [lo, hi) of it); `step()` = clip_grad_norm_ +
    Adam.step through K8.  `weight_decay` follows torch.optim.Adam (added to the gradient)."""

    def __init__(self, flat, lr: float, max_grad_norm: float, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, lo=0, hi=None):
        dev = flat.data.device
        self.flat = flat
        hi = flat.data.numel() if hi is None else hi
        self.lo, self.hi = lo, hi
        self.exp_avg = torch.zeros(hi - lo, device=dev)
        self.exp_avg_sq = torch.zeros(hi - lo, device=dev)
        self.lr = torch.full((1,), lr, device=dev, dtype=torch.float32)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.ws = torch.zeros(2, device=dev, dtype=torch.float64)
        self.betas, self.eps, self.max_grad_norm, self.weight_decay = betas, eps, max_grad_norm, weight_decay

    def step(self, grad_scale: float = 1.0, presummed: bool = False):
        """`presummed`: K31 has already left sum((grad * grad_scale)^2) in `ws` and incremented `step_count`."""
        fn = ops.adam_apply if presummed else ops.clip_adam
        fn(self.flat.data[self.lo:self.hi], self.flat.grad[self.lo:self.hi], self.exp_avg, self.exp_avg_sq, self.lr,
           self.step_count, self.ws, self.betas[0], self.betas[1], self.eps, self.max_grad_norm, grad_scale,
           self.grad_norm, self.weight_decay)

    # torch.optim-compatible state for checkpoints
    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr, "step": self.step_count}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr.copy_(sd["lr"])
        self.step_count.copy_(sd["step"])


class _PPOLossFused(torch.autograd.Function):
    """K10: the PPO loss terms of gail.py:367-408, forward and backward in one kernel launch."""

    @staticmethod
    def forward(ctx, mu, value, std, mb, cfg, stats):
        M = mu.shape[0]
        dmu = torch.empty(M, mu.shape[1], device=mu.device, dtype=torch.float32)
        dvalue = torch.empty(M, device=mu.device, dtype=torch.float32)
        dstd = torch.empty(std.shape[0], device=mu.device, dtype=torch.float32)
        ops.ppo_loss(mu, std.detach().contiguous(), value, mb["actions"], mb["old_actions_log_prob"].view(-1),
                     mb["advantages"].view(-1), mb["returns"].view(-1), mb["values"].view(-1), mb["old_mu"],
                     mb["old_sigma"], dmu, dvalue, dstd, stats, cfg["clip"], cfg["c_surr"], cfg["c_value"],
                     cfg["c_bound"], cfg["c_entropy"], cfg["clipped_value"])
        ctx.save_for_backward(dmu, dvalue, dstd)
        # the VALUE of the combined loss is never read (the logged statistics are its terms, accumulated by K13); the
        # returned scalar only roots the backward pass
        return stats[0] * cfg["c_surr"]

    @staticmethod
    def backward(ctx, g):
        dmu, dvalue, dstd = ctx.saved_tensors
        return g * dmu, (g * dvalue).unsqueeze(1), g * dstd, None, None, None


class _RowLossFused(torch.autograd.Function):
    """K12: estimator MSE (gail.py:359) / privileged-latent regulariser (:354), loss and gradient in one launch.
    `b` is a constant (the reference detaches it / computes it under no_grad)."""

    @staticmethod
    def forward(ctx, a, b, mode, out):
        M, W = a.shape
        da = torch.empty(M, (W + 3) // 4 * 4, device=a.device, dtype=torch.float32)[:, :W]
        ops.row_loss(a, b, da, out, mode)
        ctx.save_for_backward(da)
        return out[0] * 1.0

    @staticmethod
    def backward(ctx, g):
        (da,) = ctx.saved_tensors
        return da.mul_(g), None, None, None


STAT_NAMES = ("surrogate_loss", "value_loss", "b_loss", "entropy", "priv_reg_loss", "estimator_loss", "kl_mean")


class SSInfoGAIL:
    def __init__(self, env, actor_critic, discriminator, estimator, estimator_paras, motion_loader, disc_normalizer,
                 disc_history_len, disc_obs_len, num_disc_obs, obs_disc_weight_step, disc_loss_function=None,
                 num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 surrogate_loss_coef=1., value_loss_coef=5., entropy_coef=0., bounds_loss_coef=10., disc_coef=5.,
                 disc_logit_reg=0.05, disc_grad_penalty=0.2, disc_weight_decay=0.0001, lr_ac=1e-3, lr_disc=1e-3,
                 lr_q=1e-3, max_grad_norm=1.0, use_clipped_value_loss=False, schedule="fixed", desired_kl=0.01,
                 device='cpu', disc_replay_buffer_size=100000, min_std=None, us_coef=1.0, ss_coef=4.0,
                 prior_soft_coef=1e-3, info_max_coef=2.0, begin_rim=100, priv_reg_coef_schedual=[0, 0.1, 0, 1],
                 priv_reg_coef_schedual_resume=[0, 0.1, 0, 1], use_cuda_graph=True, fused_loss=True, capture_collectives=None):
        self.device, self.env = device, env
        self.desired_kl, self.schedule = desired_kl, schedule
        self.lr_disc, self.lr_q, self.min_std = lr_disc, lr_q, min_std
        self.disc_coef, self.disc_logit_reg, self.disc_grad_penalty = disc_coef, disc_logit_reg, disc_grad_penalty
        self.disc_weight_decay, self.us_coef, self.ss_coef = disc_weight_decay, us_coef, ss_coef
        self.prior_soft_coef, self.info_max_coef, self.begin_rim = prior_soft_coef, info_max_coef, begin_rim
        self.dim_c = env.dim_c
        self.info_max_coef_on = 0.0                                 # gail.py:143; ramps up after `begin_rim` updates (:251-253)
        self.disc_loss_function = disc_loss_function
        self.disc_history_len, self.disc_obs_len, self.num_disc_obs = disc_history_len, disc_obs_len, num_disc_obs
        self.obs_disc_weight_step = obs_disc_weight_step
        self.disc = discriminator.to(device) if discriminator is not None else None
        self.disc_storage = ReplayBuffer(env.num_obs_disc, self.dim_c, disc_obs_len, disc_replay_buffer_size, device)
        self.motion_loader, self.disc_normalizer = motion_loader, disc_normalizer
        cfg_env = env.cfg.env if hasattr(env.cfg, "env") else None
        g = (lambda k, d: getattr(cfg_env, k, d)) if cfg_env is not None else (lambda k, d: d)
        self.num_prop, self.num_explicit = g("num_prop", 57), g("num_explicit", 4)
        self.num_latent, self.num_hist, self.num_command = g("num_latent", 29), g("history_len", 10), g("num_command", 11)

        self.actor_critic = actor_critic.to(device)
        self.estimator = estimator.to(device)
        self.ac_flat = self.actor_critic.flatten_parameters()
        self.est_flat = self.estimator.flatten_parameters()
        self.optim_ac = FlatAdam(self.ac_flat, lr_ac, max_grad_norm)
        self.optim_estimator = FlatAdam(self.est_flat, estimator_paras["learning_rate"], max_grad_norm)
        self._lr_estimator0 = float(estimator_paras["learning_rate"])
        self.train_with_estimated_explicit = estimator_paras["train_with_estimated_explicit"]
        self.priv_reg_coef_schedual = priv_reg_coef_schedual
        self.priv_reg_counter = 0
        self.transition: Optional[RolloutStorage.Transition] = None
        self.storage: Optional[RolloutStorage] = None

        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.surrogate_loss_coef, self.value_loss_coef = surrogate_loss_coef, value_loss_coef
        self.entropy_coef, self.bounds_loss_coef = entropy_coef, bounds_loss_coef
        self.gamma, self.lam, self.max_grad_norm = gamma, lam, max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.learning_steps = 0
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.use_cuda_graph = use_cuda_graph and torch.device(device).type == "cuda"
        self._graphs = None
        self.fused_loss = fused_loss and torch.device(device).type == "cuda"
        # with > 1 rank the three NCCL all-reduces are captured inside the minibatch graph (measured: 95 % weak-scaling
        # efficiency at 2 GPUs); QA_CAPTURE_COLLECTIVES=0 keeps only forward/backward in the graph and runs the
        # all-reduces + K13/K8 eagerly between replays
        if capture_collectives is None:
            capture_collectives = os.environ.get("QA_CAPTURE_COLLECTIVES", "1") == "1"
        self.capture_collectives = capture_collectives
        self._graph_has_apply = True
        # north_star's "single NCCL allreduce of PPO gradients": the actor-critic gradients, the estimator gradients and the KL
        # scalar live in ONE arena reduced by one collective per optimiser step (QA_SINGLE_ALLREDUCE=0: three collectives);
        # gloo: tests/test_dist_gloo.py, NCCL on 2 x B200: tests/test_dist_nccl_gpu.py
        self._grad_arena = None
        if self.world_size > 1 and os.environ.get("QA_SINGLE_ALLREDUCE", "1") == "1":
            self.use_grad_arena()
        # discriminator minibatch step with one shared forward (see update_ss_info_gail); opt-in until it has been measured
        # and re-pinned on the B200 (same values up to the summation order of the weight gradients)
        self.disc_batched = os.environ.get("QA_DISC_BATCHED", "0") == "1"
        # PPO minibatch step with ONE privileged-latent encoder pass (see _forward_backward); opt-in for the same reason
        self.share_priv_latent = os.environ.get("QA_SHARE_PRIV_LATENT", "0") == "1"
        # the PPO minibatch step as a static schedule of libqa_b200 launches (ppo_plan.PpoStepPlan) instead of an autograd graph:
        # the default on CUDA with the tcgen05 layers; QA_PPO_PLAN=0 (or linear mode "fp32", the cuBLAS parity mode) keeps autograd
        self.use_plan = os.environ.get("QA_PPO_PLAN", "1") == "1"
        self._plan = None
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)   # capture runs on a side stream
        self._ppo_stats = torch.zeros(4, device=device)
        self._aux_loss = torch.zeros(2, device=device)            # priv_reg_loss, estimator_loss of the current minibatch
        self._priv_reg_coef = torch.zeros((), device=device)
        self._stats = torch.zeros(len(STAT_NAMES), device=device)
        self.last_stats = {}
        self._disc_stage = None

    def use_grad_arena(self):
        """[actor-critic grads | estimator grads | kl, pad] in one contiguous buffer; `_allreduce_grads` then needs one collective."""
        n_ac, n_est = self.ac_flat.numel, self.est_flat.numel
        n = (n_ac + n_est + 4 + 3) // 4 * 4
        self._peer = None
        if torch.device(self.device).type == "cuda" and qdist.PeerArena.available():
            # every rank's arena mapped into every process (cudaIpc, NVLink peers): the all-reduce is K31, one kernel that also
            # leaves the gradient norms for K8 -- no NCCL collective on the optimiser step (QA_PEER_ALLREDUCE=0: NCCL)
            self._peer = qdist.PeerArena(n, self.device)
            arena = self._peer.tensor
        else:
            arena = torch.zeros(n, device=self.device)
        self.ac_flat.rebind_grad(arena[:n_ac])
        self.est_flat.rebind_grad(arena[n_ac:n_ac + n_est])
        self._kl = arena[n_ac + n_est]
        self._grad_arena = arena

    def _allreduce_grads(self) -> float:
        """SUM of the flat gradients over the ranks (the 1/W goes into K8's grad_scale) and the rank-mean of the KL."""
        if self._grad_arena is not None:
            peer = getattr(self, "_peer", None)
            if peer is not None:
                n_ac, n_est = self.ac_flat.numel, self.est_flat.numel
                scale = 1.0 / self.world_size
                ops.peer_allreduce(peer.world_size, peer.rank, peer.n, peer.arena_ptrs, peer.ctrl_ptrs, seg_split=n_ac,
                                   norm_end=n_ac + n_est, sumsq_out=(self.optim_ac.ws, self.optim_estimator.ws), grad_scale=scale,
                                   step_inc=(self.optim_ac.step_count, self.optim_estimator.step_count), scale_index=n_ac + n_est)
                return scale
            scale = qdist.allreduce_flat_(self._grad_arena)
            self._kl.mul_(scale)
            return scale
        scale = qdist.allreduce_flat_(self.ac_flat.grad)
        qdist.allreduce_flat_(self.est_flat.grad)
        qdist.allreduce_mean_scalar_(self._kl)
        return scale

    @staticmethod
    def weighted_bce_loss(predictions, targets, weights):                     # gail.py:219-223
        return (weights * F.binary_cross_entropy_with_logits(predictions, targets, reduction='none')).mean()

    @staticmethod
    def weighted_mse_loss(predictions, targets, weights):                     # gail.py:225-229
        return (weights * F.mse_loss(predictions, targets, reduction='none')).mean()

    def stage_disc_inserts(self, on: bool = True):
        """Replay-buffer inserts of a rollout go to a (T,N,.) staging area at fixed addresses (so that the rollout can
        be a replayed CUDA graph) and are appended to the ring by `flush_disc_stage()` after the rollout."""
        if not on:
            self._disc_stage = None
            return
        st, ds = self.storage, self.disc_storage
        T, N = st.num_transitions_per_env, st.num_envs
        z = lambda w: torch.zeros(T, N, w, device=self.device)                  # noqa: E731
        self._disc_stage = (z(ds.states.shape[1]), z(1), z(ds.latent_c.shape[1]))

    def flush_disc_stage(self):
        if self._disc_stage is not None:
            s, e, c = self._disc_stage
            self.disc_storage.insert(s.flatten(0, 1), e.flatten(0, 1), c.flatten(0, 1))

    # ---- reference API --------------------------------------------------------------------------------
    @property
    def lr_ac(self) -> float:
        return float(self.optim_ac.lr.item())

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.transition = RolloutStorage.Transition(num_envs, actor_obs_shape, critic_obs_shape, action_shape, self.device)
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape,
                                      action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    @torch.no_grad()
    def act(self, obs, critic_obs, hist_encoding=False, normal_draw=None):
        """gail.py:176-197: estimator overwrite of the explicit privileged lanes, policy sample, value."""
        tr, ac = self.transition, self.actor_critic
        if self.train_with_estimated_explicit:
            obs_est = obs.clone()
            obs_est[:, self.num_prop:self.num_prop + self.num_explicit] = self.estimator(obs_est[:, :self.num_prop])
            tr.actions = ac.act(obs_est, hist_encoding, normal_draw=normal_draw)
        else:
            tr.actions = ac.act(obs, hist_encoding, normal_draw=normal_draw)
        tr.values = ac.evaluate(critic_obs)
        tr.actions_log_prob = ac.get_actions_log_prob(tr.actions)
        tr.action_mean, tr.action_sigma = ac.action_mean, ac.action_std
        tr.observations, tr.critic_observations = obs, critic_obs
        return tr.actions

    @torch.no_grad()
    def process_env_step(self, rewards, dones, infos, obs_disc_history_buf=None):
        """gail.py:199-212 (time-out bootstrap happens in the dtype of `rewards`, float64 when they come from
        predict_disc_reward; the storage copy rounds to fp32)."""
        tr = self.transition
        tr.rewards = rewards.clone()
        tr.dones = dones
        if 'time_outs' in infos:
            tr.rewards += self.gamma * torch.squeeze(tr.values * infos['time_outs'].unsqueeze(1).to(self.device), 1)
        if obs_disc_history_buf is not None:
            flat_hist = obs_disc_history_buf.reshape(obs_disc_history_buf.shape[0], -1)
            if self._disc_stage is not None:          # CUDA-graph rollouts: fixed-address staging, flushed per iteration
                t = self.storage.step
                self._disc_stage[0][t].copy_(flat_hist)
                self._disc_stage[1][t].copy_(self.env.latent_eps)
                self._disc_stage[2][t].copy_(self.env.latent_c)
            else:
                self.disc_storage.insert(flat_hist, self.env.latent_eps, self.env.latent_c)
        self.storage.add_transitions(tr)
        self.actor_critic.reset(dones)

    @torch.no_grad()
    def process_env_step_fused(self, heads, obs, reward_t, dones, infos, flat_hist=None, reward_terms=None):
        """`predict_disc_reward`'s tail + `process_env_step` (discriminator.py:90-118, gail.py:199-212) with the reward
        arithmetic in ONE kernel (K19) that writes storage.rewards[step] / storage.dones[step] in place.  `heads` =
        `Discriminator.heads_forward` of the normalised history; `flat_hist` = that history (N,98) when it is not already
        in the staging slot of this step."""
        tr, st, d = self.transition, self.storage, self.disc
        t = st.step
        tr.dones = dones
        ops.disc_reward(heads, obs, reward_t, d.dt, (d.reward_i_coef, d.reward_us_coef, d.reward_ss_coef, d.reward_t_coef),
                        st.rewards[t].view(-1), values=tr.values, time_outs=infos.get('time_outs'),
                        gamma=self.gamma, dones=dones, dones_out=st.dones[t].view(-1), reward_terms=reward_terms)
        if self._disc_stage is not None:
            if flat_hist is not None:
                self._disc_stage[0][t].copy_(flat_hist)
            self._disc_stage[1][t].copy_(self.env.latent_eps)
            self._disc_stage[2][t].copy_(self.env.latent_c)
        elif flat_hist is not None:
            self.disc_storage.insert(flat_hist, self.env.latent_eps, self.env.latent_c)
        st.add_transitions(tr, rewards_dones_written=True)
        self.actor_critic.reset(dones)

    @torch.no_grad()
    def compute_returns(self, last_critic_obs):
        last_values = self.actor_critic.evaluate(last_critic_obs)
        self.storage.compute_returns(last_values, self.gamma, self.lam)

    # ---- DAgger: history-encoder adaptation (gail.py:543-575) --------------------------------------------------------
    def update_dagger(self, indices: Optional[torch.Tensor] = None):
        """Fits the history encoder to the (frozen) privileged-latent encoder over the stored rollout: per minibatch
        loss = mean ||priv_latent - hist_latent||_2 (K12 forward+backward), clip_grad_norm_ over the encoder's parameters
        and Adam through K8 on the encoder's slice of the flat buffer.  The privileged latents do not change during the
        update, so they are computed once for the whole rollout.  Returns the mean loss (one host sync)."""
        st, ac = self.storage, self.actor_critic
        p, e, l, h = self.num_prop, self.num_explicit, self.num_latent, self.num_hist * self.num_prop
        opt = self._ensure_hist_optimizer()
        flat_obs = st.observations.flatten(0, 1)
        batch = flat_obs.shape[0]
        mb_size = batch // self.num_mini_batches
        if indices is None:
            indices = torch.randperm(self.num_mini_batches * mb_size, device=self.device)
        with torch.no_grad():
            priv_all = ac.infer_priv_latent(flat_obs[:, p + e:p + e + l])
        total = torch.zeros((), device=self.device)
        fused = self.fused_loss
        for _ in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                idx = indices[i * mb_size:(i + 1) * mb_size]
                obs_hist = flat_obs[idx][:, p + e + l:p + e + l + h]
                hist_latent = ac.infer_hist_latent(obs_hist)                     # with grad: torch modules (Linear / Conv1d)
                priv = priv_all[idx]
                if fused:
                    loss = _RowLossFused.apply(hist_latent, priv, 1, self._dagger_loss)
                else:
                    loss = (priv - hist_latent).norm(p=2, dim=1).mean()
                self.ac_flat.grad[opt.lo:opt.hi].zero_()
                loss.backward()
                opt.step(qdist.allreduce_flat_(self.ac_flat.grad[opt.lo:opt.hi]))       # env shards: rank-mean gradient (1/W in K8)
                total += loss.detach()
        n = self.num_learning_epochs * self.num_mini_batches
        st.clear()
        self.priv_reg_counter += 1
        return float(total.item()) / n

    def _ensure_hist_optimizer(self):
        """gail.py:99-100: Adam over the history encoder's parameters = a slice of the actor-critic's flat buffer."""
        if getattr(self, "optim_hist_encoder", None) is None:
            names = [n for n in self.ac_flat.slices if n.startswith("history_encoder.")]
            lo = min(self.ac_flat.slices[n][0] for n in names)
            hi = max(self.ac_flat.slices[n][0] + self.ac_flat.slices[n][1] for n in names)
            self.optim_hist_encoder = FlatAdam(self.ac_flat, self._lr_estimator0, self.max_grad_norm, lo=lo, hi=hi)
            self._dagger_loss = torch.zeros(1, device=self.device)
        return self.optim_hist_encoder

    # ---- checkpoints: the six torch.optim.Adam dicts of on_policy_runner.py:306-339 ------------------------------------
    @staticmethod
    def _disc_groups(disc, flat, optim_d, optim_q_eps, optim_q_c):
        wd = lambda name: {"weight_decay": 1e-3, "momentum": 0.9, "name": name}          # noqa: E731  gail.py:109-126
        trunk, head = ckpt.layout(flat, disc.trunk, "trunk."), ckpt.layout(flat, disc.linear, "linear.")
        eps, cls = ckpt.layout(flat, disc.encoder_eps, "encoder_eps."), ckpt.layout(flat, disc.classifier, "classifier.")
        return {"optim_d": [dict(adam=optim_d[0], params=trunk, extra=wd("trunk")), dict(adam=optim_d[0], params=head, extra=wd("head"))],
                "optim_q_eps": [dict(adam=optim_q_eps[0], params=trunk, extra=wd("trunk")),
                                dict(adam=optim_q_eps[1], params=eps, extra=wd("encoder_eps"))],
                "optim_q_c": [dict(adam=optim_q_c[0], params=trunk, extra=wd("trunk")),
                              dict(adam=optim_q_c[1], params=cls, extra=wd("classifier"))]}

    def _optim_groups(self):
        ac, he = self.actor_critic, self.actor_critic.history_encoder
        g = {"optim_ac": [dict(adam=self.optim_ac, params=ckpt.layout(self.ac_flat, ac), extra={"name": "actor_critic"})],
             "optim_hist_encoder": [dict(adam=self._ensure_hist_optimizer(), params=ckpt.layout(self.ac_flat, he, "history_encoder."))],
             "optim_estimator": [dict(adam=self.optim_estimator, params=ckpt.layout(self.est_flat, self.estimator))]}
        if getattr(self, "disc_flat", None) is not None:
            g.update(self._disc_groups(self.disc, self.disc_flat, self.optim_d, self.optim_q_eps, self.optim_q_c))
        return g

    def optimizer_state_dicts(self):
        """{'optim_ac', 'optim_hist_encoder', 'optim_estimator', 'optim_d', 'optim_q_eps', 'optim_q_c'} in the
        `torch.optim.Adam.state_dict()` layout of the reference (parameter numbering of gail.py:96-128)."""
        out = {k: ckpt.to_torch_state_dict(g) for k, g in self._optim_groups().items()}
        if "optim_d" not in out and self.disc is not None:
            pend = getattr(self, "_pending_disc_optim", None)
            if pend is not None:                                   # loaded, not yet stepped: hand the same dicts back
                out.update(pend)
            else:                                                  # never stepped: zero moments of the right shapes
                d = self.disc
                shapes = lambda m: [tuple(p.shape) for p in m.parameters()]              # noqa: E731
                grp = lambda m, name, lr: dict(shapes=shapes(m), lr=lr,                      # noqa: E731
                                               extra={"weight_decay": 1e-3, "momentum": 0.9, "name": name})
                dev = next(d.parameters()).device
                out["optim_d"] = ckpt.zero_torch_state_dict([grp(d.trunk, "trunk", self.lr_disc), grp(d.linear, "head", self.lr_disc)], dev)
                out["optim_q_eps"] = ckpt.zero_torch_state_dict([grp(d.trunk, "trunk", self.lr_q),
                                                                 grp(d.encoder_eps, "encoder_eps", self.lr_q)], dev)
                out["optim_q_c"] = ckpt.zero_torch_state_dict([grp(d.trunk, "trunk", self.lr_q),
                                                               grp(d.classifier, "classifier", self.lr_q)], dev)
        return out

    def load_optimizer_state_dicts(self, d):
        """Accepts the reference's torch.optim dicts (e.g. the shipped tsc/weights/bbc/model.pt) and this package's older flat
        format ({'exp_avg', 'exp_avg_sq', 'lr', 'step'})."""
        groups = self._optim_groups()
        for k in ("optim_ac", "optim_hist_encoder", "optim_estimator", "optim_d", "optim_q_eps", "optim_q_c"):
            sd = d.get(k)
            if sd is None:
                continue
            if ckpt.is_torch_state_dict(sd):
                if k in groups:
                    ckpt.from_torch_state_dict(sd, groups[k])
                else:                                              # discriminator optimisers are built on the first update
                    if getattr(self, "_pending_disc_optim", None) is None:
                        self._pending_disc_optim = {}
                    self._pending_disc_optim[k] = sd
            elif isinstance(sd, dict) and "exp_avg" in sd and k in groups and len(groups[k]) == 1:
                groups[k][0]["adam"].load_state_dict(sd)

    # ---- discriminator update (gail.py:415-541) ---------------------------------------------------------
    def _init_disc_update(self):
        """Flat discriminator parameters and the reference's three Adam optimisers (gail.py:107-128): optim_d = trunk + head,
        optim_q_eps = trunk + encoder_eps, optim_q_c = trunk + classifier, weight_decay 1e-3 each -- the trunk is stepped
        by all three, in that order, with separate moments."""
        d = self.disc
        if self.disc_loss_function == "WassersteinLoss":
            # the reference steps the critic with RMSprop in that mode (gail.py:124-125); K8 is an Adam kernel -- refuse rather
            # than train with a different optimiser (the Wasserstein REWARD mapping of predict_disc_reward is supported)
            raise NotImplementedError("discriminator update with WassersteinLoss needs RMSprop for optim_d (gail.py:124-125)")
        self.disc_flat = d.flatten_parameters()
        sl = self.disc_flat.slices
        lo = lambda n: sl[n][0]                                                # noqa: E731
        end = lambda n: sl[n][0] + sl[n][1]                                    # noqa: E731
        t0, t1 = lo("trunk.0.weight"), end("trunk.2.bias")
        l0, l1 = lo("linear.weight"), end("linear.bias")
        c0, c1 = lo("classifier.weight"), end("classifier.bias")
        e0, e1 = lo("encoder_eps.weight"), end("encoder_eps.bias")
        assert t1 == l0, "trunk and head must be adjacent in the flat layout"
        mk = lambda lr, a, b: FlatAdam(self.disc_flat, lr, 0.0, weight_decay=1e-3, lo=a, hi=b)      # noqa: E731  no clipping
        self.optim_d = [mk(self.lr_disc, t0, l1)]
        self.optim_q_eps = [mk(self.lr_q, t0, t1), mk(self.lr_q, e0, e1)]
        self.optim_q_c = [mk(self.lr_q, t0, t1), mk(self.lr_q, c0, c1)]
        self._disc_stats = torch.zeros(11, device=self.device)
        self.info_max_coef_on = 0.0
        self._info_max_coef_on = torch.zeros((), device=self.device)
        if not torch.is_tensor(getattr(self.env, "prior_parameters", None)):
            self.env.prior_parameters = torch.full((self.dim_c,), 1.0 / self.dim_c, device=self.device)
        pend = getattr(self, "_pending_disc_optim", None)
        if pend:                                                   # optimiser state loaded before the first update
            groups = self._disc_groups(d, self.disc_flat, self.optim_d, self.optim_q_eps, self.optim_q_c)
            for k, sd in pend.items():
                ckpt.from_torch_state_dict(sd, groups[k])
            self._pending_disc_optim = None

    def _disc_prepare(self, x):
        """gail.py:423-452: task-obs weighting, per-step multipliers, normalisation (no grad)."""
        L, W = self.disc_obs_len, self.num_disc_obs
        x = x.view(len(x), L, -1).clone()
        if self.env.task_obs_weight_decay:
            # the weight decays every iteration (on_policy_runner.py:224-225): it is read from a DEVICE scalar so that a
            # captured discriminator step sees the current value (a Python float would be frozen into the graph)
            w = self._task_obs_weight_dev()
            x[:, :, 3:9] *= w
            x[:, :, 33:] *= w
        x = x[:, -L:, :].reshape(len(x), -1)
        if self.obs_disc_weight_step != 0.0:
            x = x * (torch.arange(L, dtype=torch.float32, device=x.device) * self.obs_disc_weight_step + 1).repeat_interleave(W)
        if self.disc_normalizer is not None:
            with torch.no_grad():
                x = self.disc_normalizer.normalize_torch(x, x.device)
        return x

    def _task_obs_weight_dev(self, refresh: bool = None) -> torch.Tensor:
        """env.task_obs_weight as a 0-d device tensor at a fixed address.  Refreshed from the env on every call outside a
        CUDA-graph capture (`update_disc` refreshes it before replaying), never during one."""
        t = getattr(self, "_tow", None)
        if t is None:
            t = self._tow = torch.ones((), device=self.device)
        capturing = torch.device(self.device).type == "cuda" and torch.cuda.is_current_stream_capturing()
        if refresh or (refresh is None and not capturing):
            t.fill_(float(self.env.task_obs_weight))
        return t

    def update_ss_info_gail(self, sample_disc_policy, sample_disc_expert_lb, sample_disc_expert_ulb):
        """One discriminator minibatch step with the reference's signature and 11-tuple (device tensors, no host sync).
        MSELoss / BCEWithLogitsLoss / WassersteinLoss discriminator losses as in :471-481."""
        if getattr(self, "disc_flat", None) is None:
            self._init_disc_update()
        d, env = self.disc, self.env
        policy_state, policy_latent_eps, policy_latent_c = sample_disc_policy
        expert_state_lb, label_exp_lb = sample_disc_expert_lb
        g = None
        if self.disc_batched:
            # ONE input preparation and ONE trunk pass over [policy | labelled | unlabelled] rows instead of three plus the
            # gradient penalty's own forward (:492-502 re-runs the same weights on the same unlabelled batch, so its `dd` is
            # `logits_exp`): same per-row values, a third of the launches.  The penalty's input gradient is taken from the
            # shared graph, restricted to the unlabelled rows.
            sizes = [len(policy_state), len(expert_state_lb), len(sample_disc_expert_ulb)]
            x_all = self._disc_prepare(torch.cat([policy_state, expert_state_lb, sample_disc_expert_ulb], dim=0))
            policy_state, expert_state_lb, expert_state_ulb = x_all.split(sizes)
            x_req = x_all.detach().requires_grad_(True)
            d_all, eps_all, c_all = d.forward_torch(x_req)
            (logits_pi, _, logits_exp), (eps, _, _) = d_all.split(sizes), eps_all.split(sizes)
            pred_c, pred_c_lb, pred_c_ulb = c_all.split(sizes)
            (g_all,) = torch.autograd.grad(logits_exp, x_req, grad_outputs=torch.ones_like(logits_exp), create_graph=True,
                                           retain_graph=True)
            g = g_all[sizes[0] + sizes[1]:]
        else:
            policy_state = self._disc_prepare(policy_state)
            expert_state_lb = self._disc_prepare(expert_state_lb)
            expert_state_ulb = self._disc_prepare(sample_disc_expert_ulb)
            _, _, pred_c_lb = d.forward_torch(expert_state_lb)
            logits_pi, eps, pred_c = d.forward_torch(policy_state)
            logits_exp, _, pred_c_ulb = d.forward_torch(expert_state_ulb)
        ss_loss = F.cross_entropy(pred_c_lb, label_exp_lb)                      # on the soft-maxed output, as the reference
        lab_pi = torch.argmax(policy_latent_c, dim=-1)
        with torch.no_grad():                                                   # prior estimate :462-464
            prior_batch = pred_c_ulb.mean(dim=0)
            if self.world_size > 1:                                             # the union batch's mean (equal shards)
                prior_batch = qdist.allreduce_mean_scalar_(prior_batch.clone())
            env.prior_parameters.mul_(1 - self.prior_soft_coef).add_(prior_batch * self.prior_soft_coef)
        info_max_loss = torch.mean(-torch.sum(pred_c_ulb * torch.log(pred_c_ulb + 1e-20), dim=-1))
        if self.disc_loss_function == "BCEWithLogitsLoss":
            disc_exp_loss = F.binary_cross_entropy_with_logits(logits_exp, torch.ones_like(logits_exp))
            disc_pi_loss = F.binary_cross_entropy_with_logits(logits_pi, torch.zeros_like(logits_pi))
        elif self.disc_loss_function == "MSELoss":
            disc_exp_loss = F.mse_loss(logits_exp, torch.ones_like(logits_exp))
            disc_pi_loss = F.mse_loss(logits_pi, -1 * torch.ones_like(logits_pi))
        elif self.disc_loss_function == "WassersteinLoss":
            disc_exp_loss, disc_pi_loss = -logits_exp.mean(), logits_pi.mean()
        else:
            raise ValueError("Unexpected loss function specified")
        disc_loss = 0.5 * (disc_pi_loss + disc_exp_loss)
        us_loss = F.l1_loss(eps, policy_latent_eps)
        disc_logit_loss = torch.sum(torch.square(d.get_disc_logit_weights()))
        if g is None:
            sample_expert = expert_state_ulb.clone().requires_grad_(True)       # gradient penalty :492-502
            h = F.relu(F.linear(F.relu(F.linear(sample_expert, d.trunk[0].weight, d.trunk[0].bias)), d.trunk[2].weight, d.trunk[2].bias))
            dd = F.linear(h, d.linear.weight, d.linear.bias)
            (g,) = torch.autograd.grad(dd, sample_expert, grad_outputs=torch.ones_like(dd), create_graph=True, retain_graph=True)
        grad_pen_loss = torch.mean(torch.sum(torch.square(g), dim=-1))
        disc_weight_decay = torch.sum(torch.square(torch.cat(d.get_disc_weights(), dim=-1)))
        loss = (self.ss_coef * ss_loss + self._info_max_coef_on * info_max_loss + self.disc_coef * disc_loss +
                self.us_coef * us_loss + self.disc_grad_penalty * grad_pen_loss + self.disc_logit_reg * disc_logit_loss +
                self.disc_weight_decay * disc_weight_decay)
        self.disc_flat.zero_grad()
        loss.backward()
        self._disc_optim_step(qdist.allreduce_flat_(self.disc_flat.grad))       # env shards: ONE all-reduce per step, 1/W in K8
        ac = self.actor_critic
        if not ac.fixed_std and self.min_std is not None:                       # :523-524
            ac.std.data.clamp_(min=self.min_std)
        if self.disc_normalizer is not None:                                    # :527-529 (of the NORMALISED batches, as there)
            if self.disc_batched:                                               # one merge of the three batches' pooled moments
                self.disc_normalizer.update_torch(x_all, self.world_size)       # (Chan's merge is associative: same result
            else:                                                               # up to the fp32 rounding of the batch moments)
                for x in (policy_state, expert_state_lb, expert_state_ulb):
                    self.disc_normalizer.update_torch(x, self.world_size)
        with torch.no_grad():
            acc_lb = torch.mean((torch.argmax(pred_c_lb, dim=-1) == label_exp_lb).float())
            acc_pi, acc_exp = (logits_pi < 0).float().mean(), (logits_exp > 0).float().mean()
            acc_ulb = torch.mean((torch.argmax(pred_c, dim=-1) == lab_pi).float())
        return (ss_loss.detach(), info_max_loss.detach(), disc_loss.detach(), us_loss.detach(), grad_pen_loss.detach(),
                disc_logit_loss.detach(), disc_weight_decay.detach(), acc_lb, acc_pi, acc_exp, acc_ulb)

    def notify_disc_changed(self):
        """Discriminator parameters changed (update / checkpoint load): refresh the derived device copies (rollout_plan's stacked
        head matrix)."""
        for cb in getattr(self, "_disc_listeners", ()):
            cb()

    def _disc_optim_step(self, grad_scale: float = 1.0):
        opts = self.optim_d + self.optim_q_eps + self.optim_q_c                   # :519-521, in the reference's order
        flat = self.disc_flat
        chainable = (flat.data.is_cuda and len(opts) <= 8 and os.environ.get("QA_ADAM_CHAIN", "1") == "1" and
                     all(o.max_grad_norm <= 0 and o.lo % 4 == 0 and o.hi % 4 == 0 and o.betas == opts[0].betas and o.eps == opts[0].eps
                         for o in opts))
        if not chainable:
            for o in opts:
                o.step(grad_scale)
            return
        # K8c: the three optimisers' five parameter groups as ONE launch (an element of the shared trunk receives its three
        # updates in order, in registers) instead of 5 x (memset, norm / step kernel, update kernel)
        if getattr(self, "_adam_chain_ticket", None) is None:
            self._adam_chain_ticket = torch.zeros(1, device=flat.data.device, dtype=torch.int32)
        ops.adam_chain(flat.data, flat.grad, [(o.lo, o.hi, o.exp_avg, o.exp_avg_sq, o.lr, o.step_count, o.weight_decay) for o in opts],
                       self._adam_chain_ticket, opts[0].betas[0], opts[0].betas[1], opts[0].eps, grad_scale)

    def update_disc(self, expert, num_updates=None):
        """The discriminator half of SSInfoGAIL.update (gail.py:258-300): `4 * epochs * minibatches` minibatch steps of
        `T*N / that` samples each from the policy replay buffer and the labelled / unlabelled expert sets
        (`expert.preloaded_s_lb (n,98)`, `expert.preloaded_label (n,)`, `expert.preloaded_s_ulb (n,98)` as in
        motion_loader.py:513-526).  Index draws (np.random.choice with replacement there) are torch.randint on the device;
        the per-minibatch statistics are accumulated on the device and read back once.  Returns the reference's 11 means."""
        if getattr(self, "disc_flat", None) is None:
            self._init_disc_update()
        st, ds = self.storage, self.disc_storage
        n_mb = self.num_learning_epochs * self.num_mini_batches * 4 if num_updates is None else num_updates
        mb = st.num_envs * st.num_transitions_per_env // (self.num_learning_epochs * self.num_mini_batches * 4)
        if self.learning_steps >= self.begin_rim:                               # :251-253
            self.info_max_coef_on = min(self.info_max_coef * (self.learning_steps - self.begin_rim) / 10000, self.info_max_coef)
        self._info_max_coef_on.fill_(self.info_max_coef_on)
        self._task_obs_weight_dev(refresh=True)
        dev = self.device
        i_pi = torch.randint(ds.num_samples, (n_mb, mb), device=dev)
        i_lb = torch.randint(expert.preloaded_s_lb.shape[0], (n_mb, mb), device=dev)
        i_ulb = torch.randint(expert.preloaded_s_ulb.shape[0], (n_mb, mb), device=dev)
        self._disc_stats.zero_()
        if self.use_cuda_graph and (self.world_size == 1 or self.capture_collectives):
            key = (id(expert), mb, ds.states.data_ptr(), self._ensure_disc_plan(mb) is not None)
            if getattr(self, "_disc_graph_key", None) != key:
                self._capture_disc(expert, mb)
                self._disc_graph_key = key
                self._disc_stats.zero_()
            b = self._disc_mb
            for k in range(n_mb):
                b["i_pi"].copy_(i_pi[k])
                b["i_lb"].copy_(i_lb[k])
                b["i_ulb"].copy_(i_ulb[k])
                self._disc_graph.replay()
                ops._count(self._disc_graph_launches)
        else:
            for k in range(n_mb):
                self._disc_step(expert, i_pi[k], i_lb[k], i_ulb[k])
        if self.world_size > 1:                                                 # logged statistics: mean over the ranks
            qdist.allreduce_mean_scalar_(self._disc_stats)
        if hasattr(self.env, "refresh_prior"):                                  # the steps above moved env.prior_parameters
            self.env.refresh_prior()                                            # (:462-464): next rollout samples from it
        self.notify_disc_changed()
        return tuple((self._disc_stats / n_mb).tolist())

    def _ensure_disc_plan(self, mb):
        """The static-schedule discriminator step (disc_plan) when it applies: CUDA, tcgen05 layers, MSE loss, QA_DISC_PLAN != 0."""
        from . import linear
        from .disc_plan import DiscStepPlan
        if os.environ.get("QA_DISC_PLAN", "1") != "1" or linear.get_mode() != "tc" or DiscStepPlan.supported(self) is not None:
            self._disc_plan = None
            return None
        plan = getattr(self, "_disc_plan", None)
        if plan is None or plan.B != mb:
            plan = self._disc_plan = DiscStepPlan(self, mb)
        return plan

    def _disc_step(self, expert, i_pi, i_lb, i_ulb):
        plan = self._ensure_disc_plan(i_pi.shape[0])
        if plan is not None:
            plan.step(expert, i_pi, i_lb, i_ulb)
            return
        ds = self.disc_storage
        out = self.update_ss_info_gail((ds.states[i_pi], ds.latent_eps[i_pi], ds.latent_c[i_pi]),
                                       (expert.preloaded_s_lb[i_lb], expert.preloaded_label[i_lb]), expert.preloaded_s_ulb[i_ulb])
        self._disc_stats += torch.stack(out)

    def _capture_disc(self, expert, mb):
        """One discriminator minibatch step (3 x 2 gathers, ~150 torch kernels incl. the double backward, 5 K8 launches,
        the normaliser merge) as ONE CUDA graph over static index buffers; warm-up side effects are rolled back."""
        dev = self.device
        z = lambda: torch.zeros(mb, dtype=torch.int64, device=dev)             # noqa: E731
        self._disc_mb = dict(i_pi=z(), i_lb=z(), i_ulb=z())
        b = self._disc_mb
        if self.disc_normalizer is not None:
            self.disc_normalizer._device_state(dev)
        opts = self.optim_d + self.optim_q_eps + self.optim_q_c
        state = [self.disc_flat.data, self.env.prior_parameters, self._disc_stats]
        for o in opts:
            state += [o.exp_avg, o.exp_avg_sq, o.step_count]
        if not self.actor_critic.fixed_std:
            state.append(self.actor_critic.std.data)
        if self.disc_normalizer is not None:
            state += list(self.disc_normalizer._device_state(dev)[1:])
        snap = [t.clone() for t in state]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._disc_step(expert, b["i_pi"], b["i_lb"], b["i_ulb"])
        torch.cuda.current_stream().wait_stream(s)
        before = ops.launches
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._disc_step(expert, b["i_pi"], b["i_lb"], b["i_ulb"])
        self._disc_graph, self._disc_graph_launches = g, ops.launches - before
        for t, v in zip(state, snap):
            t.copy_(v)
        if self.disc_normalizer is not None:
            self.disc_normalizer.__dict__["_dev_dirty"] = True

    def update_actor_critic(self, sample):
        """The reference's per-minibatch entry point (gail.py:328-413) for code written against it: `sample` is the 11-tuple
        `RolloutStorage.mini_batch_generator` yields.  Copies it into the static minibatch buffers and runs the same fused step
        `update()` replays (estimator step, adaptive LR, clip + Adam); returns this step's six loss statistics as device scalars
        (surrogate, value, bound, entropy, priv_reg, estimator)."""
        obs, critic_obs, actions, target_values, advantages, returns, old_logp, old_mu, old_sigma = sample[:9]
        n = obs.shape[0]
        if getattr(self, "_mb", None) is None or self._mb["obs"].shape[0] != n:
            self._alloc_minibatch(n)
            if self._grad_arena is None:
                self._kl = torch.zeros((), device=self.device)
            self._graphs = None
        sch = self.priv_reg_coef_schedual
        stage = min(max((self.priv_reg_counter - sch[2]), 0) / sch[3], 1)
        plan = self._ensure_plan(n)
        if plan is not None:                                   # static schedule, eager (no graph for a one-off sample)
            plan.load(0, sample)
            p, e, l, h = self.num_prop, self.num_explicit, self.num_latent, self.num_hist * self.num_prop
            with torch.no_grad():
                plan.sets[0]["hist_latent"].copy_(self.actor_critic.infer_hist_latent(plan.sets[0]["obs"][:, p + e + l:p + e + l + h]))
            self._priv_reg_coef.fill_(stage * (sch[1] - sch[0]) + sch[0])
            before = self._stats.clone()
            plan.step(0)
            return tuple((self._stats - before)[:6])
        mb = self._mb
        for k, v in (("obs", obs), ("critic_obs", critic_obs), ("actions", actions), ("values", target_values),
                     ("advantages", advantages), ("returns", returns), ("old_actions_log_prob", old_logp), ("old_mu", old_mu),
                     ("old_sigma", old_sigma)):
            mb[k].copy_(v.reshape(mb[k].shape))
        p, e, l, h = self.num_prop, self.num_explicit, self.num_latent, self.num_hist * self.num_prop
        with torch.no_grad():
            mb["hist_latent"].copy_(self.actor_critic.infer_hist_latent(mb["obs"][:, p + e + l:p + e + l + h]))
        sch = self.priv_reg_coef_schedual
        stage = min(max((self.priv_reg_counter - sch[2]), 0) / sch[3], 1)
        self._priv_reg_coef.fill_(stage * (sch[1] - sch[0]) + sch[0])
        before = self._stats.clone()
        self._minibatch_step()
        return tuple((self._stats - before)[:6])

    # ---- PPO minibatch step ------------------------------------------------------------------------------
    def _alloc_minibatch(self, mb_size):
        st, dev = self.storage, self.device
        z = lambda w: torch.zeros(mb_size, w, device=dev)                       # noqa: E731
        # observation rows are gathered straight into a 16-byte row pitch: a legal TMA operand for the first layers
        zp = lambda w: torch.zeros(mb_size, (w + 3) // 4 * 4, device=dev)[:, :w]  # noqa: E731
        W, A = st.observations.shape[-1], st.actions.shape[-1]
        Wc = st.privileged_observations.shape[-1] if st.privileged_observations is not None else W
        self._mb = dict(obs=zp(W), critic_obs=zp(Wc), actions=z(A), values=z(1), returns=z(1),
                        old_actions_log_prob=z(1), advantages=z(1), old_mu=z(A), old_sigma=z(A),
                        hist_latent=z(self.num_latent))
        self._mb_keys = list(self._mb.keys())
        # history-encoder latents of the whole rollout: the encoder is frozen during update() (it is trained by
        # update_dagger), so its output per sample is computed once per update instead of once per epoch
        self._hist_latent_all = torch.zeros(st.num_transitions_per_env * st.num_envs, self.num_latent, device=dev)

    def _ensure_plan(self, mb_size):
        """The static-schedule step (ppo_plan) when it applies to this configuration, else None (autograd path)."""
        from . import linear
        from .ppo_plan import PpoStepPlan
        if not self.use_plan or linear.get_mode() != "tc" or PpoStepPlan.supported(self) is not None:
            if self._plan is not None:
                self._plan, self._graphs = None, None
            return None
        if self._plan is None or self._plan.M != mb_size:
            self._plan = PpoStepPlan(self, mb_size, self.num_mini_batches)
            self._graphs = None
        return self._plan

    def _gather(self, idx):
        v = self.storage.flat_views()
        v["hist_latent"] = self._hist_latent_all
        ops.gather_minibatch_windows(idx, [(v[k], 0, self._mb[k], 0, v[k].shape[1]) for k in self._mb_keys])

    @torch.no_grad()
    def _encode_history(self):
        """hist_latent (gail.py:352-353, under no_grad there too) for every stored sample, one K11 launch."""
        p, e, l, h = self.num_prop, self.num_explicit, self.num_latent, self.num_hist * self.num_prop
        flat_obs = self.storage.observations.flatten(0, 1)
        self._hist_latent_all.copy_(self.actor_critic.infer_hist_latent(flat_obs[:, p + e + l:p + e + l + h]))

    def _forward_backward(self):
        """Forward + both backward passes of one minibatch (gail.py:328-408) on the static minibatch buffers.
        Leaves gradients in the flat buffers, kl_mean in `self._kl`, and adds the loss statistics."""
        mb, ac, est = self._mb, self.actor_critic, self.estimator
        obs = mb["obs"]
        p, e, l, h = self.num_prop, self.num_explicit, self.num_latent, self.num_hist * self.num_prop
        if self.share_priv_latent:
            # the privileged-latent encoder feeds both the actor and the regulariser (gail.py:338, :352): one pass (and one
            # aligned copy of the 29 unaligned lanes) instead of two; autograd sums the two gradients into it
            priv_latent = ac.infer_priv_latent(obs[:, p + e:p + e + l])
            ac.update_distribution(obs, False, priv_latent=priv_latent)
        else:
            ac.update_distribution(obs, False)
        mu, sigma = ac.action_mean, ac.action_std
        logp = None if self.fused_loss else ac.get_actions_log_prob(mb["actions"])
        value = ac.evaluate(mb["critic_obs"])
        entropy = None if self.fused_loss else ac.entropy
        if not self.share_priv_latent:
            priv_latent = ac.infer_priv_latent(obs[:, p + e:p + e + l])
        hist_latent = mb["hist_latent"]              # gathered; computed once per update by _encode_history()
        if self.fused_loss:
            priv_reg_loss = _RowLossFused.apply(priv_latent, hist_latent, 1, self._aux_loss[0:1])
            est_loss = _RowLossFused.apply(est(obs[:, :p]), obs[:, p:p + e], 0, self._aux_loss[1:2])
        else:
            priv_reg_loss = (priv_latent - hist_latent).norm(p=2, dim=1).mean()
            est_loss = (est(obs[:, :p]) - obs[:, p:p + e]).pow(2).mean()                 # estimator (:359-365)
        self.est_flat.zero_grad()
        est_loss.backward()
        if self.fused_loss:
            cfg = dict(clip=self.clip_param, c_surr=self.surrogate_loss_coef, c_value=self.value_loss_coef,
                       c_bound=self.bounds_loss_coef, c_entropy=self.entropy_coef,
                       clipped_value=self.use_clipped_value_loss)
            main = _PPOLossFused.apply(mu, value, ac.std, mb, cfg, self._ppo_stats)
            ps = self._ppo_stats
            self._kl.copy_(ps[3])
            loss = main + self._priv_reg_coef * priv_reg_loss
        else:
            # KL for the adaptive schedule (:367-373)
            with torch.no_grad():
                osg, omu = mb["old_sigma"], mb["old_mu"]
                kl = torch.sum(torch.log(sigma / osg + 1.e-5) + (torch.square(osg) + torch.square(omu - mu)) /
                               (2.0 * torch.square(sigma)) - 0.5, dim=-1)
                self._kl.copy_(kl.mean())
            adv = mb["advantages"].squeeze(1)
            ratio = torch.exp(logp - mb["old_actions_log_prob"].squeeze(1))
            surrogate = -adv * ratio
            surrogate_clipped = -adv * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
            surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
            if self.use_clipped_value_loss:
                tv = mb["values"]
                value_clipped = tv + (value - tv).clamp(-self.clip_param, self.clip_param)
                value_loss = torch.max((value - mb["returns"]).pow(2), (value_clipped - mb["returns"]).pow(2)).mean()
            else:
                value_loss = (mb["returns"] - value).pow(2).mean()
            b_loss = (torch.clamp(mu + 1.0, max=0.) ** 2 + torch.clamp(mu - 1.0, min=0.) ** 2).sum(dim=-1).mean()
            ent = entropy.mean()
            loss = (self.surrogate_loss_coef * surrogate_loss + self.value_loss_coef * value_loss +
                    self.bounds_loss_coef * b_loss - self.entropy_coef * ent + self._priv_reg_coef * priv_reg_loss)
        self.ac_flat.zero_grad()
        loss.backward()
        if not self.fused_loss:
            with torch.no_grad():
                self._aux_loss.copy_(torch.stack([priv_reg_loss.detach(), est_loss.detach()]))
                self._ppo_stats.copy_(torch.stack([surrogate_loss.detach(), value_loss.detach(), b_loss.detach(), self._kl]))

    def _apply(self):
        """All-reduce (multi-GPU), adaptive LR on the device (:374-379), fused clip + Adam (:361-365, :409-412)."""
        scale = 1.0
        if self.world_size > 1:
            scale = self._allreduce_grads()
        pre = self.world_size > 1 and getattr(self, "_peer", None) is not None
        if not getattr(getattr(self, "_plan", None), "est_stepped_in_chain", False):
            self.optim_estimator.step(scale, presummed=pre)
        adaptive = self.desired_kl is not None and self.schedule == 'adaptive'
        if torch.device(self.device).type == "cuda":
            # K13: adaptive LR + the seven running statistics, one launch
            ops.ppo_scalars(self._ppo_stats, self.actor_critic.std.detach(), self._aux_loss[0:1], self._aux_loss[1:2],
                            self._kl.view(1), self.desired_kl if adaptive else 0.0, self.optim_ac.lr, self._stats)
        else:
            with torch.no_grad():
                ent = (1.4189385332046727 + torch.log(self.actor_critic.std.detach())).sum()
                ps, ax = self._ppo_stats, self._aux_loss
                self._stats += torch.stack([ps[0], ps[1], ps[2], ent, ax[0], ax[1], self._kl])
            if adaptive:
                lr, kl = self.optim_ac.lr, self._kl
                hi = kl > self.desired_kl * 2.0
                lo = (kl < self.desired_kl / 2.0) & (kl > 0.0)
                lr.copy_(torch.where(hi, torch.clamp(lr / 1.5, min=1e-5), torch.where(lo, torch.clamp(lr * 1.5, max=1e-2), lr)))
        self.optim_ac.step(scale, presummed=pre)

    def _minibatch_step(self):
        self._forward_backward()
        if getattr(self, "_plan", None) is not None:
            self._plan.est_stepped_in_chain = False        # autograd variant: the estimator's optimiser steps in _apply
        self._apply()

    def _capture(self):
        """Warm up on a side stream, then capture the minibatch step (one graph when single-GPU; the
        forward/backward graph and the apply graph are split around the eager NCCL all-reduce otherwise)."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        snap = [t.clone() for t in (self.ac_flat.data, self.est_flat.data, self.optim_ac.exp_avg, self.optim_ac.exp_avg_sq,
                                    self.optim_estimator.exp_avg, self.optim_estimator.exp_avg_sq, self.optim_ac.lr,
                                    self.optim_ac.step_count, self.optim_estimator.step_count, self._stats)]
        plan = self._plan
        sets = range(plan.num_sets) if plan is not None else (None,)
        fb = (lambda k: plan.forward_backward(k, step_estimator=True)) if plan is not None else (lambda k: self._forward_backward())
        with torch.cuda.stream(s):
            for _ in range(3):
                fb(sets[0])
                self._apply()
        torch.cuda.current_stream().wait_stream(s)
        self._graph_has_apply = self.world_size == 1 or self.capture_collectives
        graphs = []
        for k in sets:                                  # one graph per minibatch buffer set (the plan) or the single autograd one
            before = ops.launches
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                # with > 1 rank the NCCL all-reduce of the flat gradients is captured as a graph node, unless
                # QA_CAPTURE_COLLECTIVES=0: then the graph ends after the backward pass and _apply() runs eagerly
                fb(k)
                if self._graph_has_apply:
                    self._apply()
            graphs.append(g)
            self._graph_launches = ops.launches - before     # libqa_b200 kernels inside one replay
        self._graphs = tuple(graphs)
        # undo the warm-up / capture side effects on the trainable state
        for t, v in zip((self.ac_flat.data, self.est_flat.data, self.optim_ac.exp_avg, self.optim_ac.exp_avg_sq,
                         self.optim_estimator.exp_avg, self.optim_estimator.exp_avg_sq, self.optim_ac.lr,
                         self.optim_ac.step_count, self.optim_estimator.step_count, self._stats), snap):
            t.copy_(v)

    def update(self, indices: Optional[torch.Tensor] = None, expert=None):
        """SSInfoGAIL.update (:231-326): num_learning_epochs x num_mini_batches PPO minibatch steps over ONE permutation,
        then -- when `expert` (an object with preloaded_s_lb / preloaded_label / preloaded_s_ulb, e.g. the reference's
        MotionLoader) is given -- the 4x as many discriminator minibatch steps (:284-300).  Returns the reference's
        first six means (surrogate, value, bound, entropy, priv_reg, estimator), followed by the eleven discriminator
        means when the discriminator was updated."""
        st = self.storage
        batch = st.num_envs * st.num_transitions_per_env
        mb_size = batch // self.num_mini_batches
        if getattr(self, "_mb", None) is None or self._mb["obs"].shape[0] != mb_size:
            self._alloc_minibatch(mb_size)
            if self._grad_arena is None:
                self._kl = torch.zeros((), device=self.device)
            self._graphs = None
        self.learning_steps += 1
        sch = self.priv_reg_coef_schedual
        stage = min(max((self.priv_reg_counter - sch[2]), 0) / sch[3], 1)
        self._priv_reg_coef.fill_(stage * (sch[1] - sch[0]) + sch[0])
        if indices is None:
            indices = torch.randperm(self.num_mini_batches * mb_size, device=self.device)
        self._encode_history()
        plan = self._ensure_plan(mb_size)
        if plan is not None:
            # every minibatch is gathered ONCE per update: the generator re-uses the same slices of one permutation in every
            # epoch (rollout_storage.py:125, 140-145) and the storage does not change during the update
            for i in range(self.num_mini_batches):
                plan.gather(i, indices[i * mb_size:(i + 1) * mb_size], self._hist_latent_all)
            if self.use_cuda_graph and self._graphs is None:
                self._capture()
            self._stats.zero_()
            for _ in range(self.num_learning_epochs):
                for i in range(self.num_mini_batches):
                    if self.use_cuda_graph:
                        self._graphs[i].replay()
                        ops._count(self._graph_launches)
                        if not self._graph_has_apply:
                            self._apply()
                    else:
                        plan.step(i)
        else:
            if self.use_cuda_graph and self._graphs is None:
                self._gather(indices[:mb_size])
                self._capture()
            self._stats.zero_()
            for _ in range(self.num_learning_epochs):
                for i in range(self.num_mini_batches):
                    self._gather(indices[i * mb_size:(i + 1) * mb_size])
                    if self.use_cuda_graph:
                        self._graphs[0].replay()
                        ops._count(self._graph_launches)
                        if not self._graph_has_apply:
                            self._apply()
                    else:
                        self._minibatch_step()
        n = self.num_learning_epochs * self.num_mini_batches
        vals = (self._stats / n).tolist()                      # the one host sync of the update
        self.last_stats = dict(zip(STAT_NAMES, vals))
        disc_stats = ()
        if expert is not None:
            disc_stats = self.update_disc(expert)
        st.clear()
        self.priv_reg_counter += 1
        return tuple(vals[:6]) + tuple(disc_stats)

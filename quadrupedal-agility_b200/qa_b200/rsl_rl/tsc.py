"""TSC (task-level controller) trainer with the reference's API:

  `Actor`, `ActorCriticTSC`   tsc/rsl_rl/modules/actor_critic.py:60-284 -- same constructor arguments, method names and
                              `state_dict` keys (`actor.actor_trunk.0.weight`, `actor.scan_encoder.4.bias`, `critic.6.weight`,
                              `std`, ...)
  `RolloutStorageTSC`         tsc/rsl_rl/storage/rollout_storage.py:25-166 (two log-prob columns, mu / sigma of the 18
                              continuous actions)
  `PPO`                       tsc/rsl_rl/algorithms/ppo.py:8-282 (`act`, `process_env_step`, `compute_returns`, `update`)

underneath on the same machinery as the BBC trainer: flat parameter buffers, tcgen05 GEMM chains (K7), one gather
launch per minibatch (K6), the TSC loss graph as one forward+backward kernel (K15), fused row losses (K12), device-side
adaptive LR (K13), fused clip+Adam (K8), GAE warp-scan (K5), the whole minibatch step replayed as a CUDA graph.
Random draws can be injected (`normal_draw`, `mode_u`) for parity with the oracle.
"""
import os
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from .. import dist as qdist
from .algorithm import FlatAdam, _RowLossFused
from .linear import linear_act
from .modules import ActorCritic, FlatParams, StateHistoryEncoder, _mlp, get_activation, run_mlp

EPS = torch.finfo(torch.float32).eps


class Actor(nn.Module):
    """obs = [prop | scan | priv explicit | priv latent | history]; trunk input = [prop | scan latent | explicit | latent]
    (actor_critic.py:60-166)."""

    def __init__(self, num_prop, num_auxiliary, num_scan, num_actions_d, num_actions_c, scan_encoder_dims,
                 actor_hidden_dims, priv_encoder_dims, num_priv_latent, num_priv_explicit, num_hist, activation,
                 tanh_encoder_output=False):
        super().__init__()
        self.num_prop, self.num_auxiliary, self.num_scan, self.num_hist = num_prop, num_auxiliary, num_scan, num_hist
        self.num_actions_d, self.num_actions_c = num_actions_d, num_actions_c
        self.num_priv_latent, self.num_priv_explicit = num_priv_latent, num_priv_explicit
        self.if_scan_encode = scan_encoder_dims is not None and num_scan > 0
        act = activation if isinstance(activation, str) else "elu"
        self.activation_name = act
        if len(priv_encoder_dims) > 0:
            self.priv_encoder = _mlp([num_priv_latent] + list(priv_encoder_dims) + [num_priv_latent], act, last_act=True)
        else:
            self.priv_encoder = nn.Identity()
        self.history_encoder = StateHistoryEncoder(get_activation(act), num_prop - num_auxiliary, num_hist, num_priv_latent)
        if self.if_scan_encode:
            dims = [num_scan] + list(scan_encoder_dims)
            layers = []
            for i in range(len(dims) - 1):                                     # last layer ends in Tanh (:103-114)
                layers += [nn.Linear(dims[i], dims[i + 1]), nn.Tanh() if i == len(dims) - 2 else get_activation(act)]
            self.scan_encoder = nn.Sequential(*layers)
            self.scan_encoder_output_dim = scan_encoder_dims[-1]
        else:
            self.scan_encoder = nn.Identity()
            self.scan_encoder_output_dim = num_scan
        n_in = num_prop + self.scan_encoder_output_dim + num_priv_explicit + num_priv_latent
        self.actor_trunk = _mlp([n_in] + list(actor_hidden_dims), act, last_act=True)
        self.actor_d = nn.Linear(actor_hidden_dims[-1], num_actions_d)
        self.actor_c = nn.Linear(actor_hidden_dims[-1], num_actions_d * num_actions_c)

    def _scan_latent(self, obs):
        x = obs[:, self.num_prop:self.num_prop + self.num_scan]
        if not self.if_scan_encode:
            return x
        mods = list(self.scan_encoder)
        x = run_mlp(nn.Sequential(*mods[:-2]), x)                               # Linear+ELU stack on the GEMM chain
        return torch.tanh(linear_act(x, mods[-2].weight, mods[-2].bias, None))

    def forward(self, obs, hist_encoding: bool, eval=False, scandots_latent=None, priv_latent=None):
        scan = self._scan_latent(obs) if scandots_latent is None else scandots_latent
        o = self.num_prop + self.num_scan
        if hist_encoding:
            latent = self.infer_hist_latent(obs)
        else:                                             # a caller that also needs the latent may hand it in (one encoder pass)
            latent = self.infer_priv_latent(obs) if priv_latent is None else priv_latent
        n_in = self.num_prop + self.scan_encoder_output_dim + self.num_priv_explicit + self.num_priv_latent
        parts = [obs[:, :self.num_prop], scan, obs[:, o:o + self.num_priv_explicit], latent]
        pad = (n_in + 3) // 4 * 4 - n_in                                        # 16-byte row pitch: a legal TMA operand
        if pad:
            parts.append(torch.zeros(obs.shape[0], pad, device=obs.device, dtype=obs.dtype))
        x = torch.cat(parts, dim=1)[:, :n_in]
        return run_mlp(self.actor_trunk, x)

    def infer_priv_latent(self, obs):
        o = self.num_prop + self.num_scan + self.num_priv_explicit
        priv = obs[:, o:o + self.num_priv_latent]
        return run_mlp(self.priv_encoder, priv) if not isinstance(self.priv_encoder, nn.Identity) else priv

    def infer_hist_latent(self, obs):
        w = self.num_prop - self.num_auxiliary
        hist = obs[:, obs.shape[1] - self.num_hist * w:]
        enc = self.history_encoder
        if (obs.is_cuda and not torch.is_grad_enabled() and hist.stride(1) == 1 and enc.tsteps == 10 and w == 57
                and self.activation_name == "elu"):
            out = torch.empty(obs.shape[0], self.num_priv_latent, device=obs.device, dtype=torch.float32)
            ops.hist_encoder_fwd(hist, enc, out)                                # K11
            return out
        return enc(hist.reshape(-1, self.num_hist, w))

    def infer_scandots_latent(self, obs):
        return self._scan_latent(obs)


class ActorCriticTSC(nn.Module):
    is_recurrent = False

    def __init__(self, num_prop, num_auxiliary, num_scan, num_critic_obs, num_priv_latent, num_priv_explicit, num_hist,
                 num_actions_d, num_actions_c, scan_encoder_dims=[256, 256, 256], actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], activation='elu', init_noise_std=1.0, fixed_std=False,
                 device=torch.device('cuda'), **kwargs):
        super().__init__()
        self.kwargs = kwargs
        priv_encoder_dims = kwargs.get('priv_encoder_dims', [])
        self.num_actions_d, self.num_actions_c = num_actions_d, num_actions_c
        self.actor = Actor(num_prop, num_auxiliary, num_scan, num_actions_d, num_actions_c, scan_encoder_dims,
                           actor_hidden_dims, priv_encoder_dims, num_priv_latent, num_priv_explicit, num_hist, activation,
                           tanh_encoder_output=kwargs.get('tanh_encoder_output', False))
        self.critic = _mlp([num_critic_obs] + list(critic_hidden_dims) + [1], activation, last_act=False)
        std = init_noise_std * torch.ones(num_actions_d * num_actions_c)
        self.fixed_std = fixed_std
        self.std = std.clone().to(device) if fixed_std else nn.Parameter(std)
        self._logits = self._prob = self._logit = self._mean = self._std = None
        self.flat = None

    def flatten_parameters(self) -> FlatParams:
        self.flat = FlatParams(self)
        return self.flat

    def reset(self, dones=None):
        pass

    init_weights = staticmethod(ActorCritic.init_weights)

    def forward(self):
        raise NotImplementedError

    # ---- distributions (Categorical over modes, Normal over the continuous actions) ---------------------------
    @property
    def action_mean(self):
        return self._mean

    @property
    def action_std(self):
        return self._std

    @property
    def entropy_c(self):
        return (1.4189385332046727 + torch.log(self._std)).mean(dim=-1)

    @property
    def entropy_d(self):
        return -(self._logit * self._prob).sum(-1)

    def update_distribution(self, observations, hist_encoding=False, priv_latent=None):
        emb = self.actor(observations, hist_encoding, priv_latent=priv_latent)
        self._logits = linear_act(emb, self.actor.actor_d.weight, self.actor.actor_d.bias, None)
        prob = torch.softmax(self._logits, dim=-1)
        self._prob = prob / prob.sum(-1, keepdim=True)                          # Categorical(probs=prob)
        self._logit = torch.log(self._prob.clamp(min=EPS, max=1 - EPS))
        self._mean = linear_act(emb, self.actor.actor_c.weight, self.actor.actor_c.bias, None)
        self._std = self.std.to(self._mean.device).expand_as(self._mean)        # Normal(mean, mean*0. + std)

    def act(self, observations, hist_encoding=False, normal_draw=None, mode_u=None, **kwargs):
        self.update_distribution(observations, hist_encoding)
        if mode_u is None:
            mode_u = torch.rand(self._prob.shape[0], device=self._prob.device)
        cdf = torch.cumsum(self._prob, dim=-1)
        actions_d = torch.clamp((mode_u.unsqueeze(-1) >= cdf).sum(-1), max=self.num_actions_d - 1)
        if normal_draw is None:
            normal_draw = torch.randn_like(self._mean)
        actions_c = self._mean + self._std * normal_draw
        return torch.cat([actions_d.unsqueeze(-1).to(actions_c.dtype), actions_c], dim=-1).detach()

    def get_actions_log_prob_d(self, actions):
        return self._logit.gather(-1, actions.to(torch.int64).unsqueeze(-1)).squeeze(-1)

    def get_actions_log_prob_c(self, actions):
        var = self._std ** 2
        return (-((actions - self._mean) ** 2) / (2 * var) - torch.log(self._std) - 0.9189385332046727).sum(dim=-1)

    def act_inference(self, observations, hist_encoding=False, eval=False, scandots_latent=None, **kwargs):
        emb = self.actor(observations, hist_encoding, eval, scandots_latent)
        logits = linear_act(emb, self.actor.actor_d.weight, self.actor.actor_d.bias, None)
        actions_d = torch.argmax(torch.softmax(logits, dim=-1), dim=-1)
        actions_c = linear_act(emb, self.actor.actor_c.weight, self.actor.actor_c.bias, None)
        return torch.cat([actions_d.unsqueeze(-1).to(actions_c.dtype), actions_c], dim=-1)

    def evaluate(self, critic_observations, **kwargs):
        return run_mlp(self.critic, critic_observations)

    def reset_std(self, std, num_actions, device):
        self.std.data.copy_(std * torch.ones(num_actions, device=device))


class RolloutStorageTSC:
    """tsc/rsl_rl/storage/rollout_storage.py:25-166 on persistent device buffers; GAE through K5."""

    class Transition:
        def __init__(self):
            self.observations = self.critic_observations = self.actions = self.rewards = self.dones = None
            self.values = self.actions_log_prob_d = self.actions_log_prob_c = None
            self.action_mean = self.action_sigma = self.hidden_states = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device='cpu'):
        self.device = device
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        z = lambda *s, **k: torch.zeros(*s, device=device, **k)                # noqa: E731
        self.observations = z(T, N, *obs_shape)
        self.privileged_observations = z(T, N, *privileged_obs_shape) if privileged_obs_shape[0] is not None else None
        self.rewards, self.values, self.returns, self.advantages = z(T, N, 1), z(T, N, 1), z(T, N, 1), z(T, N, 1)
        self.actions = z(T, N, *actions_shape)
        self.dones = z(T, N, 1, dtype=torch.uint8)
        self.actions_log_prob_d, self.actions_log_prob_c = z(T, N, 1), z(T, N, 1)
        self.mu, self.sigma = z(T, N, actions_shape[0] - 1), z(T, N, actions_shape[0] - 1)
        self.num_transitions_per_env, self.num_envs = T, N
        self.step = 0
        self._gae_ws = torch.zeros(8, device=device, dtype=torch.float64) if torch.device(device).type == "cuda" else None

    def add_transitions(self, tr):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.observations[s].copy_(tr.observations)
        if self.privileged_observations is not None:
            self.privileged_observations[s].copy_(tr.critic_observations)
        self.actions[s].copy_(tr.actions)
        self.rewards[s].copy_(tr.rewards.view(-1, 1))
        self.dones[s].copy_(tr.dones.view(-1, 1))
        self.values[s].copy_(tr.values)
        self.actions_log_prob_d[s].copy_(tr.actions_log_prob_d.view(-1, 1))
        self.actions_log_prob_c[s].copy_(tr.actions_log_prob_c.view(-1, 1))
        self.mu[s].copy_(tr.action_mean)
        self.sigma[s].copy_(tr.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        ops.gae(self.rewards, self.values, self.dones, last_values.contiguous(), self.returns, self.advantages,
                self._gae_ws, gamma, lam)

    def get_statistics(self):
        """rollout_storage.py:118-125 (mean trajectory length, mean reward); like the reference it marks the last step done."""
        done = self.dones
        done[-1] = 1
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero(as_tuple=False)[:, 0]))
        return (done_indices[1:] - done_indices[:-1]).float().mean(), self.rewards.mean()

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, indices=None):
        """The reference's 12-tuple generator (:127-166): ONE permutation per call, the same slices in every epoch.  `PPO.update`
        gathers through K6 instead; this is for code written against the reference's storage."""
        mb = self.num_envs * self.num_transitions_per_env // num_mini_batches
        if indices is None:
            indices = torch.randperm(num_mini_batches * mb, device=self.device)
        obs = self.observations.flatten(0, 1)
        critic = self.privileged_observations.flatten(0, 1) if self.privileged_observations is not None else obs
        cols = [t.flatten(0, 1) for t in (self.actions, self.values, self.advantages, self.returns, self.actions_log_prob_d,
                                          self.actions_log_prob_c, self.mu, self.sigma)]
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                idx = indices[i * mb:(i + 1) * mb]
                yield (obs[idx], critic[idx], *(c[idx] for c in cols), (None, None), None)

    def flat_views(self):
        f = lambda t: t.flatten(0, 1)                                           # noqa: E731
        crit = self.privileged_observations if self.privileged_observations is not None else self.observations
        return dict(obs=f(self.observations), critic_obs=f(crit), actions=f(self.actions), values=f(self.values),
                    returns=f(self.returns), old_actions_log_prob_d=f(self.actions_log_prob_d),
                    old_actions_log_prob_c=f(self.actions_log_prob_c), advantages=f(self.advantages),
                    old_mu=f(self.mu), old_sigma=f(self.sigma))


class _PPOLossTSCFused(torch.autograd.Function):
    """K15: the loss graph of tsc ppo.py:176-262, forward and backward in one kernel launch."""

    @staticmethod
    def forward(ctx, logits, mu, value, std, mb, cfg, stats):
        M, dev = mu.shape[0], mu.device
        dlogits = torch.empty(M, 4, device=dev)[:, :logits.shape[1]]
        dmu = torch.empty(M, (mu.shape[1] + 3) // 4 * 4, device=dev)[:, :mu.shape[1]]
        dvalue, dstd = torch.empty(M, device=dev), torch.empty(std.shape[0], device=dev)
        ops.ppo_loss_tsc(logits, mu, std.detach().contiguous(), value, mb["actions"], mb["old_actions_log_prob_d"].view(-1),
                         mb["old_actions_log_prob_c"].view(-1), mb["advantages"].view(-1), mb["returns"].view(-1),
                         mb["values"].view(-1), mb["old_mu"], mb["old_sigma"], dlogits, dmu, dvalue, dstd, stats,
                         cfg["clip"], cfg["c_value"], cfg["c_entropy"], cfg["clipped_value"])
        ctx.save_for_backward(dlogits, dmu, dvalue, dstd)
        return stats[0] * 1.0            # roots the backward pass; the value of the combined loss is never read

    @staticmethod
    def backward(ctx, g):
        dlogits, dmu, dvalue, dstd = ctx.saved_tensors
        return g * dlogits, g * dmu, (g * dvalue).unsqueeze(1), g * dstd, None, None, None


class PPO:
    """tsc/rsl_rl/algorithms/ppo.py:8-383: teacher PPO (`act`, `process_env_step`, `compute_returns`, `update`, `update_dagger`)
    and the depth-student distillation updates (`update_depth_encoder` / `update_depth_actor` / `update_depth_both`)."""

    def __init__(self, actor_critic, actor_critic_bbc, estimator, estimator_paras, depth_encoder=None,
                 depth_encoder_paras=None, depth_actor=None, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2,
                 gamma=0.998, lam=0.95, value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, device='cpu', dagger_update_freq=20,
                 priv_reg_coef_schedual=[0, 0, 0], use_cuda_graph=True, fused_loss=True, **kwargs):
        self.device = device
        self.desired_kl, self.schedule = desired_kl, schedule
        self.actor_critic = actor_critic.to(device)
        self.actor_critic_bbc = actor_critic_bbc.to(device) if actor_critic_bbc is not None else None
        self.estimator = estimator.to(device)
        self.ac_flat = self.actor_critic.flatten_parameters()
        self.est_flat = self.estimator.flatten_parameters()
        self.optimizer = FlatAdam(self.ac_flat, learning_rate, max_grad_norm)
        self.estimator_optimizer = FlatAdam(self.est_flat, estimator_paras["learning_rate"], max_grad_norm)
        self.storage: Optional[RolloutStorageTSC] = None
        self.transition = RolloutStorageTSC.Transition()
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.gamma, self.lam, self.max_grad_norm = gamma, lam, max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.priv_reg_coef_schedual, self.counter = priv_reg_coef_schedual, 0
        self.priv_states_dim, self.num_prop = estimator_paras["priv_states_dim"], estimator_paras["num_prop"]
        self.num_auxiliary, self.num_scan = estimator_paras["num_auxiliary"], estimator_paras["num_scan"]
        self.train_with_estimated_states = estimator_paras["train_with_estimated_states"]
        self.if_depth = depth_encoder is not None
        self.num_actions_d = self.actor_critic.num_actions_d
        self._lr0, self.hist_encoder_optimizer = learning_rate, None
        self.share_priv_latent = os.environ.get("QA_SHARE_PRIV_LATENT", "0") == "1"   # opt-in, see SSInfoGAIL
        if self.if_depth:
            # student distillation (ppo.py:78-88): conv / GRU / batch-norm modules on cuDNN with torch.optim.Adam;
            # the three optimisers overlap on purpose, each with its own moments
            self.depth_encoder, self.depth_actor = depth_encoder.to(device), depth_actor.to(device)
            self.depth_encoder_paras = depth_encoder_paras
            lr = depth_encoder_paras["learning_rate"]
            self.depth_encoder_optimizer = torch.optim.Adam(self.depth_encoder.parameters(), lr=lr)
            self.depth_actor_optimizer = torch.optim.Adam([*self.depth_actor.parameters(), *self.depth_encoder.parameters()], lr=lr)
            self.byol_optimizer = torch.optim.Adam(self.depth_encoder.byol_learner.parameters(),
                                                   lr=depth_encoder_paras["learning_rate_byol"])
        self.CE_loss = nn.CrossEntropyLoss()
        self.world_size = (torch.distributed.get_world_size()
                           if torch.distributed.is_available() and torch.distributed.is_initialized() else 1)
        cuda = torch.device(device).type == "cuda"
        self.use_cuda_graph, self.fused_loss = use_cuda_graph and cuda, fused_loss and cuda
        self._graph = None
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        self._ppo_stats = torch.zeros(4, device=device)           # surrogate, value, entropy, kl
        self._aux_loss = torch.zeros(2, device=device)            # priv_reg, estimator
        self._priv_reg_coef = torch.zeros((), device=device)
        self._stats = torch.zeros(7, device=device)
        self._kl = torch.zeros((), device=device)
        self._mb = None
        self.last_stats = {}

    @property
    def learning_rate(self) -> float:
        return float(self.optimizer.lr.item())

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorageTSC(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape,
                                         action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def _explicit_slice(self):
        o = self.num_prop + self.num_auxiliary + self.num_scan
        return slice(o, o + self.priv_states_dim)

    @torch.no_grad()
    def act(self, obs, critic_obs, info=None, hist_encoding=False, normal_draw=None, mode_u=None):
        """ppo.py:101-125."""
        tr, ac = self.transition, self.actor_critic
        if self.train_with_estimated_states:
            obs_est = obs.clone()
            obs_est[:, self._explicit_slice()] = self.estimator(obs_est[:, :self.num_prop])
            tr.actions = ac.act(obs_est, hist_encoding, normal_draw=normal_draw, mode_u=mode_u)
        else:
            tr.actions = ac.act(obs, hist_encoding, normal_draw=normal_draw, mode_u=mode_u)
        tr.values = ac.evaluate(critic_obs)
        tr.actions_log_prob_d = ac.get_actions_log_prob_d(tr.actions[:, 0])
        tr.actions_log_prob_c = ac.get_actions_log_prob_c(tr.actions[:, 1:])
        tr.action_mean, tr.action_sigma = ac.action_mean, ac.action_std
        tr.observations, tr.critic_observations = obs, critic_obs
        return tr.actions

    @torch.no_grad()
    def act_bbc(self, obs):
        """ppo.py:127-137: the frozen low-level controller (a BBC ActorCritic) turns TSC commands into joint targets."""
        bbc = self.actor_critic_bbc
        if self.train_with_estimated_states:
            obs_est = obs.clone()
            o = self.num_prop + self.num_scan            # NB the reference indexes without num_auxiliary here (:133)
            obs_est[:, o:o + self.priv_states_dim] = self.estimator(obs_est[:, :self.num_prop])
            return bbc.act_inference(obs_est, hist_encoding=True)
        return bbc.act_inference(obs, hist_encoding=True)

    @torch.no_grad()
    def process_env_step(self, rewards, dones, infos):
        """ppo.py:139-155."""
        rewards_total = rewards.clone()
        tr = self.transition
        tr.rewards = rewards_total.clone()
        tr.dones = dones
        if 'time_outs' in infos:
            tr.rewards += self.gamma * torch.squeeze(tr.values * infos['time_outs'].unsqueeze(1).to(self.device), 1)
        self.storage.add_transitions(tr)
        tr.clear()
        self.actor_critic.reset(dones)
        return rewards_total

    @torch.no_grad()
    def compute_returns(self, last_critic_obs):
        self.storage.compute_returns(self.actor_critic.evaluate(last_critic_obs), self.gamma, self.lam)

    # ---- minibatch step ------------------------------------------------------------------------------------
    def _alloc_minibatch(self, mb_size):
        st, dev = self.storage, self.device
        z = lambda w: torch.zeros(mb_size, w, device=dev)                       # noqa: E731
        zp = lambda w: torch.zeros(mb_size, (w + 3) // 4 * 4, device=dev)[:, :w]  # noqa: E731
        W, A = st.observations.shape[-1], st.actions.shape[-1]
        Wc = st.privileged_observations.shape[-1] if st.privileged_observations is not None else W
        lat = self.actor_critic.actor.num_priv_latent
        self._mb = dict(obs=zp(W), critic_obs=zp(Wc), actions=z(A), values=z(1), returns=z(1), old_actions_log_prob_d=z(1),
                        old_actions_log_prob_c=z(1), advantages=z(1), old_mu=z(A - 1), old_sigma=z(A - 1),
                        hist_latent=z(lat))
        self._mb_keys = list(self._mb.keys())
        self._hist_latent_all = torch.zeros(st.num_transitions_per_env * st.num_envs, lat, device=dev)
        self._graph = None

    def _gather(self, idx):
        v = self.storage.flat_views()
        v["hist_latent"] = self._hist_latent_all
        ops.gather_minibatch(idx, [v[k] for k in self._mb_keys], [self._mb[k] for k in self._mb_keys])

    @torch.no_grad()
    def _encode_history(self):
        """hist_latent of every stored sample (ppo.py:183-184 runs it under inference_mode per minibatch; the history
        encoder is frozen during update(), so once per update is the same numbers)."""
        self._hist_latent_all.copy_(self.actor_critic.actor.infer_hist_latent(self.storage.observations.flatten(0, 1)))

    def _forward_backward(self):
        mb, ac, est = self._mb, self.actor_critic, self.estimator
        obs = mb["obs"]
        if self.share_priv_latent:                                          # ppo.py:176, :186 evaluate the encoder twice
            priv_latent = ac.actor.infer_priv_latent(obs)
            ac.update_distribution(obs, False, priv_latent=priv_latent)
        else:
            ac.update_distribution(obs, False)
        value = ac.evaluate(mb["critic_obs"])
        if not self.share_priv_latent:
            priv_latent = ac.actor.infer_priv_latent(obs)
        sl = self._explicit_slice()
        if self.fused_loss:
            priv_reg_loss = _RowLossFused.apply(priv_latent, mb["hist_latent"], 1, self._aux_loss[0:1])
            est_loss = _RowLossFused.apply(est(obs[:, :self.num_prop]), obs[:, sl], 0, self._aux_loss[1:2])
        else:
            priv_reg_loss = (priv_latent - mb["hist_latent"]).norm(p=2, dim=1).mean()
            est_loss = (est(obs[:, :self.num_prop]) - obs[:, sl]).pow(2).mean()
        self.est_flat.zero_grad()
        est_loss.backward()
        if self.fused_loss:
            cfg = dict(clip=self.clip_param, c_value=self.value_loss_coef, c_entropy=self.entropy_coef,
                       clipped_value=self.use_clipped_value_loss)
            main = _PPOLossTSCFused.apply(ac._logits, ac._mean, value, ac.std, mb, cfg, self._ppo_stats)
            self._kl.copy_(self._ppo_stats[3])
            loss = main + self._priv_reg_coef * priv_reg_loss
        else:
            mu, sigma = ac.action_mean, ac.action_std
            logp_d = ac.get_actions_log_prob_d(mb["actions"][:, 0])
            logp_c = ac.get_actions_log_prob_c(mb["actions"][:, 1:])
            entropy = ac.entropy_c + ac.entropy_d
            with torch.no_grad():
                osg, omu = mb["old_sigma"], mb["old_mu"]
                kl = torch.sum(torch.log(sigma / osg + 1.e-5) + (torch.square(osg) + torch.square(omu - mu)) /
                               (2.0 * torch.square(sigma)) - 0.5, dim=-1)
                self._kl.copy_(kl.mean())
            adv = mb["advantages"].squeeze(1)

            def surrogate(logp, old):
                ratio = torch.exp(logp - old.squeeze(1))
                return torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)).mean()

            surrogate_loss = surrogate(logp_d, mb["old_actions_log_prob_d"]) + surrogate(logp_c, mb["old_actions_log_prob_c"])
            if self.use_clipped_value_loss:
                tv = mb["values"]
                vc = tv + (value - tv).clamp(-self.clip_param, self.clip_param)
                value_loss = torch.max((value - mb["returns"]).pow(2), (vc - mb["returns"]).pow(2)).mean()
            else:
                value_loss = (mb["returns"] - value).pow(2).mean()
            loss = (surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy.mean() +
                    self._priv_reg_coef * priv_reg_loss)
            with torch.no_grad():
                self._aux_loss.copy_(torch.stack([priv_reg_loss.detach(), est_loss.detach()]))
                self._ppo_stats.copy_(torch.stack([surrogate_loss.detach(), value_loss.detach(), entropy.mean().detach(), self._kl]))
        self.ac_flat.zero_grad()
        loss.backward()

    def _apply(self):
        scale = 1.0
        if self.world_size > 1:
            scale = qdist.allreduce_flat_(self.ac_flat.grad)
            qdist.allreduce_flat_(self.est_flat.grad)
            qdist.allreduce_mean_scalar_(self._kl)
        self.estimator_optimizer.step(scale)
        adaptive = self.desired_kl is not None and self.schedule == 'adaptive'
        if torch.device(self.device).type == "cuda":
            ops.ppo_scalars(self._ppo_stats, self.actor_critic.std.detach(), self._aux_loss[0:1], self._aux_loss[1:2],
                            self._kl.view(1), self.desired_kl if adaptive else 0.0, self.optimizer.lr, self._stats)
        else:
            with torch.no_grad():
                ps, ax = self._ppo_stats, self._aux_loss
                self._stats += torch.stack([ps[0], ps[1], ps[2], ps[2] * 0, ax[0], ax[1], self._kl])
            if adaptive:
                lr, kl = self.optimizer.lr, self._kl
                hi = kl > self.desired_kl * 2.0
                lo = (kl < self.desired_kl / 2.0) & (kl > 0.0)
                lr.copy_(torch.where(hi, torch.clamp(lr / 1.5, min=1e-5), torch.where(lo, torch.clamp(lr * 1.5, max=1e-2), lr)))
        self.optimizer.step(scale)

    def _minibatch_step(self):
        self._forward_backward()
        self._apply()

    def _capture(self):
        state = (self.ac_flat.data, self.est_flat.data, self.optimizer.exp_avg, self.optimizer.exp_avg_sq,
                 self.estimator_optimizer.exp_avg, self.estimator_optimizer.exp_avg_sq, self.optimizer.lr,
                 self.optimizer.step_count, self.estimator_optimizer.step_count, self._stats)
        snap = [t.clone() for t in state]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._minibatch_step()
        torch.cuda.current_stream().wait_stream(s)
        before = ops.launches
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._minibatch_step()
        self._graph, self._graph_launches = g, ops.launches - before
        for t, v in zip(state, snap):
            t.copy_(v)

    def update(self, indices: Optional[torch.Tensor] = None):
        """ppo.py:159-262.  Returns the reference's 7-tuple (value, surrogate, estimator, 0, 0, priv_reg, priv_reg_coef)."""
        st = self.storage
        mb_size = st.num_envs * st.num_transitions_per_env // self.num_mini_batches
        if self._mb is None or self._mb["obs"].shape[0] != mb_size:
            self._alloc_minibatch(mb_size)
        sch = self.priv_reg_coef_schedual
        stage = min(max((self.counter - sch[2]), 0) / sch[3], 1)
        coef = stage * (sch[1] - sch[0]) + sch[0]
        self._priv_reg_coef.fill_(coef)
        if indices is None:
            indices = torch.randperm(self.num_mini_batches * mb_size, device=self.device)
        self._encode_history()
        if self.use_cuda_graph and self._graph is None:
            self._gather(indices[:mb_size])
            self._capture()
        self._stats.zero_()
        for _ in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._gather(indices[i * mb_size:(i + 1) * mb_size])
                if self.use_cuda_graph:
                    self._graph.replay()
                    ops._count(self._graph_launches)
                else:
                    self._minibatch_step()
        n = self.num_learning_epochs * self.num_mini_batches
        vals = (self._stats / n).tolist()                      # the one host sync of the update
        self.last_stats = dict(surrogate_loss=vals[0], value_loss=vals[1], entropy=vals[2], priv_reg_loss=vals[4],
                               estimator_loss=vals[5], kl_mean=vals[6])
        st.clear()
        self.update_counter()
        return vals[1], vals[0], vals[5], 0.0, 0.0, vals[4], coef

    def update_counter(self):
        self.counter += 1

    # ---- DAgger: history-encoder adaptation (ppo.py:284-314) -----------------------------------------------------------
    def update_dagger(self, indices: Optional[torch.Tensor] = None):
        """Fits `actor.history_encoder` to the frozen privileged-latent encoder over the stored rollout (the buffers still
        hold it after `update()`'s `storage.clear()`, exactly what the reference relies on): per minibatch
        mean ||priv_latent - hist_latent||_2 (K12 forward+backward), clip_grad_norm_ over the encoder, Adam (K8) on the
        encoder's slice of the flat buffer at the PPO learning rate the optimiser was built with (:64).  Returns the mean loss."""
        st, actor = self.storage, self.actor_critic.actor
        if self.hist_encoder_optimizer is None:
            names = [n for n in self.ac_flat.slices if n.startswith("actor.history_encoder.")]
            lo = min(self.ac_flat.slices[n][0] for n in names)
            hi = max(self.ac_flat.slices[n][0] + self.ac_flat.slices[n][1] for n in names)
            self.hist_encoder_optimizer = FlatAdam(self.ac_flat, self._lr0, self.max_grad_norm, lo=lo, hi=hi)
            self._dagger_loss = torch.zeros(1, device=self.device)
        opt = self.hist_encoder_optimizer
        flat_obs = st.observations.flatten(0, 1)
        mb_size = flat_obs.shape[0] // self.num_mini_batches
        if indices is None:
            indices = torch.randperm(self.num_mini_batches * mb_size, device=self.device)
        with torch.no_grad():
            priv_all = actor.infer_priv_latent(flat_obs)
        total = torch.zeros((), device=self.device)
        w = actor.num_prop - actor.num_auxiliary                              # the actor's proprioception incl. auxiliary lanes
        for _ in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                idx = indices[i * mb_size:(i + 1) * mb_size]
                hist = flat_obs[idx][:, flat_obs.shape[1] - actor.num_hist * w:]
                hist_latent = actor.history_encoder(hist.reshape(-1, actor.num_hist, w))      # with grad: torch modules
                priv = priv_all[idx]
                if self.fused_loss:
                    loss = _RowLossFused.apply(hist_latent, priv, 1, self._dagger_loss)
                else:
                    loss = (priv - hist_latent).norm(p=2, dim=1).mean()
                self.ac_flat.grad[opt.lo:opt.hi].zero_()
                loss.backward()
                opt.step(qdist.allreduce_flat_(self.ac_flat.grad[opt.lo:opt.hi]))       # env shards: rank-mean gradient (1/W in K8)
                total += loss.detach()
        st.clear()
        self.update_counter()
        return float(total.item()) / (self.num_learning_epochs * self.num_mini_batches)

    # ---- depth-student distillation (ppo.py:316-383) --------------------------------------------------------
    def update_depth_encoder(self, depth_latent_batch, scandots_latent_batch):
        if not self.if_depth:
            return None
        loss = (scandots_latent_batch.detach() - depth_latent_batch).norm(p=2, dim=1).mean()
        self.depth_encoder_optimizer.zero_grad()
        loss.backward()
        qdist.allreduce_mean_grads_(list(self.depth_encoder.parameters()))
        nn.utils.clip_grad_norm_(self.depth_encoder.parameters(), self.max_grad_norm)
        self.depth_encoder_optimizer.step()
        return loss.item()

    def depth_actor_losses(self, actions_student, actions_teacher, yaw_student, yaw_teacher, obst_student, obst_teacher):
        """The three distillation terms of :329-336 as tensors: (mode cross-entropy + continuous L2, yaw L2 with lane
        weights (2, 0.5), obstacle-type cross-entropy -- on the ALREADY soft-maxed student lanes, as the reference does)."""
        nd = self.num_actions_d
        d_loss = self.CE_loss(actions_student[:, :nd], actions_teacher[:, 0].detach().to(torch.int64))
        c_loss = (actions_teacher[:, 1:].detach() - actions_student[:, nd:]).norm(p=2, dim=1).mean()
        scale = torch.tensor([2.0, 0.5], device=yaw_teacher.device)
        yaw_loss = ((yaw_teacher.detach() - yaw_student) * scale).norm(p=2, dim=1).mean()
        obst_loss = self.CE_loss(obst_student, torch.argmax(obst_teacher, dim=-1))
        return d_loss + c_loss, yaw_loss, obst_loss

    def update_depth_actor(self, actions_student_batch, actions_teacher_batch, yaw_student_batch, yaw_teacher_batch,
                           obst_type_buffer_student, obst_type_buffer_teacher, depth_batch, byol_perm=None):
        """One distillation step through the whole rollout's graph (student actor + depth encoder incl. the GRU), then six
        BYOL minibatch steps over the rollout's shuffled depth images with the EMA target update after each (:338-358).
        Returns (depth_actor_loss, yaw_loss, obst_type_loss, mean_byol_loss)."""
        if not self.if_depth:
            return None
        actor_loss, yaw_loss, obst_loss = self.depth_actor_losses(
            actions_student_batch, actions_teacher_batch, yaw_student_batch, yaw_teacher_batch, obst_type_buffer_student,
            obst_type_buffer_teacher)
        loss = actor_loss + yaw_loss + obst_loss
        self.depth_actor_optimizer.zero_grad()
        loss.backward()
        # env-sharded ranks (SURVEY 8e, config "TSC-student ... 2xB200"): one gradient all-reduce per optimiser step
        qdist.allreduce_mean_grads_([*self.depth_actor.parameters(), *self.depth_encoder.parameters()])
        nn.utils.clip_grad_norm_(self.depth_actor.parameters(), self.max_grad_norm)
        self.depth_actor_optimizer.step()
        n = depth_batch.size(0)
        bs = n // 6
        perm = torch.randperm(n) if byol_perm is None else byol_perm
        depth_batch = depth_batch[perm.to(depth_batch.device)]
        byol_total = torch.zeros((), device=depth_batch.device)
        learner = self.depth_encoder.byol_learner
        for i in range(0, n, bs):
            byol_loss = learner(depth_batch[i:i + bs])
            self.byol_optimizer.zero_grad()
            byol_loss.backward()
            qdist.allreduce_mean_grads_(list(learner.parameters()))           # batch-norm statistics: SyncBatchNorm (byol.py:44-46)
            self.byol_optimizer.step()
            byol_total += byol_loss.detach()
            learner.update_moving_average()
        stats = torch.stack([actor_loss.detach(), yaw_loss.detach(), obst_loss.detach(), byol_total / (n // bs)])
        a, y, o, b = stats.tolist()                                     # one host sync for the four statistics
        return a, y, o, b

    def update_depth_both(self, depth_latent_batch, scandots_latent_batch, actions_student_batch, actions_teacher_batch):
        if not self.if_depth:
            return None
        enc_loss = (scandots_latent_batch.detach() - depth_latent_batch).norm(p=2, dim=1).mean()
        act_loss = (actions_teacher_batch.detach() - actions_student_batch).norm(p=2, dim=1).mean()
        self.depth_actor_optimizer.zero_grad()
        (enc_loss + act_loss).backward()
        qdist.allreduce_mean_grads_([*self.depth_actor.parameters(), *self.depth_encoder.parameters()])
        nn.utils.clip_grad_norm_([*self.depth_actor.parameters(), *self.depth_encoder.parameters()], self.max_grad_norm)
        self.depth_actor_optimizer.step()
        return enc_loss.item(), act_loss.item()

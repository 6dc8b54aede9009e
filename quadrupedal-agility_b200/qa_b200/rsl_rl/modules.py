"""ActorCritic / Estimator / Discriminator with the reference's constructor signatures, method names
and `state_dict` keys (bbc/rsl_rl/modules/actor_critic.py:62-225, modules/estimator.py:12-40,
algorithms/discriminator.py:12-118), so reference checkpoints load key-for-key.

What differs from the reference is underneath: every parameter of a network is a VIEW into one flat fp32
buffer (`FlatParams`), with a matching flat gradient buffer -- the unit the NCCL all-reduce and the fused
clip+Adam kernel (K8) operate on; and the dense layers run through `qa_b200.rsl_rl.linear`, the single
dense-contraction site of the hot path.
"""
import types
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .linear import linear_act, mlp_chain


def get_activation(act_name):
    table = {"elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "crelu": nn.ReLU, "lrelu": nn.LeakyReLU,
             "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
    if act_name not in table:
        print("invalid activation function!")
        return None
    return table[act_name]()


def _mlp(dims: List[int], act_name: str, last_act: bool) -> nn.Sequential:
    """Linear/activation stack whose module indices match the reference's nn.Sequential layout
    (Linear at 0, 2, 4, ...)."""
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2 or last_act:
            layers.append(get_activation(act_name))
    return nn.Sequential(*layers)


def _stack_spec(seq: nn.Sequential):
    """[(weight, bias)], [act] of a Linear/ELU/ReLU stack, or None if the Sequential holds anything else."""
    mods = list(seq)
    layers, acts, i = [], [], 0
    while i < len(mods):
        m = mods[i]
        if not isinstance(m, nn.Linear):
            return None
        act = None
        if i + 1 < len(mods) and isinstance(mods[i + 1], (nn.ELU, nn.ReLU)):
            act = "elu" if isinstance(mods[i + 1], nn.ELU) else "relu"
            i += 1
        layers.append((m.weight, m.bias))
        acts.append(act)
        i += 1
    return layers, acts


def run_mlp(seq: nn.Sequential, x: torch.Tensor, head: nn.Linear = None) -> torch.Tensor:
    """Evaluates a Linear/activation stack (optionally followed by a linear `head`) as one GEMM chain with the bias and
    activation fused into each GEMM's epilogue."""
    spec = _stack_spec(seq)
    if spec is None:
        x = seq(x)
        return x if head is None else linear_act(x, head.weight, head.bias, None)
    layers, acts = spec
    if head is not None:
        layers, acts = layers + [(head.weight, head.bias)], acts + [None]
    return mlp_chain(x, layers, acts)


class FlatParams:
    """Re-homes the parameters of `module` into ONE contiguous fp32 buffer (+ a gradient twin).
    `module.state_dict()` / `load_state_dict()` keep working: the Parameters are views."""

    def __init__(self, module: nn.Module):
        """Every tensor starts on a 16-byte boundary and the rows of 2-D weights are padded to a multiple of 4
        floats (the padding stays zero: zero gradient, zero Adam update), so that the tcgen05 GEMM can read the
        weights straight out of the flat buffer through TMA.  The Parameters are (possibly strided) views."""
        params = [p for p in module.parameters()]
        names = {id(p): n for n, p in module.named_parameters()}
        dev = params[0].device
        plan, off = [], 0
        for p in params:
            if p.dim() == 2:
                rows, k = p.shape
                kp = (k + 3) // 4 * 4
                n = rows * kp
            else:
                n = (p.numel() + 3) // 4 * 4
            plan.append((p, off, n))
            off += n
        self.numel = off
        self.data = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.slices = {}

        def view(buf, p, o, n):
            if p.dim() == 2:
                rows, k = p.shape
                return buf[o:o + n].view(rows, n // rows)[:, :k]
            return buf[o:o + p.numel()].view(p.shape)

        for p, o, n in plan:
            view(self.data, p, o, n).copy_(p.data)
            p.data = view(self.data, p, o, n)
            p.grad = view(self.grad, p, o, n)
            self.slices[names[id(p)]] = (o, n)
        self._plan, self._view = plan, view

    def zero_grad(self):
        self.grad.zero_()

    def rebind_grad(self, buf: torch.Tensor) -> None:
        """Move the gradient twin into `buf` (a contiguous fp32 slice of `numel` elements, e.g. of an arena shared with other
        parameter sets so that ONE collective reduces them all); every Parameter's `.grad` becomes a view of it."""
        if buf.numel() != self.numel or buf.dtype != torch.float32 or not buf.is_contiguous():
            raise ValueError("rebind_grad: need a contiguous fp32 buffer of the flat size")
        buf.copy_(self.grad)
        self.grad = buf
        for p, o, n in self._plan:
            p.grad = self._view(buf, p, o, n)


class StateHistoryEncoder(nn.Module):
    """Linear(57->30)+act per time step, Conv1d(30->20,k4,s2)+act, Conv1d(20->10,k2)+act, Linear(30->out)+act
    (actor_critic.py:9-59; tsteps = 10 is the shipped configuration)."""

    def __init__(self, activation_fn, input_size, tsteps, output_size, tanh_encoder_output=False):
        super().__init__()
        self.activation_fn = activation_fn
        self.tsteps = tsteps
        ch = 10
        self.encoder = nn.Sequential(nn.Linear(input_size, 3 * ch), self.activation_fn)
        spec = {50: [(3 * ch, 2 * ch, 8, 4), (2 * ch, ch, 5, 1), (ch, ch, 5, 1)],
                10: [(3 * ch, 2 * ch, 4, 2), (2 * ch, ch, 2, 1)],
                20: [(3 * ch, 2 * ch, 6, 2), (2 * ch, ch, 4, 2)]}
        if tsteps not in spec:
            raise ValueError("tsteps must be 10, 20 or 50")
        conv = []
        for cin, cout, k, s in spec[tsteps]:
            conv += [nn.Conv1d(cin, cout, kernel_size=k, stride=s), self.activation_fn]
        conv.append(nn.Flatten())
        self.conv_layers = nn.Sequential(*conv)
        self.linear_output = nn.Sequential(nn.Linear(ch * 3, output_size), self.activation_fn)

    def forward(self, obs):
        nd, T = obs.shape[0], self.tsteps
        proj = self.encoder(obs.reshape(nd * T, -1))
        out = self.conv_layers(proj.reshape(nd, T, -1).permute(0, 2, 1))
        return self.linear_output(out)


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, num_prop, num_hist, num_explicit, num_latent,
                 num_command, actor_hidden_dims=[256, 256, 256], critic_hidden_dims=[256, 256, 256],
                 priv_encoder_dims=[256, 256], activation='elu', init_noise_std=1.0, fixed_std=False,
                 train_with_estimated_latent=False, **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs.keys())))
        super().__init__()
        self.num_actor_obs, self.num_critic_obs = num_actor_obs, num_critic_obs
        self.train_with_estimated_latent = train_with_estimated_latent
        self.num_prop, self.num_explicit, self.num_latent = num_prop, num_explicit, num_latent
        self.num_hist, self.num_command = num_hist, num_command
        self.activation_name = activation
        # registration order == the reference's, so state_dict() key order (and the flat layout) match
        std = init_noise_std * torch.ones(num_actions)
        self.fixed_std = fixed_std
        if fixed_std:
            self.std = std.clone()
        else:
            self.std = nn.Parameter(std)
        if len(priv_encoder_dims) > 0:
            self.priv_encoder = _mlp([num_latent] + list(priv_encoder_dims) + [num_latent], activation, last_act=True)
        else:
            self.priv_encoder = nn.Identity()
        self.history_encoder = StateHistoryEncoder(get_activation(activation), num_prop, num_hist, num_latent)
        self.actor_trunk = _mlp([num_actor_obs] + list(actor_hidden_dims), activation, last_act=True)
        self.actor_head = nn.Linear(actor_hidden_dims[-1], num_actions)
        self.critic_trunk = _mlp([num_critic_obs] + list(critic_hidden_dims), activation, last_act=True)
        self.critic_head = nn.Linear(critic_hidden_dims[-1], 1)
        for m in self.actor_trunk.modules():                       # actor_critic.py:131-135
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)
        self.distribution = None
        self._mean = self._std = None
        self.flat = None

    # ---- flat storage ---------------------------------------------------------------------------------
    def flatten_parameters(self) -> FlatParams:
        self.flat = FlatParams(self)
        return self.flat

    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    # ---- distribution ----------------------------------------------------------------------------------
    @property
    def action_mean(self):
        return self._mean

    @property
    def action_std(self):
        return self._std

    @property
    def entropy(self):
        # Normal.entropy().sum(-1) = sum(0.5 + 0.5 log(2 pi) + log(std))
        return (1.4189385332046727 + torch.log(self._std)).sum(dim=-1)

    def _split(self, observations):
        p, e, l, h = self.num_prop, self.num_explicit, self.num_latent, self.num_hist * self.num_prop
        return (observations[:, :p], observations[:, p:p + e], observations[:, p + e:p + e + l],
                observations[:, p + e + l:p + e + l + h], observations[:, p + e + l + h:])

    def _actor_mean(self, observations, hist_encoding: bool, priv_latent=None):
        obs_prop, obs_explicit, obs_latent, obs_hist, obs_command = self._split(observations)
        if self.train_with_estimated_latent:
            if hist_encoding:
                obs_latent = self.infer_hist_latent(obs_hist)
            else:                                         # a caller that also needs the latent may hand it in (one encoder pass)
                obs_latent = self.infer_priv_latent(obs_latent) if priv_latent is None else priv_latent
        # [prop | explicit | latent | command | zero pad]: ONE concatenation whose row pitch is a multiple of 16 bytes, so
        # the result is a legal TMA operand for the first actor layer (the pad columns are sliced off again)
        n_in = self.num_actor_obs
        pad = (n_in + 3) // 4 * 4 - n_in
        parts = [obs_prop, obs_explicit, obs_latent, obs_command]
        if pad:
            parts.append(torch.zeros(observations.shape[0], pad, device=observations.device, dtype=observations.dtype))
        x = torch.cat(parts, dim=-1)[:, :n_in]
        return run_mlp(self.actor_trunk, x, head=self.actor_head)

    @staticmethod
    def init_weights(sequential, scales):
        """Orthogonal re-initialisation of a stack's Linear layers, gain per layer (actor_critic.py:147-151; unused there too)."""
        for gain, layer in zip(scales, (m for m in sequential if isinstance(m, nn.Linear))):
            nn.init.orthogonal_(layer.weight, gain=gain)

    def update_distribution(self, observations, hist_encoding: bool, priv_latent=None):
        mean = self._actor_mean(observations, hist_encoding, priv_latent)
        self._mean = mean
        # Normal(mean, mean*0. + std) (actor_critic.py:189-190): the broadcast is a stride-0 view, not a materialised tensor
        self._std = self.std.to(mean.device).expand_as(mean)

    def act(self, observations, hist_encoding=False, **kwargs):
        self.update_distribution(observations, hist_encoding)
        noise = kwargs.get("normal_draw")
        if noise is None:
            noise = torch.randn_like(self._mean)
        return (self._mean + self._std * noise).detach()              # Normal(mean, std).sample()

    def get_actions_log_prob(self, actions):
        var = self._std ** 2
        return (-((actions - self._mean) ** 2) / (2 * var) - torch.log(self._std) - 0.9189385332046727).sum(dim=-1)

    def act_inference(self, observations, hist_encoding=True):
        return self._actor_mean(observations, hist_encoding)

    def infer_priv_latent(self, obs):
        return run_mlp(self.priv_encoder, obs) if not isinstance(self.priv_encoder, nn.Identity) else obs

    def infer_hist_latent(self, obs):
        enc = self.history_encoder
        fused_ok = (obs.is_cuda and not torch.is_grad_enabled() and obs.dim() == 2 and obs.stride(1) == 1
                    and enc.tsteps == 10 and self.activation_name == "elu" and self.num_prop == 57)
        if fused_ok:                                      # K11: forward only (inference / no-grad uses)
            from .. import ops
            out = torch.empty(obs.shape[0], self.num_latent, device=obs.device, dtype=torch.float32)
            ops.hist_encoder_fwd(obs, enc, out)
            return out
        return enc(obs.reshape(-1, self.num_hist, self.num_prop))

    def evaluate(self, critic_observations, **kwargs):
        return run_mlp(self.critic_trunk, critic_observations, head=self.critic_head)


class Estimator(nn.Module):
    """MLP input_dim -> hidden_dims -> output_dim, no activation on the output (estimator.py:12-40)."""

    def __init__(self, input_dim, output_dim, hidden_dims=[256, 128, 64], activation="elu", **kwargs):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.estimator = _mlp([input_dim] + list(hidden_dims) + [output_dim], activation, last_act=False)
        self.flat = None

    def flatten_parameters(self) -> FlatParams:
        self.flat = FlatParams(self)
        return self.flat

    def forward(self, input):
        return run_mlp(self.estimator, input)

    def inference(self, input):
        with torch.no_grad():
            return run_mlp(self.estimator, input)


class Discriminator(nn.Module):
    """Trunk 98->512->256 (ReLU) + heads d(1) / classifier(dim_c) / encoder_eps(1); per-step style reward
    `predict_disc_reward` (discriminator.py:12-118); its update is `SSInfoGAIL.update_ss_info_gail`."""

    def __init__(self, env, input_dim, num_disc_obs, dim_c, dt, disc_loss_function, reward_i_normalizer,
                 reward_i_coef, reward_us_coef, reward_ss_coef, reward_t_coef, disc_history_len, disc_obs_len,
                 obs_disc_weight_step, hidden_units, device):
        super().__init__()
        self.device, self.env = device, env
        self.input_dim, self.num_disc_obs, self.dim_c, self.dt = input_dim, num_disc_obs, dim_c, dt
        self.disc_loss_function = disc_loss_function
        self.reward_i_normalizer = reward_i_normalizer
        self.disc_history_len, self.disc_obs_len = disc_history_len, disc_obs_len
        self.obs_disc_weight_step = obs_disc_weight_step
        self.reward_i_coef, self.reward_us_coef = reward_i_coef, reward_us_coef
        self.reward_ss_coef, self.reward_t_coef = reward_ss_coef, reward_t_coef
        self.trunk = _mlp([input_dim] + list(hidden_units), "relu", last_act=True)
        self.linear = nn.Linear(hidden_units[-1], 1)
        self.classifier = nn.Linear(hidden_units[-1], dim_c)
        self.encoder_eps = nn.Linear(hidden_units[-1], 1)
        for m in self.trunk.modules():
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)
        nn.init.uniform_(self.linear.weight, -1.0, 1.0)
        nn.init.zeros_(self.linear.bias)

    def flatten_parameters(self) -> FlatParams:
        self.flat = FlatParams(self)
        return self.flat

    def forward_torch(self, x):
        """forward() on plain torch ops: the discriminator update needs double backward (gradient penalty, gail.py:492-502),
        which the tcgen05 autograd nodes do not provide; its batches are 1 228 rows, launch-bound either way."""
        h = F.relu(F.linear(F.relu(F.linear(x, self.trunk[0].weight, self.trunk[0].bias)), self.trunk[2].weight, self.trunk[2].bias))
        d = F.linear(h, self.linear.weight, self.linear.bias)
        eps = F.linear(h, self.encoder_eps.weight, self.encoder_eps.bias)
        c = torch.softmax(F.linear(h, self.classifier.weight, self.classifier.bias), -1)
        return d, eps, torch.clamp(c, 1e-20, torch.inf)

    def forward(self, x):
        x = run_mlp(self.trunk, x)
        d = linear_act(x, self.linear.weight, self.linear.bias, None)
        eps = linear_act(x, self.encoder_eps.weight, self.encoder_eps.bias, None)
        c = torch.softmax(linear_act(x, self.classifier.weight, self.classifier.bias, None), -1)
        return d, eps, torch.clamp(c, 1e-20, torch.inf)

    def heads_forward(self, x):
        """[d | eps | classifier logits | 0] (M, 8) from the normalised input: trunk on the GEMM chain, the three heads as ONE
        GEMM over their concatenated weights (padded to 8 rows so the output row pitch is TMA-legal)."""
        h = run_mlp(self.trunk, x)
        w = torch.cat([self.linear.weight, self.encoder_eps.weight, self.classifier.weight,
                       torch.zeros(1, self.linear.weight.shape[1], device=x.device)], dim=0)
        b = torch.cat([self.linear.bias, self.encoder_eps.bias, self.classifier.bias, torch.zeros(1, device=x.device)])
        return linear_act(h, w, b, None)

    def predict_disc_reward(self, reward_t, obs, obs_disc, normalizer=None):
        """Returns (rewards, reward_i, reward_us, reward_ss, reward_t); like the reference, the semi-supervised
        term goes through a float64 cross-entropy on the already soft-maxed classifier output (:68-69, :108),
        so `rewards` and `reward_ss` are float64."""
        label_eps = obs[:, -self.dim_c - 1].clone().unsqueeze(-1)
        label_c = F.one_hot(torch.argmax(obs[:, -self.dim_c:], dim=-1), num_classes=self.dim_c)
        od = obs_disc.clone()
        if self.env.task_obs_weight_decay:
            od[:, :, 3:9] *= self.env.task_obs_weight
            od[:, :, 33:] *= self.env.task_obs_weight
        od = od[:, -self.disc_obs_len:, :].reshape(len(od), -1)
        if self.obs_disc_weight_step != 0.0:
            mult = (torch.arange(self.disc_obs_len, dtype=torch.float32, device=od.device) *
                    self.obs_disc_weight_step + 1).repeat_interleave(self.num_disc_obs)
            od = od * mult
        with torch.no_grad():
            if normalizer is not None:
                od = normalizer.normalize_torch(od, od.device)
            d, eps, c = self.forward(od)
            if self.disc_loss_function == "BCEWithLogitsLoss":
                reward_i = -torch.log(torch.maximum(1 - 1 / (1 + torch.exp(-d)), torch.tensor(0.0001, device=d.device)))
            elif self.disc_loss_function == "MSELoss":
                reward_i = torch.clamp(1 - (1 / 4) * torch.square(d - 1), min=0)
            elif self.disc_loss_function == "WassersteinLoss":             # :97-99: running normalisation of the critic output
                reward_i = self.reward_i_normalizer.normalize_torch(d, d.device)
                if hasattr(self.reward_i_normalizer, "update_torch"):
                    self.reward_i_normalizer.update_torch(d)
                else:
                    self.reward_i_normalizer.update(d.cpu().numpy())
            else:
                raise ValueError("Unexpected style reward mapping specified")
            reward_us = -torch.abs(eps - label_eps)
            reward_ss = -F.cross_entropy(c, label_c.to(torch.float64), reduction="none").unsqueeze(1)
            reward_i = reward_i * self.dt
            reward_us = reward_us * self.dt
            reward_ss = reward_ss * self.dt
            rewards = (self.reward_i_coef * reward_i + self.reward_us_coef * reward_us +
                       self.reward_ss_coef * reward_ss + self.reward_t_coef * reward_t)
        return rewards.squeeze(), reward_i.squeeze(), reward_us.squeeze(), reward_ss.squeeze(), reward_t.squeeze()

    def get_disc_logit_weights(self):
        return torch.flatten(self.linear.weight)

    def get_disc_weights(self):
        ws = [torch.flatten(m.weight) for m in self.trunk.modules() if isinstance(m, nn.Linear)]
        ws.append(torch.flatten(self.linear.weight))
        return ws


class ActorCriticBBC(ActorCritic):
    """The frozen low-level controller as the TSC fork constructs it (tsc/rsl_rl/modules/actor_critic.py:286-447): the BBC
    `ActorCritic` with the fork's argument list -- `num_prop` INCLUDES the auxiliary lanes and `num_auxiliary` is subtracted
    (:312) -- and the fork's default `train_with_estimated_latent=True` (the controller acts on the history-encoder latent).
    Same parameters / `state_dict` keys as `ActorCritic`: a BBC checkpoint's `actor_critic` loads into it (`load_bbc`)."""

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, num_prop, num_auxiliary, num_hist, num_explicit, num_latent,
                 num_command, actor_hidden_dims=[256, 256, 256], critic_hidden_dims=[256, 256, 256], priv_encoder_dims=[256, 256],
                 activation='elu', init_noise_std=1.0, fixed_std=False, train_with_estimated_latent=True, **kwargs):
        super().__init__(num_actor_obs, num_critic_obs, num_actions, num_prop - num_auxiliary, num_hist, num_explicit, num_latent,
                         num_command, actor_hidden_dims=actor_hidden_dims, critic_hidden_dims=critic_hidden_dims,
                         priv_encoder_dims=priv_encoder_dims, activation=activation, init_noise_std=init_noise_std,
                         fixed_std=fixed_std, train_with_estimated_latent=train_with_estimated_latent, **kwargs)


class DiscriminatorTSC(Discriminator):
    """The TSC fork's constructor and call signature (tsc/rsl_rl/algorithms/discriminator.py:12-110): no env argument, no
    task-observation weighting, the running normaliser is a member and `predict_disc_reward(reward_t, obs, obs_disc)` uses it.
    Same parameters / `state_dict` keys as the BBC class, so a BBC checkpoint's `disc` loads into it (`load_bbc`, :647-660)."""

    def __init__(self, input_dim, num_disc_obs, dim_c, dt, disc_loss_function, reward_i_normalizer, reward_i_coef,
                 reward_us_coef, reward_ss_coef, reward_t_coef, disc_obs_len, hidden_units, normalizer, device):
        env = types.SimpleNamespace(task_obs_weight_decay=False, task_obs_weight=1.0, dim_c=dim_c)
        super().__init__(env, input_dim, num_disc_obs, dim_c, dt, disc_loss_function, reward_i_normalizer, reward_i_coef,
                         reward_us_coef, reward_ss_coef, reward_t_coef, disc_obs_len, disc_obs_len, 0.0, hidden_units, device)
        self.normalizer = normalizer

    def predict_disc_reward(self, reward_t, obs, obs_disc, normalizer=None):
        return super().predict_disc_reward(reward_t, obs, obs_disc, normalizer=self.normalizer if normalizer is None else normalizer)

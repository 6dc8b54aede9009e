"""CPU ORACLE (test infrastructure, not product code) for the rsl_rl trainer half of the hot path.

Plain-torch restatements, each citing the reference lines it follows (all under
/root/reference/bbc/rsl_rl unless stated).  Pinned by `oracle/gen_golden_trainer.py`, which runs the
UNMODIFIED reference classes (RolloutStorage, ActorCritic, Estimator, Discriminator, SSInfoGAIL) on the
same seeded inputs in the build container, asserts agreement with this file and commits the vectors
to `tests/golden/trainer_*.npz`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may
import this module.
"""
import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------
# a14  RolloutStorage.compute_returns  (storage/rollout_storage.py:97-111)
# ------------------------------------------------------------------------------------------

def compute_returns(rewards, values, dones, last_values, gamma: float, lam: float):
    """rewards/values (T,N,1) f32, dones (T,N,1) u8, last_values (N,1).  Returns (returns, advantages)."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        next_is_not_terminal = 1.0 - dones[step].float()
        delta = rewards[step] + next_is_not_terminal * gamma * next_values - values[step]
        advantage = delta + next_is_not_terminal * gamma * lam * advantage
        returns[step] = advantage + values[step]
    advantages = returns - values
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages


# ------------------------------------------------------------------------------------------
# a13  RolloutStorage.mini_batch_generator index plan  (storage/rollout_storage.py:122-157)
# ------------------------------------------------------------------------------------------

def mini_batch_slices(indices: torch.Tensor, num_mini_batches: int, num_epochs: int):
    """ONE permutation per update, the same `num_mini_batches` slices reused in every epoch."""
    mb = indices.numel() // num_mini_batches
    for _ in range(num_epochs):
        for i in range(num_mini_batches):
            yield indices[i * mb:(i + 1) * mb]


# ------------------------------------------------------------------------------------------
# a16  ActorCritic / Estimator forward  (modules/actor_critic.py:62-225, modules/estimator.py:12-40)
#      functional form over a state_dict with the reference's parameter names
# ------------------------------------------------------------------------------------------
NUM_PROP, NUM_EXPLICIT, NUM_LATENT, NUM_HIST, NUM_COMMAND = 57, 4, 29, 10, 11


def _mlp(x, sd, prefix: str, idxs: List[int], last_act: bool, act=F.elu):
    for k, i in enumerate(idxs):
        x = F.linear(x, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"])
        if k < len(idxs) - 1 or last_act:
            x = act(x)
    return x


def infer_priv_latent(sd, obs_latent):
    """priv_encoder: Linear(29,64) ELU Linear(64,29) ELU  (actor_critic.py:96-108)."""
    return _mlp(obs_latent, sd, "priv_encoder", [0, 2], last_act=True)


def infer_hist_latent(sd, obs_hist):
    """StateHistoryEncoder tsteps=10 (actor_critic.py:9-59)."""
    nd = obs_hist.shape[0]
    x = obs_hist.reshape(nd * NUM_HIST, NUM_PROP)
    x = F.elu(F.linear(x, sd["history_encoder.encoder.0.weight"], sd["history_encoder.encoder.0.bias"]))
    x = x.reshape(nd, NUM_HIST, -1).permute(0, 2, 1)
    x = F.elu(F.conv1d(x, sd["history_encoder.conv_layers.0.weight"], sd["history_encoder.conv_layers.0.bias"], stride=2))
    x = F.elu(F.conv1d(x, sd["history_encoder.conv_layers.2.weight"], sd["history_encoder.conv_layers.2.bias"], stride=1))
    x = x.flatten(1)
    return F.elu(F.linear(x, sd["history_encoder.linear_output.0.weight"], sd["history_encoder.linear_output.0.bias"]))


def actor_mean(sd, observations, hist_encoding: bool, train_with_estimated_latent: bool = True):
    """update_distribution / act_inference (actor_critic.py:171-214): returns the action mean."""
    p, e, l, h = NUM_PROP, NUM_EXPLICIT, NUM_LATENT, NUM_HIST * NUM_PROP
    obs_prop = observations[:, :p]
    obs_explicit = observations[:, p:p + e]
    obs_latent = observations[:, p + e:p + e + l]
    obs_hist = observations[:, p + e + l:p + e + l + h]
    obs_command = observations[:, p + e + l + h:]
    if train_with_estimated_latent:
        obs_latent = infer_hist_latent(sd, obs_hist) if hist_encoding else infer_priv_latent(sd, obs_latent)
    x = torch.cat([obs_prop, obs_explicit, obs_latent, obs_command], dim=-1)
    x = _mlp(x, sd, "actor_trunk", [0, 2, 4], last_act=True)
    return F.linear(x, sd["actor_head.weight"], sd["actor_head.bias"])


def critic_value(sd, critic_obs):
    """evaluate (actor_critic.py:222-225)."""
    x = _mlp(critic_obs, sd, "critic_trunk", [0, 2, 4], last_act=True)
    return F.linear(x, sd["critic_head.weight"], sd["critic_head.bias"])


def estimator_forward(sd, obs_prop):
    """Estimator 57->128->64->4 (estimator.py:24-36)."""
    return _mlp(obs_prop, sd, "estimator", [0, 2, 4], last_act=False)


def normal_log_prob(actions, mean, std):
    """torch.distributions.Normal.log_prob summed over the action dim (actor_critic.py:195-196)."""
    var = std ** 2
    return (-((actions - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(dim=-1)


def normal_entropy(std):
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)).sum(dim=-1)


def act(sd_ac, sd_est, obs, critic_obs, normal_draw, hist_encoding: bool = False):
    """SSInfoGAIL.act (algorithms/gail.py:176-197) with the N(0,1) draw injected.
    Returns dict(actions, values, actions_log_prob, action_mean, action_sigma)."""
    obs_est = obs.clone()
    obs_est[:, NUM_PROP:NUM_PROP + NUM_EXPLICIT] = estimator_forward(sd_est, obs_est[:, :NUM_PROP])
    mean = actor_mean(sd_ac, obs_est, hist_encoding)
    std = mean * 0. + sd_ac["std"]
    actions = mean + std * normal_draw                       # Normal.sample(): loc + eps * scale
    values = critic_value(sd_ac, critic_obs)
    return dict(actions=actions, values=values, actions_log_prob=normal_log_prob(actions, mean, std),
                action_mean=mean, action_sigma=std)


# ------------------------------------------------------------------------------------------
# a12  Discriminator.predict_disc_reward  (algorithms/discriminator.py:64-118) + Normalizer.normalize_torch
#      (utils/utils.py:97-103)
# ------------------------------------------------------------------------------------------

def predict_disc_reward(sd_disc, reward_t, obs, obs_disc_hist, norm_mean, norm_var, dt: float,
                        task_obs_weight: float, task_obs_weight_decay: bool = True,
                        coefs: Tuple[float, float, float, float] = (1.0, 0.01, 0.2, 0.2),
                        dim_c: int = 5, disc_obs_len: int = 2, obs_disc_weight_step: float = 0.0,
                        norm_eps: float = 1e-4, clip_obs: float = 10.0):
    """reward_t (N,1); obs (N,671); obs_disc_hist (N,2,49); norm_mean/var float64 numpy-like tensors.
    Returns the reference's 5-tuple; `rewards` and `reward_ss` are float64 like the reference (:108)."""
    label_eps = obs[:, -dim_c - 1].clone().unsqueeze(-1)
    label_c = F.one_hot(torch.argmax(obs[:, -dim_c:], dim=-1), num_classes=dim_c)
    od = obs_disc_hist.clone()
    if task_obs_weight_decay:
        od[:, :, 3:9] *= task_obs_weight
        od[:, :, 33:] *= task_obs_weight
    od = od[:, -disc_obs_len:, :].reshape(len(od), -1)
    mult = (torch.arange(disc_obs_len, dtype=torch.float32, device=od.device) * obs_disc_weight_step + 1)
    mult = mult.view(1, -1, 1).repeat(od.size(0), 1, od.shape[1] // disc_obs_len).view(len(od), -1)
    od = od * mult
    mean_t = torch.as_tensor(norm_mean, dtype=torch.float64).to(device=od.device, dtype=torch.float32)
    std_t = torch.sqrt((torch.as_tensor(norm_var, dtype=torch.float64) + norm_eps).to(device=od.device, dtype=torch.float32))
    x = torch.clamp((od - mean_t) / std_t, -clip_obs, clip_obs)
    x = _mlp(x, sd_disc, "trunk", [0, 2], last_act=True, act=F.relu)
    d = F.linear(x, sd_disc["linear.weight"], sd_disc["linear.bias"])
    eps = F.linear(x, sd_disc["encoder_eps.weight"], sd_disc["encoder_eps.bias"])
    c = torch.clamp(torch.softmax(F.linear(x, sd_disc["classifier.weight"], sd_disc["classifier.bias"]), -1),
                    1e-20, torch.inf)
    reward_i = torch.clamp(1 - (1 / 4) * torch.square(d - 1), min=0)                    # MSELoss mapping :97-98
    reward_us = -F.l1_loss(eps, label_eps, reduction="none")
    reward_ss = -F.cross_entropy(c, label_c.to(float), reduction="none").unsqueeze(1)   # f64, double softmax
    reward_i = reward_i * dt
    reward_us = reward_us * dt
    reward_ss = reward_ss * dt
    ci, cu, cs, ct = coefs
    rewards = ci * reward_i + cu * reward_us + cs * reward_ss + ct * reward_t
    return rewards.squeeze(), reward_i.squeeze(), reward_us.squeeze(), reward_ss.squeeze(), reward_t.squeeze()


# ------------------------------------------------------------------------------------------
# a15  SSInfoGAIL.update_actor_critic losses  (algorithms/gail.py:328-413)
# ------------------------------------------------------------------------------------------

def ppo_losses(sd_ac, sd_est, batch: Dict[str, torch.Tensor], clip_param=0.2, priv_reg_coef=0.0,
               surrogate_loss_coef=2.0, value_loss_coef=5.0, bounds_loss_coef=0.0, entropy_coef=0.01,
               use_clipped_value_loss=True):
    """Forward half of one PPO minibatch step.  `batch` keys: obs, critic_obs, actions, target_values,
    advantages, returns, old_actions_log_prob, old_mu, old_sigma.  Returns dict of the scalar losses,
    kl_mean, and the total `ppo_loss` / `estimator_loss` (differentiable w.r.t. sd tensors)."""
    obs = batch["obs"]
    mu = actor_mean(sd_ac, obs, hist_encoding=False)
    sigma = mu * 0. + sd_ac["std"]
    logp = normal_log_prob(batch["actions"], mu, sigma)
    value = critic_value(sd_ac, batch["critic_obs"])
    entropy = normal_entropy(sigma)
    p, e, l = NUM_PROP, NUM_EXPLICIT, NUM_LATENT
    priv_latent = infer_priv_latent(sd_ac, obs[:, p + e:p + e + l])
    with torch.no_grad():
        hist_latent = infer_hist_latent(sd_ac, obs[:, p + e + l:p + e + l + NUM_HIST * NUM_PROP])
    priv_reg_loss = (priv_latent - hist_latent.detach()).norm(p=2, dim=1).mean()
    est = estimator_forward(sd_est, obs[:, :p])
    estimator_loss = (est - obs[:, p:p + e]).pow(2).mean()
    with torch.no_grad():
        old_sigma, old_mu = batch["old_sigma"], batch["old_mu"]
        kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5) +
                       (torch.square(old_sigma) + torch.square(old_mu - mu)) / (2.0 * torch.square(sigma)) - 0.5, dim=-1)
        kl_mean = torch.mean(kl)
    adv = torch.squeeze(batch["advantages"])
    ratio = torch.exp(logp - torch.squeeze(batch["old_actions_log_prob"]))
    surrogate = -adv * ratio
    surrogate_clipped = -adv * torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param)
    surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
    if use_clipped_value_loss:
        tv = batch["target_values"]
        value_clipped = tv + (value - tv).clamp(-clip_param, clip_param)
        value_loss = torch.max((value - batch["returns"]).pow(2), (value_clipped - batch["returns"]).pow(2)).mean()
    else:
        value_loss = (batch["returns"] - value).pow(2).mean()
    soft_bound = 1.0
    mu_loss_high = torch.maximum(mu - soft_bound, torch.tensor(0, device=mu.device)) ** 2
    mu_loss_low = torch.minimum(mu + soft_bound, torch.tensor(0, device=mu.device)) ** 2
    b_loss = (mu_loss_low + mu_loss_high).sum(axis=-1)
    ppo_loss = (surrogate_loss_coef * surrogate_loss + value_loss_coef * value_loss +
                bounds_loss_coef * b_loss.mean() - entropy_coef * entropy.mean() + priv_reg_coef * priv_reg_loss)
    return dict(ppo_loss=ppo_loss, estimator_loss=estimator_loss, surrogate_loss=surrogate_loss,
                value_loss=value_loss, b_loss=b_loss.mean(), entropy=entropy.mean(), priv_reg_loss=priv_reg_loss,
                kl_mean=kl_mean, mu=mu, sigma=sigma, value=value)


def adaptive_lr(lr: float, kl_mean: float, desired_kl: float = 0.01) -> float:
    """gail.py:375-378."""
    if kl_mean > desired_kl * 2.0:
        return max(1e-5, lr / 1.5)
    if kl_mean < desired_kl / 2.0 and kl_mean > 0.0:
        return min(1e-2, lr * 1.5)
    return lr


def priv_reg_coef(counter: int, sched=(0, 0.1, 1000, 2000)) -> float:
    """gail.py:354-357."""
    stage = min(max((counter - sched[2]), 0) / sched[3], 1)
    return stage * (sched[1] - sched[0]) + sched[0]


# ------------------------------------------------------------------------------------------
# 8(f)-1  SSInfoGAIL.update_ss_info_gail  (algorithms/gail.py:415-541), MSELoss discriminator
# ------------------------------------------------------------------------------------------

def disc_forward(sd, x):
    """Discriminator.forward (discriminator.py:64-69): d, eps, clamped soft-maxed class probabilities."""
    h = _mlp(x, sd, "trunk", [0, 2], last_act=True, act=F.relu)
    d = F.linear(h, sd["linear.weight"], sd["linear.bias"])
    eps = F.linear(h, sd["encoder_eps.weight"], sd["encoder_eps.bias"])
    c = torch.softmax(F.linear(h, sd["classifier.weight"], sd["classifier.bias"]), -1)
    return d, eps, torch.clamp(c, 1e-20, torch.inf)


def disc_prepare(x, task_obs_weight, norm_mean, norm_var, disc_obs_len=2, num_disc_obs=49, obs_disc_weight_step=0.0,
                 task_obs_weight_decay=True, norm_eps=1e-4, clip_obs=10.0):
    """:423-452: task-obs weighting, per-step multipliers, normalisation of a (B, 98) batch of disc-obs histories."""
    x = x.view(len(x), disc_obs_len, -1).clone()
    if task_obs_weight_decay:
        x[:, :, 3:9] *= task_obs_weight
        x[:, :, 33:] *= task_obs_weight
    x = x[:, -disc_obs_len:, :].reshape(len(x), -1)
    mult = (torch.arange(disc_obs_len, dtype=torch.float32) * obs_disc_weight_step + 1)
    x = x * mult.view(1, -1, 1).repeat(len(x), 1, num_disc_obs).view(len(x), -1)
    mean_t = torch.as_tensor(norm_mean, dtype=torch.float64).to(torch.float32)
    std_t = torch.sqrt((torch.as_tensor(norm_var, dtype=torch.float64) + norm_eps).to(torch.float32))
    return torch.clamp((x - mean_t) / std_t, -clip_obs, clip_obs)


def disc_losses(sd, policy_state, policy_latent_eps, policy_latent_c, expert_lb, label_lb, expert_ulb, dim_c=5,
                ss_coef=1.0, info_max_coef_on=0.0, disc_coef=1.0, us_coef=1.0, disc_grad_penalty=0.1, disc_logit_reg=0.05,
                disc_weight_decay=0.0001):
    """:454-514 on already-normalised batches.  Returns the differentiable total loss and the logged terms."""
    _, _, pred_c_lb = disc_forward(sd, expert_lb)
    ss_loss = F.cross_entropy(pred_c_lb, label_lb)                                # CE on soft-maxed output (quirk)
    logits_pi, eps, pred_c = disc_forward(sd, policy_state)
    logits_exp, _, pred_c_ulb = disc_forward(sd, expert_ulb)
    pred_c_ulb_mean = torch.mean(pred_c_ulb, dim=0)
    info_max_loss = torch.mean(-torch.sum(pred_c_ulb * torch.log(pred_c_ulb + 1e-20), dim=-1))
    disc_exp_loss = torch.mean(F.mse_loss(logits_exp, torch.ones_like(logits_exp), reduction='none'))
    disc_pi_loss = torch.mean(F.mse_loss(logits_pi, -1 * torch.ones_like(logits_pi), reduction='none'))
    disc_loss = 0.5 * (disc_pi_loss + disc_exp_loss)
    us_loss = F.l1_loss(eps, policy_latent_eps)
    disc_logit_loss = torch.sum(torch.square(torch.flatten(sd["linear.weight"])))
    sample = expert_ulb.clone().requires_grad_(True)                              # gradient penalty :492-502
    h = _mlp(sample, sd, "trunk", [0, 2], last_act=True, act=F.relu)
    dd = F.linear(h, sd["linear.weight"], sd["linear.bias"])
    (g,) = torch.autograd.grad(dd, sample, grad_outputs=torch.ones_like(dd), create_graph=True, retain_graph=True)
    grad_pen_loss = torch.mean(torch.sum(torch.square(g), dim=-1))
    ws = torch.cat([torch.flatten(sd["trunk.0.weight"]), torch.flatten(sd["trunk.2.weight"]), torch.flatten(sd["linear.weight"])])
    weight_decay = torch.sum(torch.square(ws))
    loss = (ss_coef * ss_loss + info_max_coef_on * info_max_loss + disc_coef * disc_loss + us_coef * us_loss +
            disc_grad_penalty * grad_pen_loss + disc_logit_reg * disc_logit_loss + disc_weight_decay * weight_decay)
    with torch.no_grad():
        lab_pi = torch.argmax(policy_latent_c, dim=-1)
        acc = dict(acc_lb=torch.mean((torch.argmax(pred_c_lb, -1) == label_lb).float()),
                   acc_pi=(logits_pi < 0).float().mean(), acc_exp=(logits_exp > 0).float().mean(),
                   acc_ulb=torch.mean((torch.argmax(pred_c, -1) == lab_pi).float()))
    return dict(loss=loss, ss_loss=ss_loss, info_max_loss=info_max_loss, disc_loss=disc_loss, us_loss=us_loss,
                grad_pen_loss=grad_pen_loss, disc_logit_loss=disc_logit_loss, disc_weight_decay=weight_decay,
                pred_c_ulb_mean=pred_c_ulb_mean.detach(), **acc)


def disc_optimizers(sd, lr_disc=5e-4, lr_q=1e-3):
    """The three Adam optimisers of gail.py:107-128: the trunk is stepped by ALL of them, weight_decay 1e-3 each."""
    trunk = [sd[k] for k in ("trunk.0.weight", "trunk.0.bias", "trunk.2.weight", "trunk.2.bias")]
    grp = lambda ps: {'params': ps, 'weight_decay': 1e-3}                      # noqa: E731
    return (torch.optim.Adam([grp(trunk), grp([sd["linear.weight"], sd["linear.bias"]])], lr=lr_disc),
            torch.optim.Adam([grp(trunk), grp([sd["encoder_eps.weight"], sd["encoder_eps.bias"]])], lr=lr_q),
            torch.optim.Adam([grp(trunk), grp([sd["classifier.weight"], sd["classifier.bias"]])], lr=lr_q))


def normalizer_update(mean, var, count, arr):
    """RunningMeanStd.update (utils/utils.py:63-83) on float64 numpy-like tensors; `arr` float32 (B, D)."""
    bm = arr.mean(dim=0).double()                                                # np.mean of a float32 array is float32
    bv = arr.var(dim=0, unbiased=False).double()
    bc = arr.shape[0]
    delta = bm - mean
    tot = count + bc
    new_mean = mean + delta * bc / tot
    m2 = var * count + bv * bc + delta.square() * count * bc / (count + bc)
    return new_mean, m2 / (count + bc), tot

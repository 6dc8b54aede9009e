"""CPU ORACLE (test infrastructure, not product code) for the TSC trainer: `ActorCriticTSC` forward and the loss
graph of `PPO.update`.  Plain-torch restatements citing /root/reference/tsc/rsl_rl; pinned by
`oracle/gen_golden_tsc.py` against the UNMODIFIED reference classes (fixture tests/golden/tsc_trainer_seed3.npz).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this module.
"""
import math

import torch
import torch.nn.functional as F

from trainer import _mlp, adaptive_lr, normal_log_prob, priv_reg_coef  # noqa: F401  (shared definitions)

NUM_PROP, NUM_AUX, NUM_SCAN, NUM_EXPLICIT, NUM_LATENT, NUM_HIST = 65, 8, 132, 4, 29, 10
HIST_W = NUM_PROP - NUM_AUX                 # 57
NUM_ACT_D, NUM_ACT_C = 3, 6
EPS = torch.finfo(torch.float32).eps


def hist_latent(sd, obs):
    """Actor.infer_hist_latent (modules/actor_critic.py:160-162) + StateHistoryEncoder (:12-58)."""
    nd = obs.shape[0]
    x = obs[:, -NUM_HIST * HIST_W:].reshape(nd * NUM_HIST, HIST_W)
    p = "actor.history_encoder."
    x = F.elu(F.linear(x, sd[p + "encoder.0.weight"], sd[p + "encoder.0.bias"]))
    x = x.reshape(nd, NUM_HIST, -1).permute(0, 2, 1)
    x = F.elu(F.conv1d(x, sd[p + "conv_layers.0.weight"], sd[p + "conv_layers.0.bias"], stride=2))
    x = F.elu(F.conv1d(x, sd[p + "conv_layers.2.weight"], sd[p + "conv_layers.2.bias"], stride=1))
    return F.elu(F.linear(x.flatten(1), sd[p + "linear_output.0.weight"], sd[p + "linear_output.0.bias"]))


def priv_latent(sd, obs):
    """Actor.infer_priv_latent (:156-158)."""
    o = NUM_PROP + NUM_SCAN + NUM_EXPLICIT
    return _mlp(obs[:, o:o + NUM_LATENT], sd, "actor.priv_encoder", [0, 2], last_act=True)


def scan_latent(sd, obs):
    """Actor.scan_encoder: 132 -> 128 -> 64 -> 32, ELU, ELU, Tanh (:103-117)."""
    x = obs[:, NUM_PROP:NUM_PROP + NUM_SCAN]
    x = F.elu(F.linear(x, sd["actor.scan_encoder.0.weight"], sd["actor.scan_encoder.0.bias"]))
    x = F.elu(F.linear(x, sd["actor.scan_encoder.2.weight"], sd["actor.scan_encoder.2.bias"]))
    return torch.tanh(F.linear(x, sd["actor.scan_encoder.4.weight"], sd["actor.scan_encoder.4.bias"]))


def actor_embedding(sd, obs, hist_encoding: bool):
    """Actor.forward (:137-154)."""
    o = NUM_PROP + NUM_SCAN
    latent = hist_latent(sd, obs) if hist_encoding else priv_latent(sd, obs)
    x = torch.cat([obs[:, :NUM_PROP], scan_latent(sd, obs), obs[:, o:o + NUM_EXPLICIT], latent], dim=1)
    return _mlp(x, sd, "actor.actor_trunk", [0, 2, 4], last_act=True)


def heads(sd, emb):
    """logits of the mode head and mean of the continuous head (ActorCriticTSC.act, :250-259)."""
    return (F.linear(emb, sd["actor.actor_d.weight"], sd["actor.actor_d.bias"]),
            F.linear(emb, sd["actor.actor_c.weight"], sd["actor.actor_c.bias"]))


def critic_value(sd, critic_obs):
    return _mlp(critic_obs, sd, "critic", [0, 2, 4, 6], last_act=False)


def estimator_forward(sd, x):
    return _mlp(x, sd, "estimator", [0, 2, 4], last_act=False)


def categorical(prob):
    """torch.distributions.Categorical(probs=prob): normalised probs and the clamped log it works with
    (probs_to_logits: log(clamp(p, eps, 1-eps)))."""
    p = prob / prob.sum(-1, keepdim=True)
    return p, torch.log(p.clamp(min=EPS, max=1 - EPS))


def sample_mode(prob, u):
    """Inverse-CDF draw of the mode index from a uniform u (the injected replacement of Categorical.sample)."""
    cdf = torch.cumsum(prob, dim=-1)
    return torch.clamp((u.unsqueeze(-1) >= cdf).sum(-1), max=prob.shape[-1] - 1)


def act(sd_ac, sd_est, obs, critic_obs, normal_draw, mode_u, hist_encoding=False):
    """PPO.act (algorithms/ppo.py:101-125) with the random draws injected."""
    obs_est = obs.clone()
    o = NUM_PROP + NUM_SCAN                 # NB :108-110 writes at num_prop(57) + num_auxiliary(8) + num_scan
    obs_est[:, o:o + NUM_EXPLICIT] = estimator_forward(sd_est, obs_est[:, :HIST_W])
    emb = actor_embedding(sd_ac, obs_est, hist_encoding)
    logits, mean = heads(sd_ac, emb)
    prob = torch.softmax(logits, dim=-1)
    a_d = sample_mode(prob, mode_u)
    std = mean * 0. + sd_ac["std"]
    a_c = mean + std * normal_draw
    p, logit = categorical(prob)
    return dict(actions=torch.cat([a_d.unsqueeze(-1).to(mean.dtype), a_c], dim=-1), values=critic_value(sd_ac, critic_obs),
                actions_log_prob_d=logit.gather(-1, a_d.unsqueeze(-1)).squeeze(-1),
                actions_log_prob_c=normal_log_prob(a_c, mean, std), action_mean=mean, action_sigma=std, prob=prob)


def ppo_losses(sd_ac, sd_est, batch, clip_param=0.2, priv_reg_coef=0.0, value_loss_coef=1.0, entropy_coef=0.01,
               use_clipped_value_loss=True):
    """Forward half of one minibatch step of PPO.update (algorithms/ppo.py:159-262).  batch keys: obs, critic_obs,
    actions (M,19), target_values, advantages, returns, old_actions_log_prob_d, old_actions_log_prob_c, old_mu,
    old_sigma."""
    obs = batch["obs"]
    emb = actor_embedding(sd_ac, obs, hist_encoding=False)
    logits, mu = heads(sd_ac, emb)
    prob = torch.softmax(logits, dim=-1)
    p, logit = categorical(prob)
    a_d = batch["actions"][:, 0].to(torch.int64)
    logp_d = logit.gather(-1, a_d.unsqueeze(-1)).squeeze(-1)
    sigma = mu * 0. + sd_ac["std"]
    logp_c = normal_log_prob(batch["actions"][:, 1:], mu, sigma)
    value = critic_value(sd_ac, batch["critic_obs"])
    entropy_c = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).mean(dim=-1)          # :236-237 (mean, not sum)
    entropy_d = -(logit.clamp(min=torch.finfo(logit.dtype).min) * p).sum(-1)                 # Categorical.entropy
    entropy = entropy_c + entropy_d
    pl = priv_latent(sd_ac, obs)
    with torch.no_grad():
        hl = hist_latent(sd_ac, obs)
    priv_reg_loss = (pl - hl.detach()).norm(p=2, dim=1).mean()
    o = NUM_PROP + NUM_SCAN
    estimator_loss = (estimator_forward(sd_est, obs[:, :HIST_W]) - obs[:, o:o + NUM_EXPLICIT]).pow(2).mean()
    with torch.no_grad():
        osg, omu = batch["old_sigma"], batch["old_mu"]
        kl = torch.sum(torch.log(sigma / osg + 1.e-5) + (torch.square(osg) + torch.square(omu - mu)) /
                       (2.0 * torch.square(sigma)) - 0.5, axis=-1)
        kl_mean = torch.mean(kl)
    adv = torch.squeeze(batch["advantages"])

    def surrogate(logp, old):
        ratio = torch.exp(logp - torch.squeeze(old))
        return torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param)).mean()

    surrogate_loss = surrogate(logp_d, batch["old_actions_log_prob_d"]) + surrogate(logp_c, batch["old_actions_log_prob_c"])
    if use_clipped_value_loss:
        tv = batch["target_values"]
        vc = tv + (value - tv).clamp(-clip_param, clip_param)
        value_loss = torch.max((value - batch["returns"]).pow(2), (vc - batch["returns"]).pow(2)).mean()
    else:
        value_loss = (batch["returns"] - value).pow(2).mean()
    hi = torch.maximum(mu - 1.0, torch.tensor(0, device=mu.device)) ** 2
    lo = torch.minimum(mu + 1.0, torch.tensor(0, device=mu.device)) ** 2
    b_loss = (lo + hi).sum(axis=-1)
    loss = (surrogate_loss + value_loss_coef * value_loss - entropy_coef * entropy.mean() + priv_reg_coef * priv_reg_loss +
            0.0 * b_loss.mean())
    return dict(ppo_loss=loss, estimator_loss=estimator_loss, surrogate_loss=surrogate_loss, value_loss=value_loss,
                priv_reg_loss=priv_reg_loss, entropy=entropy.mean(), kl_mean=kl_mean, mu=mu, sigma=sigma, value=value,
                prob=prob)


def tsc_priv_reg_coef(counter, sched=(0, 0.1, 500, 1000)):
    """ppo.py:186-187 with the go2 agility schedule (legged_robot_config.py:399)."""
    return priv_reg_coef(counter, sched)


def update_dagger(sd_ac, obs, lr, epochs, max_grad_norm=1.0):
    """PPO.update_dagger (algorithms/ppo.py:284-314) over ONE minibatch holding every row of `obs` (the mean of the row norms
    does not depend on the generator's permutation): `epochs` Adam steps on the history encoder towards the frozen
    privileged-latent encoder.  Returns (mean loss, {name: updated tensor} of the history encoder)."""
    sd = {k: v.clone() for k, v in sd_ac.items()}
    enc_names = [k for k in sd if k.startswith("actor.history_encoder.")]
    for k in enc_names:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in enc_names], lr=lr)                  # hist_encoder_optimizer (:64)
    total = 0.0
    for _ in range(epochs):
        with torch.no_grad():
            target = priv_latent(sd, obs)
        loss = (target.detach() - hist_latent(sd, obs)).norm(p=2, dim=1).mean()
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_([sd[k] for k in enc_names], max_grad_norm)
        opt.step()
        total += float(loss.detach())
    return total / epochs, {k: sd[k].detach() for k in enc_names}

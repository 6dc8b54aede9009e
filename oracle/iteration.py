"""One full training iteration through the ORACLE port of the reference's PyTorch path (oracle/bbc_env.py + oracle/trainer.py).

TEST / BASELINE INFRASTRUCTURE ONLY.  Called by bench.py's baseline legs -- `cpu_baseline` and `--impl reference` on the host
cores, `torch_gpu_baseline` with `device` = the bench GPU (the reference's own single-GPU path of north_star's >= 4x target) --
never by the product package (it lived in qa_b200/pipeline.py in round 1 and was moved here: VERDICT r1 weak item 9).
Mirrors bbc/rsl_rl/runners/on_policy_runner.py:156-225 (rollout loop, compute_returns, update) over recorded state.
"""
import time

import torch

import bbc_env as O
import trainer as OT

CARRIED = ("last_actions", "last_torques_org", "last_dof_vel", "last_root_vel", "obs_history_buf", "episode_length_buf",
           "last_contacts", "commands", "latent_eps", "latent_c", "episode_sums", "feet_air_time", "obs_disc_buf",
           "action_history_buf")


def oracle_iteration(cfg, static, snaps, draws, table, weights, rollout_steps=None, minibatch_steps=20,
                         gamma=0.99, lam=0.95, num_mini_batches=4, device="cpu"):
    """The same iteration through the oracle port of the reference's PyTorch path (test infrastructure; called only by
    bench.py's baseline legs: cpu_baseline / --impl reference on the host cores, torch_gpu_baseline with `device` = the
    GPU, i.e. the reference's own single-GPU path of north_star's >= 4x target).  `rollout_steps` <= T env steps and
    `minibatch_steps` <= 20 PPO minibatch steps are executed (a bounded sample); every input must already live on `device`.
    Returns dict(t_rollout, t_gae, t_update, rollout_steps, minibatch_steps)."""
    dev = torch.device(device)
    if dev.type == "cuda":
        _zeros, _ones, _clock = torch.zeros, torch.ones, time.perf_counter

        def zeros(*a, **k):
            return _zeros(*a, device=dev, **k)

        def ones(*a, **k):
            return _ones(*a, device=dev, **k)

        def clock():
            torch.cuda.synchronize(dev)
            return _clock()
    else:
        zeros, ones, clock = torch.zeros, torch.ones, time.perf_counter
    T, N = len(snaps), cfg.num_envs
    R = T if rollout_steps is None else min(rollout_steps, T)
    g = torch.Generator().manual_seed(7)
    carried = {k: snaps[0][k].clone() for k in CARRIED}
    sd_ac = {k: v.clone().requires_grad_(True) for k, v in weights["ac"].items()}
    sd_est = {k: v.clone().requires_grad_(True) for k, v in weights["est"].items()}
    opt_a = torch.optim.Adam(list(sd_ac.values()), lr=1e-3)
    opt_e = torch.optim.Adam(list(sd_est.values()), lr=1e-4)
    W = 671
    st = dict(obs=zeros(T, N, W), actions=zeros(T, N, 12), rewards=zeros(T, N, 1),
              dones=zeros(T, N, 1, dtype=torch.uint8), values=zeros(T, N, 1), logp=zeros(T, N, 1),
              mu=zeros(T, N, 12), sigma=ones(T, N, 12))
    obs = zeros(N, W)
    disc_hist = torch.stack([carried["obs_disc_buf"]] * 2, dim=1)
    time_outs = zeros(N, dtype=torch.bool)
    t0 = clock()
    with torch.no_grad():
        for t in range(R):
            a = OT.act(sd_ac, sd_est, obs, obs, torch.randn(N, 12, generator=g).to(dev))
            hist, act = O.action_push(cfg, carried["action_history_buf"], a["actions"], delay=0)
            s = dict(snaps[t])
            s.update({k: carried[k] for k in CARRIED})
            s["action_history_buf"], s["actions"] = hist, act
            for _ in range(cfg.decimation):
                _, s["torques_org"] = O.compute_torques(cfg, {**static, "dof_state": s["dof_state"]}, act.clone())
            out = O.post_physics_step(cfg, static, s, draws[t], table, t + 1)
            done = out["reset_buf"]
            with_term = torch.where(done[:, None], carried["obs_disc_buf"], out["obs_disc_buf"])
            disc_hist = torch.stack([disc_hist[:, 1], with_term], dim=1)
            rew = OT.predict_disc_reward(weights["disc"], out["rew_buf"].unsqueeze(1), obs, disc_hist,
                                         weights["norm_mean"], weights["norm_var"], cfg.dt, 1.0)[0]
            if bool(done.any()):
                time_outs = out["time_out_buf"]
            rew = rew + gamma * torch.squeeze(a["values"] * time_outs.unsqueeze(1), 1)       # gail.py:203-205
            st["obs"][t], st["actions"][t], st["rewards"][t, :, 0] = obs, a["actions"], rew.float()
            st["dones"][t, :, 0], st["values"][t], st["logp"][t, :, 0] = done, a["values"], a["actions_log_prob"]
            st["mu"][t], st["sigma"][t] = a["action_mean"], a["action_sigma"]
            disc_hist = torch.where(done[:, None, None], out["obs_disc_buf"].unsqueeze(1).expand(-1, 2, -1), disc_hist)
            obs = out["obs_buf"]
            for k in CARRIED:
                carried[k] = out[k]
    t1 = clock()
    with torch.no_grad():
        last_values = OT.critic_value(sd_ac, obs)
        returns, adv = OT.compute_returns(st["rewards"], st["values"], st["dones"], last_values, gamma, lam)
    t2 = clock()
    flat = lambda x: x.flatten(0, 1)                                                        # noqa: E731
    idx = torch.randperm(T * N, generator=g).to(dev)
    mb = (T * N) // num_mini_batches
    lr = 1e-3
    for k in range(minibatch_steps):
        i = idx[(k % num_mini_batches) * mb:((k % num_mini_batches) + 1) * mb]
        batch = dict(obs=flat(st["obs"])[i], critic_obs=flat(st["obs"])[i], actions=flat(st["actions"])[i],
                     target_values=flat(st["values"])[i], advantages=flat(adv)[i], returns=flat(returns)[i],
                     old_actions_log_prob=flat(st["logp"])[i], old_mu=flat(st["mu"])[i], old_sigma=flat(st["sigma"])[i])
        L = OT.ppo_losses(sd_ac, sd_est, batch)
        opt_e.zero_grad()
        L["estimator_loss"].backward()
        torch.nn.utils.clip_grad_norm_(list(sd_est.values()), 1.0)
        opt_e.step()
        lr = OT.adaptive_lr(lr, float(L["kl_mean"]))
        for pg in opt_a.param_groups:
            pg["lr"] = lr
        opt_a.zero_grad()
        L["ppo_loss"].backward()
        torch.nn.utils.clip_grad_norm_(list(sd_ac.values()), 1.0)
        opt_a.step()
    t3 = clock()
    return dict(t_rollout=t1 - t0, t_gae=t2 - t1, t_update=t3 - t2, rollout_steps=R, minibatch_steps=minibatch_steps)

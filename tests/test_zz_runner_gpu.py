"""Device tests of the runner level (all green on the B200 since the round-1 driver run, GPUTEST_r01; host halves -- deque
accounting, tag list, optimiser-dict codec, the runner's learn loop, the batched discriminator step's gradients -- are covered
on CPU in tests/test_train_log.py, test_checkpoint.py, test_runner_host.py, test_disc_batched.py).

* `SSInfoGAIL.update_actor_critic(sample)` and one full-size PPO minibatch step (BASELINE config 0) against golden / oracle;
* `OnPolicyRunner.learn` with a log_dir: episode bookkeeping, the reference's scalar tags, checkpoint round trip
  (bbc/rsl_rl/runners/on_policy_runner.py:118-339);
* the opt-in batched discriminator step against the default one; K2 against the oracle at 32 768 envs in one launch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_update_actor_critic_entry_point_matches_reference_golden():
    """`SSInfoGAIL.update_actor_critic(sample)` (gail.py:328-413), the reference's per-minibatch entry point, on the golden
    minibatch of tests/test_trainer_gpu.py::test_ppo_minibatch_step_matches_reference_golden."""
    from helpers import assert_close
    from qa_b200 import synthetic
    from test_trainer_gpu import NET_ATOL, NET_RTOL, build, load_golden
    torch.backends.cuda.matmul.allow_tf32 = False
    g = load_golden()
    alg, env, norm = build(synthetic.make_weights(3))
    alg.priv_reg_counter = int(g["in.priv_reg_counter"])
    b = {k: g["in.batch." + k].to(DEV) for k in ("actions", "target_values", "advantages", "returns", "old_actions_log_prob",
                                                  "old_mu", "old_sigma")}
    obs = g["in.obs"].to(DEV)
    sample = (obs, obs, b["actions"], b["target_values"], b["advantages"], b["returns"], b["old_actions_log_prob"], b["old_mu"],
              b["old_sigma"], (None, None), None)
    out = alg.update_actor_critic(sample)
    torch.cuda.synchronize()
    for v, k in zip(out, ("surrogate_loss", "value_loss", "b_loss", "entropy", "priv_reg_loss", "estimator_loss")):
        assert_close(f"ppo.{k}", v.cpu(), g[f"ppo.{k}"], rtol=NET_RTOL, atol=NET_ATOL)
    assert abs(alg.lr_ac - float(g["ppo.lr_new"])) < 1e-9


def test_ppo_minibatch_step_at_full_size_matches_oracle():
    """BASELINE config 0 at its full size: one PPO minibatch step (24 576 rows gathered by K6 out of a recorded 4096 x 24
    RolloutStorage whose returns / advantages come from K5) against the CPU oracle's loss graph on the same rows."""
    import trainer as OT
    from helpers import assert_close
    from qa_b200 import synthetic
    from test_trainer_gpu import NET_ATOL, NET_RTOL, build
    torch.backends.cuda.matmul.allow_tf32 = False
    N, T = 4096, 24
    w = synthetic.make_weights(3)
    alg, env, norm = build(w, n_envs=N)
    st = alg.storage
    g = torch.Generator().manual_seed(2)
    st.observations.copy_(0.5 * torch.randn(T, N, 671, generator=g))
    st.privileged_observations.copy_(st.observations)
    with torch.no_grad():
        for t in range(T):
            o = st.observations[t]
            alg.act(o, o, normal_draw=torch.randn(N, 12, generator=g).to(DEV))
            tr = alg.transition
            st.actions[t], st.values[t] = tr.actions, tr.values
            st.actions_log_prob[t, :, 0], st.mu[t], st.sigma[t] = tr.actions_log_prob, tr.action_mean, tr.action_sigma
    # make the step non-trivial: the behaviour policy differs a little from the current one
    st.mu.add_(0.05 * torch.randn(T, N, 12, generator=g).to(DEV))
    st.sigma.mul_(1.05)
    st.actions_log_prob.add_(0.05 * torch.randn(T, N, 1, generator=g).to(DEV))
    st.rewards.copy_(0.05 * torch.rand(T, N, 1, generator=g))
    st.dones.copy_((torch.rand(T, N, 1, generator=g) < 0.02).byte())
    st.compute_returns(torch.zeros(N, 1, device=DEV), 0.99, 0.95)
    mb_size = T * N // 4
    idx = torch.randperm(T * N, generator=g).to(DEV)
    alg._alloc_minibatch(mb_size)
    alg._kl = torch.zeros((), device=DEV)
    alg._encode_history()
    alg._gather(idx[:mb_size])
    coef = OT.priv_reg_coef(1500)
    alg._priv_reg_coef.fill_(coef)
    alg._stats.zero_()
    mb = {k: v.clone() for k, v in alg._mb.items()}
    alg._minibatch_step()
    torch.cuda.synchronize()
    stats = alg._stats.cpu()
    batch = dict(obs=mb["obs"].cpu(), critic_obs=mb["critic_obs"].cpu(), actions=mb["actions"].cpu(), target_values=mb["values"].cpu(),
                 advantages=mb["advantages"].cpu(), returns=mb["returns"].cpu(), old_actions_log_prob=mb["old_actions_log_prob"].cpu(),
                 old_mu=mb["old_mu"].cpu(), old_sigma=mb["old_sigma"].cpu())
    # the gathered rows are the storage rows the permutation names (K6), with K5's normalised advantages
    rows = idx[:mb_size].cpu()
    assert torch.equal(batch["obs"], st.observations.flatten(0, 1).cpu()[rows])
    assert abs(float(st.advantages.mean())) < 1e-4 and abs(float(st.advantages.std()) - 1.0) < 1e-3
    with torch.no_grad():
        L = OT.ppo_losses(w["ac"], w["est"], batch, priv_reg_coef=coef)
    for i, k in enumerate(("surrogate_loss", "value_loss", "b_loss", "entropy", "priv_reg_loss", "estimator_loss")):
        assert_close(f"ppo.{k}", stats[i], L[k].float(), rtol=NET_RTOL, atol=NET_ATOL)
    assert_close("kl_mean", stats[6], L["kl_mean"].float(), rtol=1e-3, atol=1e-5)
    assert abs(alg.lr_ac - OT.adaptive_lr(1e-3, float(L["kl_mean"]))) < 1e-9


def test_runner_learn_books_episodes_logs_reference_tags_and_round_trips_checkpoint(tmp_path):
    """`OnPolicyRunner.learn` with a log_dir (on_policy_runner.py:118-233): device-staged episode bookkeeping (fused K19
    reward terms == torch path), the reference's scalar tags, and `save` / `load` with the six torch.optim-layout dicts."""
    import bench
    from qa_b200.pipeline import BbcIteration
    torch.backends.cuda.matmul.allow_tf32 = False
    N, T = 256, 4
    cfg, static, snaps, table = bench.build_workload(0, DEV, n_envs=N, steps=T)
    runs = []
    for fused in (False, True):
        it = BbcIteration(cfg, static, snaps, table, device=DEV, seed=77, use_cuda_graph=False)
        r = it.runner
        r.fused_rollout, r.log_dir, r.save_interval = fused, str(tmp_path / f"run{int(fused)}"), 1
        torch.manual_seed(5)
        r.learn(2)
        runs.append(r)
    a, b = runs[0].writer.scalars, runs[1].writer.scalars
    for tag in ("Loss/surrogate_loss", "Loss/value_loss", "Loss/estimator_loss", "Loss/hist_latent_loss", "Loss/mean_noise_std",
                "LR/lr_ac", "Perf/total_fps", "Perf/collection time", "Perf/learning_time", "Episode/rew_torques",
                "Episode/rew_tracking_lin_vel", "Train/mean_reward", "Train/mean_reward_i", "Train/mean_reward_us",
                "Train/mean_reward_ss", "Train/mean_reward_t", "Train/mean_episode_length"):
        assert tag in a and tag in b, tag
        assert [s for s, _ in b[tag]] == [0, 1][-len(b[tag]):] and all(np.isfinite(v) for _, v in b[tag]), tag
    for tag in ("Train/mean_reward", "Train/mean_reward_i", "Train/mean_reward_us", "Train/mean_reward_ss", "Train/mean_reward_t",
                "Train/mean_episode_length", "Episode/rew_torques"):            # first iteration: identical policy in both runs
        va, vb = a[tag][0][1], b[tag][0][1]
        assert abs(va - vb) <= 5e-3 * abs(va) + 1e-4, (tag, va, vb)
    assert len(runs[1].book.len_buffer) > 0
    r = runs[1]
    path = f"{r.log_dir}/model.pt"
    d = torch.load(path, map_location="cpu", weights_only=False)
    assert list(d) == ['actor_critic', 'estimator', 'disc', 'optim_ac', 'optim_hist_encoder', 'optim_estimator', 'optim_d',
                       'optim_q_eps', 'optim_q_c', 'disc_normalizer', 'reward_i_normalizer', 'iter', 'infos'] and d["iter"] == 2
    assert float(d["optim_ac"]["state"][0]["step"]) == 2 * 20 and float(d["optim_hist_encoder"]["state"][0]["step"]) == 20
    it2 = BbcIteration(cfg, static, snaps, table, device=DEV, seed=3, use_cuda_graph=False)
    it2.runner.load(path)
    assert it2.runner.current_learning_iteration == 2
    for (k, v), w in zip(r.alg.actor_critic.state_dict().items(), it2.runner.alg.actor_critic.state_dict().values()):
        assert torch.equal(v, w), k
    for o1, o2 in ((r.alg.optim_ac, it2.runner.alg.optim_ac), (r.alg.optim_estimator, it2.runner.alg.optim_estimator),
                   (r.alg.optim_hist_encoder, it2.runner.alg.optim_hist_encoder)):
        assert torch.equal(o1.exp_avg, o2.exp_avg) and torch.equal(o1.exp_avg_sq, o2.exp_avg_sq)
        assert int(o1.step_count) == int(o2.step_count) and float(o1.lr) == pytest.approx(float(o2.lr))


def test_batched_discriminator_step_matches_three_pass_step_on_device():
    """`QA_DISC_BATCHED=1` path (one shared trunk pass, DESIGN.md open item 3) against the default three-pass step, eager and as
    a CUDA graph: the 11 statistics of the first minibatch step agree; post-step parameters agree up to Adam's sign-like
    response to summation-order noise in the weight gradients."""
    import types
    from qa_b200 import synthetic
    from test_trainer_gpu import build
    res = []
    for batched, graph in ((False, False), (True, False), (True, True)):
        alg, env, norm = build(synthetic.make_weights(3), n_envs=64)
        alg.use_cuda_graph, alg.disc_batched = graph, batched
        env.task_obs_weight, env.prior_parameters = 0.9, torch.full((5,), 0.2, device=DEV)
        gen = torch.Generator().manual_seed(4)
        alg.disc_storage.insert(torch.randn(900, 98, generator=gen).to(DEV), torch.rand(900, 1, generator=gen).to(DEV),
                                torch.nn.functional.one_hot(torch.randint(0, 5, (900,), generator=gen), 5).float().to(DEV))
        expert = types.SimpleNamespace(preloaded_s_lb=torch.randn(500, 98, generator=gen).to(DEV),
                                       preloaded_label=torch.randint(0, 5, (500,), generator=gen).to(DEV),
                                       preloaded_s_ulb=torch.randn(700, 98, generator=gen).to(DEV))
        torch.manual_seed(11)
        stats = alg.update_disc(expert, num_updates=1)
        norm.sync_host()
        res.append((stats, torch.cat([v.reshape(-1) for v in alg.disc.state_dict().values()]).clone().cpu(), norm.mean.copy(),
                    max(alg.lr_disc, alg.lr_q)))
    for other in res[1:]:
        for a, b in zip(res[0][0], other[0]):
            assert abs(a - b) <= 1e-4 * abs(a) + 1e-6, (res[0][0], other[0])
        d = (other[1] - res[0][1]).abs()
        assert float(d.max()) <= 3 * 2.5 * res[0][3] and float((d > 2e-5).float().mean()) < 1e-2, float(d.max())
        assert np.allclose(res[0][2], other[2], rtol=1e-5, atol=1e-7)


def test_post_physics_matches_oracle_at_32768():
    """K2 at eight times the per-GPU size of BASELINE config 5 (32 768 envs in ONE launch: 4 096 CTAs, several waves, unlike the
    one-wave 4 096-env case) against the CPU oracle, with the push step; masks, indices and counters bit-exact."""
    import bbc_env as O
    from helpers import mocap_table
    from qa_b200 import synthetic
    from qa_b200.config import BbcEnvConfig
    from test_env_gpu import check_against, make_env
    cfg = BbcEnvConfig(num_envs=32768)
    static = synthetic.make_static(cfg, seed=4321)
    snap = synthetic.make_snapshot(cfg, seed=4321, step=0)
    draws = synthetic.make_rng_draws(cfg, seed=4321, step=0)
    table = mocap_table()
    draws["mocap_clip_idx"] = table.sample_clip(draws["rt_c_idx"], draws["mocap_clip_u"])
    threads = torch.get_num_threads()
    torch.set_num_threads(8)
    try:
        want = O.post_physics_step(cfg, static, snap, draws, table, 400)
    finally:
        torch.set_num_threads(threads)
    env = make_env(cfg, static, snap, draws, 399, bulk="tiled", table=table)
    env.post_physics_step()
    torch.cuda.synchronize()
    check_against(env, want, None, want["terminal_disc_states"], "n32768[399]")
    assert int(want["reset_buf"].sum()) > 200


def test_production_rng_mode_is_bit_exact_against_the_oracle_fed_with_the_same_philox_stream():
    """Production mode (in-kernel Philox4x32-10, no injected draws) is not only "in range and reproducible"
    (tests/test_env_gpu.py::test_philox_mode_properties): `oracle/philox.py` materialises the kernel's counter layout as dense
    parity draws, and the oracle fed with them must predict the production-mode kernel on every buffer, masks bit-exact --
    with and without the push step."""
    import bbc_env as O
    import philox as P
    from helpers import mocap_table
    from qa_b200 import synthetic
    from qa_b200.config import BbcEnvConfig
    from test_env_gpu import check_against, make_env
    cfg = BbcEnvConfig(num_envs=4096)
    static = synthetic.make_static(cfg, seed=9)
    snap = synthetic.make_snapshot(cfg, seed=9, step=0)
    table = mocap_table()
    lanes = torch.nonzero(static["noise_scale_vec"]).flatten().tolist()
    cdf = P.prior_cdf(static["prior_parameters"].tolist(), cfg.latent_c_temperature)
    for counter_before in (3, 399):
        step = counter_before + 1                                   # post_physics_step advances the counter before the launch
        d = P.k2_draws(cfg.num_envs, 671, lanes, cdf, table.mode_offset.numpy(), table.mode_clips.numpy(), table.mode_cdf.numpy(),
                       seed=7, step=step)                           # make_env builds the env with seed 7
        draws = {k: torch.from_numpy(v) for k, v in d.items()}
        want = O.post_physics_step(cfg, static, snap, draws, table, step)
        env = make_env(cfg, static, snap, None, counter_before, table=table)
        env.post_physics_step()
        torch.cuda.synchronize()
        check_against(env, want, None, want["terminal_disc_states"], f"philox[{counter_before}]")
        assert int(want["reset_buf"].sum()) > 20


def test_depth_update_production_rng_is_bit_exact_against_the_oracle():
    """K14 in production mode (in-kernel Philox) against `oracle/tsc_depth.update_depth_buffer` fed with the same stream
    (`oracle/philox.depth_draws`): two consecutive frames, fill and shift, bit-exact."""
    import philox as P
    import tsc_depth as OD
    from qa_b200.depth import DepthBuffer
    from test_tsc_depth import FAR, NEAR, NOISE, _synthetic
    N, seed = 64, 77
    images, ep, buf0, _, _, _ = _synthetic(N, seed=3)
    db = DepthBuffer(N, device=DEV, seed=seed)
    db.depth_buffer.copy_(buf0)
    db.set_batched_images(images.to(DEV))
    want = buf0
    for frame, ep_t in enumerate((ep, ep + 1)):
        d = {k: torch.from_numpy(v) for k, v in P.depth_draws(N, 58, 87, seed=seed, step=frame + 1).items()}   # step counts updates
        want = OD.update_depth_buffer(want, images, ep_t, NEAR, FAR, NOISE, d["noise_scale_u"], d["offset_u"], d["pixel_u"])
        db.update_depth_buffer(ep_t.to(DEV))
        assert torch.equal(db.depth_buffer.cpu(), want), frame


def test_tsc_env_production_rng_is_bit_exact_against_the_oracle():
    """K16 / K17 with the in-kernel Philox reset randomisation against the TSC oracle fed with the same stream."""
    import philox as P
    import tsc_env as OE
    from helpers import assert_close
    from qa_b200 import synthetic
    from test_tsc_env import KEYS, collect, run_kernels
    N, seed = 4096, 7
    st = synthetic.make_tsc_static(N, seed)
    sn = synthetic.make_tsc_snapshot(N, st, seed)
    env, ids, terminal, _ = run_kernels(N, seed, parity=False)
    # the launch was keyed by the step counter AFTER post_physics_step advanced it, and by the env's seed
    dr = {k: torch.from_numpy(v) for k, v in P.tsc_draws(N, seed=env.seed, step=env.common_step_counter).items()}
    cfg = OE.TscCfg(num_envs=N)
    want = OE.post_physics_post(cfg, st, OE.post_physics_pre(cfg, st, sn, dr), sn["rigid_body_state_post"])
    got = collect(env)
    for k in KEYS:
        assert_close(k, got[k], want[k])
    assert torch.equal(ids.cpu(), want["reset_env_ids"]) and len(ids) > 20
    assert_close("terminal", terminal, want["terminal_disc_states"])


def test_post_physics_on_a_step_without_any_reset():
    """A step on which no env resets, times out or resamples (the oracle is pinned against the reference on exactly this case,
    oracle/gen_golden.py): empty compaction, every buffer equal, latched `extras` quantities untouched."""
    import bbc_env as O
    from helpers import mocap_table
    from qa_b200 import synthetic
    from qa_b200.config import BbcEnvConfig
    from test_env_gpu import check_against, make_env
    for N in (64, 4096):
        cfg = BbcEnvConfig(num_envs=N)
        static = synthetic.make_static(cfg, seed=11)
        snap = synthetic.make_snapshot(cfg, seed=11, step=0, reset_frac=0.0, plant_frac=0.0)
        draws = synthetic.make_rng_draws(cfg, seed=11, step=0)
        table = mocap_table()
        draws["mocap_clip_idx"] = table.sample_clip(draws["rt_c_idx"], draws["mocap_clip_u"])
        want = O.post_physics_step(cfg, static, snap, draws, table, 6)
        assert int(want["reset_buf"].sum()) == 0
        env = make_env(cfg, static, snap, draws, 5, table=table)
        means0, latched0 = env._episode_rew_means.clone(), env._time_outs_latched.clone()
        env.post_physics_step()
        torch.cuda.synchronize()
        check_against(env, want, None, want["terminal_disc_states"], f"zero-reset[{N}]")
        assert int(env._reset_count.item()) == 0
        assert torch.equal(env._episode_rew_means, means0) and torch.equal(env._time_outs_latched, latched0)

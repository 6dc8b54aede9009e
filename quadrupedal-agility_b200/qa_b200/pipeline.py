"""One training iteration of the BBC loop over recorded state, as `bench.py` and `smoke()` drive it.

Mirrors the reference loop order (bbc/rsl_rl/runners/on_policy_runner.py:156-225): for each of the
T=24 steps `act -> env.step -> (disc reward) -> process_env_step`, then `compute_returns` and
`update`.  IsaacGym is replaced by `RecordedPhysics` over T synthetic snapshots.

Two entry points with identical device work:
  run_resident()  simulator state already in HBM (kernel-level throughput, bench `value`)
  run_host()      each step's simulator state arrives from PINNED HOST memory (what a host-side
                  physics backend would hand over) and the iteration's mean reward is read back --
                  the public-API end-to-end number (bench `e2e`).
"""
import time
from typing import Dict, List

import torch

from . import ops
from . import config as K
from .legged_robot import LeggedRobot, RecordedPhysics

SIM_KEYS = ("root_states", "dof_state", "rigid_body_state", "contact_forces")
CARRIED = ("last_actions", "last_torques_org", "last_dof_vel", "last_root_vel", "obs_history_buf",
           "episode_length_buf", "last_contacts", "commands", "latent_eps", "latent_c", "episode_sums",
           "feet_air_time", "obs_disc_buf", "action_history_buf")


class BbcIteration:
    dtype_name = "f32"
    workload_name = "bbc_go2_locomotion_4096x24: rollout (act, env step, disc reward, storage) + GAE + PPO update"
    stage_names = ["act (estimator+actor+critic)", "action_push", "pd_torques x4",
                   "post_physics_bbc (fused obs/reward/termination/reset)", "compact_resets", "predict_disc_reward",
                   "process_env_step", "gae", "PPO update: 5 epochs x 4 minibatches (gather, fwd/bwd graph, clip+Adam)"]

    def __init__(self, cfg, static, snaps: List[Dict[str, torch.Tensor]], table, device, seed=1234, world_size=1,
                 bulk_store=True, use_cuda_graph=True, tiled=True):
        from .config import bbc_train_cfg
        from .rsl_rl.runner import OnPolicyRunner
        self.cfg, self.device, self.T = cfg, torch.device(device), len(snaps)
        dev, N, T = self.device, cfg.num_envs, len(snaps)
        # pinned host copies of the simulator tensors (run_host) and device-resident snapshots (run_resident)
        self.host_snaps = [{k: s[k].contiguous().pin_memory() for k in SIM_KEYS} for s in snaps]
        self.dev_snaps = [{k: s[k].to(dev) for k in SIM_KEYS} for s in snaps]
        self.staging = {k: torch.empty_like(self.dev_snaps[0][k]) for k in SIM_KEYS}
        self.phys_resident = RecordedPhysics(self.dev_snaps)
        self.phys_staged = RecordedPhysics([self.staging])
        self.env = LeggedRobot(cfg, self.phys_resident, static, table, device=dev, seed=seed, bulk_store=bulk_store, tiled=tiled)
        self.env.global_counter = 1
        torch.manual_seed(seed)
        train_cfg = bbc_train_cfg()
        train_cfg["algorithm"]["use_cuda_graph"] = use_cuda_graph
        self.runner = OnPolicyRunner(self.env, train_cfg, log_dir=None, device=dev)
        self.env.load_state({k: v.to(dev) for k, v in snaps[0].items() if k in CARRIED})
        self.phys_resident.cursor = -1
        self.obs, self.critic_obs = self.env.get_observations(), self.env.get_privileged_observations()
        self.runner._disc_hist = torch.stack([self.env.get_disc_observations()] * 2, dim=1)
        self.result_host = torch.zeros(8).pin_memory()
        self.k2_traffic_bytes = None
        self.h2d_bytes_per_iteration = T * sum(self.staging[k].numel() * self.staging[k].element_size() for k in SIM_KEYS)
        self.d2h_bytes_per_iteration = 4 * 8
        self._launch0 = ops.launches
        self._iters = 0
        self._k2_pairs = []
        self._phase_events = []
        self.graph_rollout = use_cuda_graph
        self._rollout_graphs = {}

    # ---- bookkeeping ---------------------------------------------------------------------------------
    def reset_counters(self):
        self._launch0 = ops.launches
        self._iters = 0
        self.env.k2_events = None
        self._k2_pairs = []
        self._phase_events = []

    def phase_ms(self):
        """(collection, learning) device milliseconds per iteration -- the reference's Perf/collection time and
        Perf/learning_time split (on_policy_runner.py:208-228)."""
        n = max(len(self._phase_events), 1)
        return (sum(e[0].elapsed_time(e[1]) for e in self._phase_events) / n,
                sum(e[1].elapsed_time(e[2]) for e in self._phase_events) / n)

    @property
    def launch_count(self):
        """libqa_b200 kernels per iteration."""
        return (ops.launches - self._launch0) // max(self._iters, 1)

    def k2_time_ms(self):
        return sum(a.elapsed_time(b) for a, b in self._k2_pairs), len(self._k2_pairs)

    # ---- the iteration (on_policy_runner.py:152-225) -----------------------------------------------------
    def _learn(self):
        alg = self.runner.alg
        with torch.no_grad():
            alg.compute_returns(self.critic_obs)
        return alg.update()

    # ---- CUDA-graph rollouts ---------------------------------------------------------------------------------
    def _rollout_eager(self, host: bool):
        env, runner = self.env, self.runner
        env.physics = self.phys_staged if host else self.phys_resident
        with torch.no_grad():
            for t in range(self.T):
                if host:
                    for k in SIM_KEYS:                           # the physics backend's hand-over: pinned host -> HBM
                        self.staging[k].copy_(self.host_snaps[t][k], non_blocking=True)
                self.obs, self.critic_obs = runner.rollout_step(self.obs, self.critic_obs)

    def _capture_rollout(self, host: bool):
        """Captures the whole T-step rollout (every torch op and libqa_b200 launch of `rollout_step`, and in host
        mode the T x 4 pinned-host -> HBM copies) as ONE CUDA graph.  Python-side state that the steps advance
        (ping-pong parity, snapshot cursor, storage.step) is periodic in T; what is not periodic -- the step counter
        that drives Philox, the push schedule and the contact rings -- lives on the device (K2 step_state)."""
        env, runner, alg = self.env, self.runner, self.runner.alg
        assert self.T % 2 == 0
        env.use_device_step_counter(True)
        if alg._disc_stage is None:
            alg.stage_disc_inserts(True)
        hist0 = runner._disc_hist.clone()
        runner._disc_hist = hist0
        obs0, crit0 = self.obs, self.critic_obs
        alg.storage.clear()
        self.phys_resident.cursor = -1
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                               # warm-up on a side stream (allocator, cuBLAS handles)
            self._rollout_eager(host)
            alg.storage.clear()
            runner._disc_hist = hist0.copy_(runner._disc_hist)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        before = ops.launches
        g = torch.cuda.CUDAGraph()
        self.obs, self.critic_obs = obs0, crit0
        with torch.cuda.graph(g):
            self._rollout_eager(host)
            hist0.copy_(runner._disc_hist)
        runner._disc_hist = hist0
        assert self.obs.data_ptr() == obs0.data_ptr()
        alg.storage.clear()
        return g, ops.launches - before

    def _rollout(self, host: bool):
        if not self.graph_rollout:
            self._rollout_eager(host)
            return
        key = "host" if host else "resident"
        if key not in self._rollout_graphs:
            self._rollout_graphs[key] = self._capture_rollout(host)
        g, n = self._rollout_graphs[key]
        g.replay()
        ops._count(n)
        self.runner.alg.flush_disc_stage()

    def run_resident(self, profile_phases=False):
        if profile_phases:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        self._rollout(host=False)
        if profile_phases:
            ev[1].record()
        stats = self._learn()
        if profile_phases:
            ev[2].record()
            self._phase_events.append(ev)
        self._iters += 1
        return stats

    def time_k2_only(self, reps=20):
        """Roofline leg: the T fused post-physics launches of one rollout (T different state snapshots), alone in a
        CUDA graph, timed with CUDA events around each replay.  Returns (total ms, launches)."""
        env = self.env
        env.physics = self.phys_resident
        env.use_device_step_counter(True)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            env.post_physics_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.phys_resident.cursor = -1
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(self.T):
                env._k2_only_step()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(reps):
            flush.fill_(1)
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total, reps * self.T

    def time_disc_update(self, reps=3):
        """Device milliseconds of one full discriminator update (gail.py:284-300: 4 x epochs x minibatches steps) over the
        replay buffer this rollout filled and synthetic expert sets of the reference's preload size (200 000 x 98)."""
        import types
        alg, dev = self.runner.alg, self.device
        g = torch.Generator().manual_seed(99)
        n = 200000
        expert = types.SimpleNamespace(preloaded_s_lb=torch.randn(n, 98, generator=g).to(dev),
                                       preloaded_label=torch.randint(0, 5, (n,), generator=g).to(dev),
                                       preloaded_s_ulb=torch.randn(n, 98, generator=g).to(dev))
        if alg.disc_storage.num_samples == 0:
            return None
        alg.storage.step = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        alg.update_disc(expert)                                  # warm-up + graph capture
        e0.record()
        for _ in range(reps):
            alg.update_disc(expert)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    def run_host(self):
        runner = self.runner
        self._rollout(host=True)
        with torch.no_grad():
            mean_rew = runner.alg.storage.rewards.mean()
        stats = self._learn()                                    # reads the loss statistics back (one D2H)
        self.result_host[0:1].copy_(mean_rew.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._iters += 1
        return float(self.result_host[0]), stats


def cpu_oracle_iteration(O, OT, cfg, static, snaps, draws, table, weights, rollout_steps=None, minibatch_steps=20,
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

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
        # run_host: two staging sets, so that the host->HBM copies of step t+1 (copy stream) overlap the kernels of step t
        self.staging = [{k: torch.empty_like(self.dev_snaps[0][k]) for k in SIM_KEYS} for _ in range(2)]
        self.phys_resident = RecordedPhysics(self.dev_snaps)
        self.phys_staged = RecordedPhysics(self.staging)
        self._copy_stream = torch.cuda.Stream(device=dev)
        self.env = LeggedRobot(cfg, self.phys_resident, static, table, device=dev, seed=seed, bulk_store=bulk_store, tiled=tiled)
        self.env.global_counter = 1
        torch.manual_seed(seed)
        train_cfg = bbc_train_cfg()
        train_cfg["algorithm"]["use_cuda_graph"] = use_cuda_graph
        self.runner = OnPolicyRunner(self.env, train_cfg, log_dir=None, device=dev)
        self.runner.full_rollouts = True                   # every rollout here runs all T steps: deferred reward tails are legal
        self.env.load_state({k: v.to(dev) for k, v in snaps[0].items() if k in CARRIED})
        self.phys_resident.cursor = -1
        self.obs, self.critic_obs = self.env.get_observations(), self.env.get_privileged_observations()
        self.runner._disc_hist = torch.stack([self.env.get_disc_observations()] * 2, dim=1)
        self.result_host = torch.zeros(8).pin_memory()
        self.k2_traffic_bytes = None
        self.h2d_bytes_per_iteration = T * sum(self.staging[0][k].numel() * self.staging[0][k].element_size() for k in SIM_KEYS)
        self.d2h_bytes_per_iteration = 4 * 8
        self._launch0 = ops.launches
        self._iters = 0
        self._k2_pairs = []
        self._phase_events = []
        self.graph_rollout = use_cuda_graph
        self._rollout_graphs = {}

    def release_graphs(self):
        """Drop every captured CUDA graph (rollout, PPO minibatch steps, discriminator step): graphs that hold captured NCCL
        collectives must die before `destroy_process_group()`."""
        alg = self.runner.alg
        self._rollout_graphs = {}
        alg._graphs = None
        alg._disc_graph = None
        alg._disc_graph_key = None

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
            if not host:
                for t in range(self.T):
                    self.obs, self.critic_obs = runner.rollout_step(self.obs, self.critic_obs)
                runner.finish_rollout()
                return
            # the physics backend's hand-over: pinned host -> HBM, double buffered on a copy stream (a parallel branch of the
            # captured graph): step t's kernels wait for copy t, copy t+2 waits for step t's kernels (its staging set is free)
            main, cs = torch.cuda.current_stream(), self._copy_stream
            self.phys_staged.cursor = -1
            cs.wait_stream(main)
            freed = [None, None]
            for t in range(self.T):
                b = t % 2
                with torch.cuda.stream(cs):
                    if freed[b] is not None:
                        cs.wait_event(freed[b])
                    for k in SIM_KEYS:
                        self.staging[b][k].copy_(self.host_snaps[t][k], non_blocking=True)
                    landed = torch.cuda.Event()
                    landed.record(cs)
                main.wait_event(landed)
                self.obs, self.critic_obs = runner.rollout_step(self.obs, self.critic_obs)
                freed[b] = torch.cuda.Event()
                freed[b].record(main)
            runner.finish_rollout()
            main.wait_stream(cs)

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
        alg = self.runner.alg
        if self.env.task_obs_weight_decay and hasattr(alg, "_task_obs_weight_dev"):
            alg._task_obs_weight_dev(refresh=True)       # K18 reads the (decaying) weight from this device scalar: a replay sees today's
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

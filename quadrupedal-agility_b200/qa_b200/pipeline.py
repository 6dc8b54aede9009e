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

    def __init__(self, cfg, static, snaps: List[Dict[str, torch.Tensor]], table, device, seed=1234, world_size=1,
                 gamma=0.99, lam=0.95):
        self.cfg, self.device, self.T = cfg, torch.device(device), len(snaps)
        dev, N, T = self.device, cfg.num_envs, len(snaps)
        # pinned host copies of the simulator tensors (run_host) and device-resident snapshots (run_resident)
        self.host_snaps = [{k: s[k].contiguous().pin_memory() for k in SIM_KEYS} for s in snaps]
        self.dev_snaps = [{k: s[k].to(dev) for k in SIM_KEYS} for s in snaps]
        self.staging = {k: torch.empty_like(self.dev_snaps[0][k]) for k in SIM_KEYS}
        self.phys_resident = RecordedPhysics(self.dev_snaps)
        self.phys_staged = RecordedPhysics([self.staging])
        self.env = LeggedRobot(cfg, self.phys_resident, static, table, device=dev, seed=seed)
        self.env.load_state({k: v.to(dev) for k, v in snaps[0].items() if k in CARRIED})
        self.env.global_counter = 1
        g = torch.Generator().manual_seed(seed + 99)
        self.actions = [torch.randn(N, K.NUM_ACTIONS, generator=g).to(dev) for _ in range(T)]
        self.gamma, self.lam = gamma, lam
        f = dict(device=dev, dtype=torch.float32)
        self.rewards = torch.zeros(T, N, 1, **f)
        self.values = torch.zeros(T, N, 1, **f)
        self.dones = torch.zeros(T, N, 1, device=dev, dtype=torch.uint8)
        self.last_values = torch.zeros(N, 1, **f)
        self.returns = torch.zeros(T, N, 1, **f)
        self.advantages = torch.zeros(T, N, 1, **f)
        self.gae_ws = torch.zeros(8, device=dev, dtype=torch.float64)
        self.mean_reward_host = torch.zeros(1).pin_memory()
        self.k2_traffic_bytes = None
        self.h2d_bytes_per_iteration = T * sum(self.staging[k].numel() * self.staging[k].element_size() for k in SIM_KEYS)
        self.d2h_bytes_per_iteration = 4
        self._launch0 = ops.launches
        self._iters = 0

    workload_name = "bbc_go2_locomotion_4096x24_env_gae (trainer stages pending)"
    stage_names = ["action_push", "pd_torques x4", "post_physics_bbc (fused obs/reward/termination/reset)",
                   "compact_resets", "gae"]

    # ---- bookkeeping ---------------------------------------------------------------------------------
    def reset_counters(self):
        self._launch0 = ops.launches
        self._iters = 0
        self.env.k2_events = None
        self._k2_pairs = []

    @property
    def launch_count(self):
        """libqa_b200 kernels per iteration."""
        return (ops.launches - self._launch0) // max(self._iters, 1)

    def k2_time_ms(self):
        pairs = getattr(self, "_k2_pairs", [])
        return sum(a.elapsed_time(b) for a, b in pairs), len(pairs)

    # ---- the iteration -----------------------------------------------------------------------------------
    def _collect_step(self, t):
        env = self.env
        obs, priv, rew, reset, ids, count, term = env.step_device(self.actions[t])
        self.rewards[t, :, 0].copy_(rew)
        self.dones[t, :, 0].copy_(reset)

    def _learn(self):
        ops.gae(self.rewards, self.values, self.dones, self.last_values, self.returns, self.advantages, self.gae_ws,
                self.gamma, self.lam)

    def run_resident(self, profile_k2=False):
        env = self.env
        env.physics = self.phys_resident
        env.k2_events = [] if profile_k2 else None
        for t in range(self.T):
            self._collect_step(t)
        self._learn()
        if profile_k2:
            self._k2_pairs = getattr(self, "_k2_pairs", []) + env.k2_events
            env.k2_events = None
        self._iters += 1

    def run_host(self):
        env = self.env
        env.physics = self.phys_staged
        for t in range(self.T):
            for k in SIM_KEYS:                                   # the physics backend's hand-over: pinned host -> HBM
                self.staging[k].copy_(self.host_snaps[t][k], non_blocking=True)
            self._collect_step(t)
        self._learn()
        self.mean_reward_host.copy_(self.rewards.mean().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._iters += 1
        return float(self.mean_reward_host[0])


def cpu_oracle_iteration(O, OT, cfg, static, snaps, draws, table, gamma=0.99, lam=0.95):
    """The same iteration through the CPU oracle (test infrastructure; called only by bench.py's
    cpu_baseline / --impl reference legs).  Returns (seconds, stage description)."""
    T, N = len(snaps), cfg.num_envs
    g = torch.Generator().manual_seed(7)
    carried = {k: snaps[0][k].clone() for k in CARRIED}
    rewards, dones = torch.zeros(T, N, 1), torch.zeros(T, N, 1, dtype=torch.uint8)
    values, last_values = torch.zeros(T, N, 1), torch.zeros(N, 1)
    t0 = time.perf_counter()
    for t in range(T):
        actions = torch.randn(N, 12, generator=g)
        hist, act = O.action_push(cfg, carried["action_history_buf"], actions, delay=0)
        s = dict(snaps[t])
        s.update({k: carried[k] for k in CARRIED})
        s["action_history_buf"], s["actions"] = hist, act
        for _ in range(cfg.decimation):
            _, s["torques_org"] = O.compute_torques(cfg, {**static, "dof_state": s["dof_state"]}, act.clone())
        out = O.post_physics_step(cfg, static, s, draws[t], table, t + 1)
        rewards[t, :, 0] = out["rew_buf"]
        dones[t, :, 0] = out["reset_buf"]
        for k in CARRIED:
            carried[k] = out[k]
    OT.compute_returns(rewards, values, dones, last_values, gamma, lam)
    sec = time.perf_counter() - t0
    return sec, "action_push, compute_torques x4, post_physics_step, compute_returns"

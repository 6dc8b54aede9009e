#!/usr/bin/env python
"""bench.py -- env-steps/sec of the BBC go2_locomotion hot path on N B200s (BASELINE.json metric).

One "step" = one training iteration of the reference loop (on_policy_runner.py:156-225) over synthetic
recorded state: T=24 env steps of 4096 envs per GPU (policy act -> action push -> 4x PD torques -> fused
post-physics -> disc reward -> storage) followed by GAE and the PPO update, IsaacGym excluded
(SURVEY.md 8d).  value = T*N*n_gpus / seconds_per_iteration, exactly `Perf/total_fps`
(on_policy_runner.py:256) without the physics time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the reference algorithm's CPU path (the oracle port, all host threads) on a
bounded sample of the same workload; it is the only place besides `cpu_baseline` where `oracle/` runs.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

# The dense layers of the ActorCritic / estimator / discriminator run on the tensor cores: TF32 operands, fp32
# accumulate (10-bit mantissa inputs; the fp32 path stays the default for the parity tests).
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
T_STEPS = 24
ENVS_PER_GPU = 4096
# one name for the workload in both arms (the driver compares the lines' `config`); == pipeline.BbcIteration.workload_name
WORKLOAD_NAME = "bbc_go2_locomotion_4096x24: rollout (act, env step, disc reward, storage) + GAE + PPO update"
K2_BYTES_PER_ENV = 11158          # algorithmic bytes of the fused obs/reward kernel per env-step


def k2_traffic():
    """DRAM bytes per K2 launch from the committed `ncu --set full` capture (profiles/k2_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if not os.path.isfile(p):
        return None
    t = json.load(open(p))
    return t["dram_bytes_read"] + t["dram_bytes_write"] if t.get("n_envs") == ENVS_PER_GPU else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc, self.gpu = None, gpu_index
        self.path = f"/tmp/qa_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(rank, device, n_envs=ENVS_PER_GPU, steps=T_STEPS):
    """Synthetic recorded rollout for this rank: T state snapshots + static per-env constants (seed 1234+rank)."""
    from qa_b200 import synthetic
    from qa_b200.config import BbcEnvConfig
    from qa_b200.mocap import MocapTable
    cfg = BbcEnvConfig(num_envs=n_envs)
    seed = 1234 + rank
    static = synthetic.make_static(cfg, seed=seed)
    snaps = [synthetic.make_snapshot(cfg, seed=seed, step=t) for t in range(steps)]
    table = MocapTable.from_npz(os.path.join(ROOT, "tests", "golden", "mocap_lb_table.npz"))
    return cfg, static, snaps, table


SIM_KEYS = ("root_states", "dof_state", "rigid_body_state", "contact_forces")


def run_ours(args):
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    from qa_b200.pipeline import BbcIteration
    from qa_b200.rsl_rl import linear
    linear.set_mode("tc" if args.linear == "tc" else "fp32")
    cfg, static, snaps, table = build_workload(rank, dev)
    it = BbcIteration(cfg, static, snaps, table, device=dev, seed=1234 + rank, world_size=world,
                      bulk_store=bool(args.k2_bulk), tiled=(args.k2_bulk == 2))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ----------------------------------------------------------
    for _ in range(args.warmup):
        it.run_resident()
    it.reset_counters()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)                                                 # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        ev0.record()
        it.run_resident(profile_phases=True)
        ev1.record()
        ev1.synchronize()
        total_ms += ev0.elapsed_time(ev1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    collect_ms, learn_ms = it.phase_ms()
    launches = it.launch_count
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = T_STEPS * cfg.num_envs * world / (ms_per_step * 1e-3)

    # ---- end-to-end timing through the public API with HOST buffers ("e2e") ---------------------------
    for _ in range(max(1, args.warmup // 2)):
        it.run_host()
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ev0.record()
        it.run_host()
        ev1.record()
        ev1.synchronize()
        e2e_ms += ev0.elapsed_time(ev1)
    barrier()
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = T_STEPS * cfg.num_envs * world / (float(t.item()) / args.steps * 1e-3)

    disc_ms = None
    if world == 1:                                  # SURVEY 8f-1, reported beside the metric (not part of it): one full
        disc_ms = it.time_disc_update()             # discriminator update = 80 minibatch steps of 1228 x 3 samples
    k2_ms, k2_launches = it.time_k2_only()          # roofline leg: the 24 K2 launches of a rollout, alone in a graph
    if rank == 0:
        pk, pk_src = peaks()
        k2_avg_s = (k2_ms / max(k2_launches, 1)) * 1e-3
        achieved = K2_BYTES_PER_ENV * cfg.num_envs / k2_avg_s / 1e9 if k2_launches else None
        cpu = cpu_baseline_sample() if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "collection_ms": collect_ms, "learning_ms": learn_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 GEMM operands / f32 accumulate (wide ActorCritic layers on tcgen05); f32 everything else incl. the output heads", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "envs_per_gpu": cfg.num_envs, "steps_per_env": T_STEPS,
                       "stages": it.stage_names, "rng": "in-kernel Philox4x32-10",
                       "l2": "256 MiB flush between timed iterations; per-step working set 46 MB < 126 MB L2",
                       "parallelism": f"env-sharded dp{world}",
                       "collectives_per_optimiser_step": (1 if it.runner.alg._grad_arena is not None else 3) if world > 1 else 0,
                       "advantage_normalisation": "per rank (each rank normalises its own 98 304 samples, rollout_storage.py:111: "
                                                  "W ranks behave like W reference runs of 4096 envs whose gradients are averaged)",
                       "ppo_step": "static schedule (ppo_plan)" if it.runner.alg._plan is not None else "autograd"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": it.h2d_bytes_per_iteration,
                    "d2h_bytes_per_step": it.d2h_bytes_per_iteration},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_post_physics_bbc", "achieved": achieved,
                         "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s",
                         "frac": (achieved / pk["hbm_gbs"]) if achieved else None,
                         "us_per_launch": k2_avg_s * 1e6, "launches_timed": k2_launches,
                         "how": "CUDA events around graph replays of the rollout's 24 K2 launches (24 distinct state "
                                "snapshots), 256 MiB L2 flush before each replay",
                         "algorithmic_bytes_per_launch": K2_BYTES_PER_ENV * cfg.num_envs, "traffic": k2_traffic()},
            "clocks": clocks,
        }
        if disc_ms is not None:
            line["disc_update_ms"] = disc_ms
        if world == 1:
            line["tsc_env"] = time_tsc_env(dev)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1:
            line["roofline_gemm"] = gemm_roofline_sample(dev, pk)
            line["roofline_32768"] = k2_roofline_32768(dev, pk)
            if not args.no_tsc:
                for key, student in (("tsc_teacher", False), ("tsc_student", True)):
                    try:                                         # BASELINE configs 3 / 4 beside the headline (guarded like the baselines)
                        leg = tsc_leg(dev, student, steps=3, warmup=2)
                        n, T = leg["envs_per_gpu"], leg["steps_per_env"]
                        leg["value"] = T * n / (leg["ms_per_step"] * 1e-3)
                        leg["e2e_value"] = T * n / (leg["ms_per_step_host_fed"] * 1e-3)
                        leg["unit"] = UNIT
                        line[key] = leg
                    except Exception as e:                       # noqa: BLE001
                        line[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            if not args.no_fp32_value:
                line["value_fp32_linear"] = fp32_linear_value(args, rank, dev)
        if world == 1 and not args.no_torch_gpu_baseline:
            line["torch_gpu_baseline"] = torch_gpu_baseline_sample(dev)
        print(json.dumps(line, default=str), flush=True)
    if world > 1:
        # Tear down in dependency order: the CUDA graphs that captured NCCL collectives first, then the process group (round 1
        # destroyed the group while those graphs were alive and blocked for minutes).  A watchdog keeps a hang here from
        # costing the run: the JSON line is already out.
        import gc
        import threading
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        it.release_graphs()
        del it
        gc.collect()
        torch.cuda.synchronize()
        wd = threading.Timer(30.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.destroy_process_group()
        wd.cancel()


TSC_BYTES_PER_ENV = 12634      # algorithmic bytes of the TSC post-physics pair (K16 + K17) per env-step


def time_tsc_env(dev, n_envs=ENVS_PER_GPU, steps=T_STEPS, reps=10):
    """Config 3 of BASELINE.json (TSC teacher, 4096 envs): the two post-physics kernels K16 + K17 of `steps` env steps
    (distinct synthetic snapshots) replayed as one CUDA graph; returns a dict reported beside the metric."""
    from qa_b200 import ops, synthetic
    from qa_b200.legged_robot_tsc import LeggedRobotTSC, RecordedPhysicsTSC, TscEnvConfig
    st = synthetic.make_tsc_static(n_envs, seed=1234)
    snaps = [synthetic.make_tsc_snapshot(n_envs, st, seed=1234, step=t) for t in range(4)]
    dev_snaps = [{k: v.to(dev).contiguous() for k, v in s.items() if isinstance(v, torch.Tensor)} for s in snaps]
    env = LeggedRobotTSC(TscEnvConfig(num_envs=n_envs), RecordedPhysicsTSC(dev_snaps), st, device=dev, seed=1234)
    env.load_state(snaps[0])
    env.action_hl_history_buf = dev_snaps[0]["action_hl_history_buf"]

    def body():
        for _ in range(steps):
            env.physics.refresh()
            env.common_step_counter += 1
            a = env._args()
            ops.post_physics_tsc(env._const, a, "pre")
            ops.post_physics_tsc(env._const, a, "post")

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    us = total / (reps * steps) * 1e3
    pk, _ = peaks()
    achieved = TSC_BYTES_PER_ENV * n_envs / (us * 1e-6) / 1e9
    return {"workload": f"tsc_go2_agility_teacher_{n_envs}: qa_post_physics_tsc_pre + _post per env step", "us_per_step": us,
            "env_steps_per_sec_env_only": n_envs / (us * 1e-6), "achieved_gbs": achieved, "frac_of_hbm_peak": achieved / pk["hbm_gbs"],
            "algorithmic_bytes_per_env_step": TSC_BYTES_PER_ENV}


TSC_KEYS = ("root_states", "dof_state", "contact_forces", "rigid_body_state", "obst_dof_state", "rigid_body_state_post")
K14_BYTES_PER_ENV = 25440 + 20184 + 40368      # camera image read + depth history read + history written, per env


def _cuda_time_ms(fn, steps, flush):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


class TscWorkload:
    """BASELINE configs 3 and 4 over synthetic recorded state: the TSC env (K16 / K17 + K0 / K1, K14 for the student) under
    `OnPolicyRunnerTSC` with the frozen BBC controller.  `host=True` feeds every step's simulator tensors (and camera images)
    from pinned host memory inside the timed region."""

    def __init__(self, dev, n_envs, T, seed, student):
        from qa_b200 import synthetic
        from qa_b200.config import tsc_train_cfg
        from qa_b200.legged_robot_tsc import LeggedRobotTSC, RecordedPhysicsTSC, TscEnvConfig
        from qa_b200.rsl_rl.tsc_runner import OnPolicyRunnerTSC
        self.dev, self.N, self.T, self.student = dev, n_envs, T, student
        st = synthetic.make_tsc_static(n_envs, seed)
        n_snaps = 4
        snaps = [synthetic.make_tsc_snapshot(n_envs, st, seed, step=t) for t in range(n_snaps)]
        self.host = [{k: v.contiguous().pin_memory() for k, v in s.items() if isinstance(v, torch.Tensor) and k in TSC_KEYS} for s in snaps]
        self.dsn = [{k: v.to(dev).contiguous() for k, v in s.items() if isinstance(v, torch.Tensor)} for s in snaps]
        self.env = env = LeggedRobotTSC(TscEnvConfig(num_envs=n_envs), RecordedPhysicsTSC(self.dsn), st, device=dev, seed=seed)
        env.load_state(snaps[0])
        self.images_host = None
        if student:
            from qa_b200.depth import DepthBuffer
            depth = DepthBuffer(n_envs, device=dev, seed=seed + 5)
            g = torch.Generator().manual_seed(seed)
            self.images_host = [(-(0.1 + 6.0 * torch.rand(n_envs, 60, 106, generator=g))).pin_memory() for _ in range(2)]
            self.images = self.images_host[0].to(dev)
            depth.set_batched_images(self.images)
            env.attach_depth(depth)
        env.post_physics_step()
        cfg = tsc_train_cfg(use_camera=student)
        if student:
            cfg["depth_encoder"]["num_steps_per_env"] = T
        else:
            cfg["runner"]["num_steps_per_env"] = T
        torch.manual_seed(seed)
        self.r = r = OnPolicyRunnerTSC(env, cfg, device=dev)
        self.obs, self.obs_bbc = env.get_observations(), env.get_observations_bbc().clone()
        r._disc_hist = torch.stack([env.get_observations_disc()] * 2, dim=1)
        self.critic, self.infos = self.obs, {}
        self.h2d_bytes = T * (sum(v.numel() * v.element_size() for v in self.host[0].values()) +
                              (self.images_host[0].numel() * 4 if student else 0))
        self.result_host = torch.zeros(8).pin_memory()
        if student:
            self.hist = torch.zeros(n_envs, env.cfg.action_buf_len, r.num_actions, device=dev)
            self.sinfos = {"depth": env.depth_buffer.clone()[:, -1], "delta_yaw_ok": torch.ones(n_envs, dtype=torch.bool, device=dev)}
            r.alg.depth_encoder.train()
            r.alg.depth_actor.train()
            self.keys = ("depth", "depth_latent", "scandots_latent", "actions_teacher", "actions_student", "yaw_student", "yaw_teacher",
                         "obst_student", "obst_teacher", "delta_yaw_ok")

    def _feed(self, t):
        i = t % len(self.host)
        for k, v in self.host[i].items():
            self.dsn[i][k].copy_(v, non_blocking=True)
        if self.student:
            self.images.copy_(self.images_host[t % 2], non_blocking=True)

    def iteration(self, host=False):
        r, env = self.r, self.env
        if not self.student:
            with torch.no_grad():
                for t in range(self.T):
                    if host:
                        self._feed(t)
                    self.obs, self.obs_bbc, self.critic, self.infos = r.rollout_step(self.obs, self.obs_bbc, self.critic, self.infos)
                r.alg.compute_returns(self.critic)
            out = r.alg.update()
            if host:
                self.result_host[0] = float(out[0])
            return out
        buf = {k: [] for k in self.keys}
        for t in range(self.T):
            if host:
                self._feed(t)
            self.obs, self.obs_bbc, self.sinfos, self.hist, _rew, _dones = r.student_step(self.obs, self.obs_bbc, self.sinfos, self.hist,
                                                                                          buf, False)
        cat = {k: torch.cat(v, dim=0) for k, v in buf.items() if k != "delta_yaw_ok"}
        out = r.alg.update_depth_actor(cat["actions_student"], cat["actions_teacher"], cat["yaw_student"], cat["yaw_teacher"],
                                       cat["obst_student"], cat["obst_teacher"], cat["depth"])
        r.alg.depth_encoder.detach_hidden_states()
        r._student_last = tuple(t.detach() for t in r._student_last)
        return out


def tsc_leg(dev, student, world=1, rank=0, n_envs=None, T=T_STEPS, steps=5, warmup=3):
    """One TSC workload (config 3: teacher iteration = rollout + GAE + PPO.update, tsc on_policy_runner.py:201-228 + ppo.py:160-282;
    config 4: student iteration = depth rollout + update_depth_actor, :278-441): device-resident value, host-fed e2e, and the
    roofline of the workload's own fused kernel(s) -- K16 + K17 (teacher) / K14 (student).  Returns a dict of rank-local numbers."""
    n_envs = n_envs or (2048 if student else ENVS_PER_GPU)
    w = TscWorkload(dev, n_envs, T, 1234 + rank, student)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        w.iteration()
    from qa_b200 import ops
    l0 = ops.launches
    ms = _cuda_time_ms(lambda: w.iteration(), steps, flush)
    launches = (ops.launches - l0) // max(steps, 1)
    w.iteration(host=True)
    ms_host = _cuda_time_ms(lambda: w.iteration(host=True), steps, flush)
    out = {"envs_per_gpu": n_envs, "steps_per_env": T, "ms_per_step": ms, "ms_per_step_host_fed": ms_host,
           "h2d_bytes_per_step": w.h2d_bytes, "d2h_bytes_per_step": 32, "gpu_launches": launches}
    if student:
        # K14 alone: T launches in a graph
        depth, env = w.env.depth, w.env
        g = torch.cuda.CUDAGraph()
        depth.update_depth_buffer(env.episode_length_buf)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(T):
                depth.update_depth_buffer(env.episode_length_buf)
        us = _cuda_time_ms(g.replay, 10, flush) / T * 1e3
        out["kernel"] = {"name": "k_depth_update (K14)", "us_per_launch": us, "bytes_per_env": K14_BYTES_PER_ENV}
        out["kernel"]["achieved_gbs"] = K14_BYTES_PER_ENV * n_envs / (us * 1e-6) / 1e9
    return out


def cpu_reference_sample(n_envs, rollout_steps, minibatch_steps, threads):
    """The reference algorithm's CPU path (oracle port): `rollout_steps` of the 24 env steps, GAE, and
    `minibatch_steps` of the 20 PPO minibatch steps.  Returns (extrapolated seconds per full iteration, detail)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import iteration as oracle_iteration
    from qa_b200 import synthetic
    torch.set_num_threads(threads)
    cfg, static, snaps, table = build_workload(0, "cpu", n_envs=n_envs, steps=T_STEPS)
    draws = [synthetic.make_rng_draws(cfg, seed=1234, step=t) for t in range(rollout_steps)]
    for d in draws:
        d["mocap_clip_idx"] = table.sample_clip(d["rt_c_idx"], d["mocap_clip_u"])
    r = oracle_iteration.oracle_iteration(cfg, static, snaps, draws, table, synthetic.make_weights(1),
                                          rollout_steps=rollout_steps, minibatch_steps=minibatch_steps)
    full = r["t_rollout"] * (T_STEPS / rollout_steps) + r["t_gae"] + r["t_update"] * (20 / minibatch_steps)
    return full, r


def _cpu_line(threads, rollout_steps, minibatch_steps, full, r):
    return {"value": T_STEPS * ENVS_PER_GPU / full, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"oracle/ port of the reference on torch {torch.__version__} CPU, {threads} threads, {ENVS_PER_GPU} envs: "
                       f"{rollout_steps}/24 rollout steps ({r['t_rollout']:.2f} s), GAE ({r['t_gae']:.3f} s), "
                       f"{minibatch_steps}/20 PPO minibatch steps of 24576 ({r['t_update']:.2f} s), "
                       + (f"one full iteration measured = {full:.2f} s" if rollout_steps == T_STEPS and minibatch_steps == 20 else
                          f"extrapolated to one full iteration = {full:.2f} s"))}


def gemm_roofline_sample(dev, pk, reps=20):
    """The kernel that dominates the step by TIME (the tcgen05 GEMM, ~50 % of an iteration; K2 is ~1 %): the critic's first
    layer at minibatch size (24576 x 512 x 671) forward with the fused bias + ELU epilogue, and its dX / dW contractions, each
    timed alone with CUDA events after a 256 MiB L2 flush.  Tensor-core roofline: TF32 operands run at half the bf16 rate, so
    the peak is MEASURED_PEAKS.json's bf16 figure / 2.  Guarded like the baseline legs."""
    try:
        from qa_b200 import ops
        M, N, K = 24576, 512, 671
        kp = (K + 3) // 4 * 4
        g = torch.Generator().manual_seed(0)
        x = torch.randn(M, kp, generator=g).to(dev)[:, :K]
        w = (torch.randn(N, kp, generator=g) / K ** 0.5).to(dev)[:, :K]
        b = torch.randn(N, generator=g).to(dev)
        y = torch.empty(M, N, device=dev)
        gz = torch.randn(M, N, generator=g).to(dev)
        dx = torch.empty(M, kp, device=dev)[:, :K]
        dw = torch.zeros(N, kp, device=dev)[:, :K]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn):
            for _ in range(3):
                fn()
            tot = 0.0
            for _ in range(reps):
                flush.fill_(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / reps * 1e3                                  # us per launch

        t_fwd = timed(lambda: ops.linear_fwd(x, w, b, y, "elu"))
        t_dx = timed(lambda: ops.linear_bwd(gz, None, w, dx=dx))
        t_dw = timed(lambda: ops.linear_bwd(gz, x, None, dw=dw))
        flops = 2.0 * M * N * K
        peak = pk["bf16_tflops"] / 2.0
        ach = flops / t_fwd / 1e6
        return {"bound": "tensor", "kernel": "k_gemm_tf32 (tcgen05, TF32 operands / fp32 accumulate)", "shape": [M, N, K],
                "achieved": ach, "peak": peak, "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 runs at half the bf16 rate)",
                "unit": "TFLOP/s", "frac": ach / peak, "us_per_launch": t_fwd, "dx_us_per_launch": t_dx, "dw_us_per_launch": t_dw,
                "dx_tflops": flops / t_dx / 1e6, "dw_tflops": flops / t_dw / 1e6, "flops_per_launch": flops,
                "how": f"CUDA events around single launches, 256 MiB L2 flush before each, mean of {reps}"}
    except Exception as e:                                               # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def k2_roofline_32768(dev, pk, n_envs=32768, n_snaps=4, launches=8, reps=10):
    """"the 32 768-env config is the honest bandwidth test" -- the fused post-physics kernel at 32 768 envs on ONE
    GPU (4096 CTAs, several waves: loads, compute and stores of different tiles overlap, unlike the one-wave 4096-env launch).
    `launches` K2 launches over `n_snaps` distinct state snapshots in a CUDA graph, 256 MiB L2 flush before each replay; the
    per-launch working set (366 MB) is larger than L2 by itself.  Guarded like the baseline legs."""
    try:
        from qa_b200 import synthetic
        from qa_b200.config import BbcEnvConfig
        from qa_b200.legged_robot import LeggedRobot, RecordedPhysics
        from qa_b200.mocap import MocapTable
        cfg = BbcEnvConfig(num_envs=n_envs)
        static = synthetic.make_static(cfg, seed=4321)
        snaps = [synthetic.make_snapshot(cfg, seed=4321, step=t) for t in range(n_snaps)]
        table = MocapTable.from_npz(os.path.join(ROOT, "tests", "golden", "mocap_lb_table.npz"))
        dev_snaps = [{k: s[k].to(dev) for k in SIM_KEYS} for s in snaps]
        env = LeggedRobot(cfg, RecordedPhysics(dev_snaps), static, table, device=dev, seed=4321)
        env.load_state({k: v.to(dev) for k, v in snaps[0].items() if k not in SIM_KEYS})
        env.use_device_step_counter(True)
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            env._k2_only_step()
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(launches):
                env._k2_only_step()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(reps):
            flush.fill_(1)
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        us = total / (reps * launches) * 1e3
        ach = K2_BYTES_PER_ENV * n_envs / (us * 1e-6) / 1e9
        return {"bound": "hbm", "kernel": "k_post_physics_bbc", "envs": n_envs, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "us_per_launch": us, "launches_timed": reps * launches,
                "algorithmic_bytes_per_launch": K2_BYTES_PER_ENV * n_envs,
                "how": f"CUDA events around graph replays of {launches} K2 launches ({n_snaps} distinct snapshots), 256 MiB L2 flush before each"}
    except Exception as e:                                               # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def fp32_linear_value(args, rank, dev, steps=3):
    """The same iteration with every dense layer in full fp32 (QA_LINEAR_MODE=fp32: cuBLAS sgemm through autograd, the parity-test
    mode) -- printed beside `value` because the headline runs the ActorCritic GEMMs with TF32 operands (north_star puts them on
    the tensor cores; there is no fp32 tensor-core mode).  Device-resident timing like `value`."""
    try:
        from qa_b200.pipeline import BbcIteration
        from qa_b200.rsl_rl import linear
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        prev = linear.get_mode()
        linear.set_mode("fp32")
        try:
            cfg, static, snaps, table = build_workload(rank, dev)
            it = BbcIteration(cfg, static, snaps, table, device=dev, seed=1234 + rank)
            for _ in range(3):
                it.run_resident()
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ms = 0.0
            for _ in range(steps):
                flush.fill_(1)
                torch.cuda.synchronize()
                e0.record()
                it.run_resident()
                e1.record()
                e1.synchronize()
                ms += e0.elapsed_time(e1)
            it.release_graphs()
        finally:
            linear.set_mode(prev)
            torch.backends.cuda.matmul.allow_tf32 = tf32
        ms /= steps
        return {"value": T_STEPS * cfg.num_envs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
                "dtype": "f32 everywhere (cuBLAS sgemm, allow_tf32=False)"}
    except Exception as e:                                               # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def torch_gpu_baseline_sample(dev, n_envs=ENVS_PER_GPU):
    """the reference's own single-GPU path -- its PyTorch op sequence (the oracle port, eager torch kernels, fp32)
    on the SAME B200 and workload, one full iteration (24 env steps + GAE + 20 PPO minibatch steps) after a short warm-up.
    This is the number north_star's ">= 4x the reference's single-GPU env-steps/sec" refers to; a reported baseline, guarded so
    that a failure here can never cost the benchmark line."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import iteration as oracle_iteration
        from qa_b200 import synthetic
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False                  # the reference runs fp32 matmuls (torch default)
        mv = lambda d: {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}     # noqa: E731
        cfg, static, snaps, table = build_workload(0, "cpu", n_envs=n_envs)
        draws = []
        for t in range(T_STEPS):
            d = synthetic.make_rng_draws(cfg, seed=1234, step=t)
            d["mocap_clip_idx"] = table.sample_clip(d["rt_c_idx"], d["mocap_clip_u"])
            draws.append(mv(d))
        static, snaps, table = mv(static), [mv(s) for s in snaps], table.to(dev)
        w = synthetic.make_weights(1)
        w = {k: (mv(v) if isinstance(v, dict) else v.to(dev)) for k, v in w.items()}
        run = lambda rs, ms: oracle_iteration.oracle_iteration(cfg, static, snaps, draws, table, w, rollout_steps=rs,   # noqa: E731
                                                               minibatch_steps=ms, device=dev)
        run(2, 2)                                                        # warm-up: allocator, cuBLAS handles, autotuning
        r = run(T_STEPS, 20)
        torch.backends.cuda.matmul.allow_tf32 = tf32
        full = r["t_rollout"] + r["t_gae"] + r["t_update"]
        return {"value": T_STEPS * n_envs / full, "unit": UNIT, "kind": "port",
                "ms_per_step": full * 1e3, "collection_ms": r["t_rollout"] * 1e3, "learning_ms": (r["t_gae"] + r["t_update"]) * 1e3,
                "sample": f"oracle/ port of the reference's PyTorch path, eager fp32 torch {torch.__version__} kernels on the same GPU, "
                          f"{n_envs} envs, one full iteration (24 env steps, GAE, 20 PPO minibatch steps of {6 * n_envs}), wall clock "
                          f"with device synchronisation"}
    except Exception as e:                                               # noqa: BLE001  (a baseline must not break the line)
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def run_tsc(args):
    """`--workload tsc_teacher | tsc_student`: the same JSON contract for BASELINE configs 3 / 4 (weak scaling over env shards;
    the PPO / distillation gradients are all-reduced per optimiser step as in the BBC path)."""
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    from qa_b200.rsl_rl import linear
    linear.set_mode("tc" if args.linear == "tc" else "fp32")
    student = args.workload == "tsc_student"
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    leg = tsc_leg(dev, student, world, rank, steps=args.steps, warmup=args.warmup)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([leg["ms_per_step"], leg["ms_per_step_host_fed"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_host = (float(v) for v in t.tolist())
    n, T = leg["envs_per_gpu"], leg["steps_per_env"]
    if rank == 0:
        name = ("tsc_go2_agility_student (--use_camera, depth path)" if student else "tsc_go2_agility_teacher (--exptid base)")
        line = {"metric": METRIC, "value": T * n * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32 GEMM operands / f32 accumulate; conv / GRU of the depth student on cuDNN (tf32)", "data": "synthetic",
                "config": {"workload": f"{name}: {n} envs per GPU x {T} steps, rollout + update", "envs_per_gpu": n, "steps_per_env": T,
                           "l2": "256 MiB flush between timed iterations", "parallelism": f"env-sharded dp{world}"},
                "e2e": {"value": T * n * world / (ms_host * 1e-3), "unit": UNIT, "h2d_bytes_per_step": leg["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": leg["d2h_bytes_per_step"]},
                "gpu_launches": leg["gpu_launches"], "clocks": clocks}
        if "kernel" in leg:
            pk, _ = peaks()
            k = leg["kernel"]
            line["roofline"] = {"bound": "hbm", "kernel": k["name"], "achieved": k["achieved_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                                "frac": k["achieved_gbs"] / pk["hbm_gbs"], "us_per_launch": k["us_per_launch"], "traffic": None}
        elif world == 1:
            te = time_tsc_env(dev)
            pk, _ = peaks()
            line["roofline"] = {"bound": "hbm", "kernel": "k_post_physics_tsc_pre + _post (K16 + K17)", "achieved": te["achieved_gbs"],
                                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": te["frac_of_hbm_peak"], "us_per_launch": te["us_per_step"],
                                "traffic": None}
        print(json.dumps(line, default=str), flush=True)
    if world > 1:
        import threading
        dist.barrier()
        torch.cuda.synchronize()
        wd = threading.Timer(30.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.destroy_process_group()
        wd.cancel()


def cpu_baseline_sample():
    threads = os.cpu_count() or 1
    rs, ms = 12, 6                                   # ~10-20 s of CPU work
    full, r = cpu_reference_sample(ENVS_PER_GPU, rs, ms, threads)
    return _cpu_line(threads, rs, ms, full, r)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # the WHOLE iteration (24/24 env steps, GAE, 20/20 PPO minibatch steps of 24 576): measured, not extrapolated -- about 3 s per
    # step on 16 host cores, so the driver's --steps / --warmup stay within minutes
    rs, ms = T_STEPS, 20
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(ENVS_PER_GPU, 2, 1, threads)
    fulls = []
    for _ in range(args.steps):
        full, r = cpu_reference_sample(ENVS_PER_GPU, rs, ms, threads)
        fulls.append(full)
    full = sum(fulls) / len(fulls)
    value = T_STEPS * ENVS_PER_GPU / full
    world = int(os.environ.get("WORLD_SIZE", 1))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "envs_per_gpu": ENVS_PER_GPU, "steps_per_env": T_STEPS,
                       "sample": "host cores, one FULL iteration per step (24 env steps + GAE + 20 PPO minibatch steps), "
                                 "see cpu_baseline.sample"},
            "cpu_baseline": _cpu_line(threads, rs, ms, full, r),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line, default=str), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--linear", default="tc", choices=["tc", "cublas"],
                    help="dense layers: tc = hand-written tcgen05 TF32 forward (default), cublas = library TF32 GEMMs")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true",
                    help="skip the reference's PyTorch path timed on the same GPU (reported beside the metric)")
    ap.add_argument("--no-fp32-value", action="store_true", help="skip the full-fp32 (cuBLAS) run of the same iteration")
    ap.add_argument("--workload", default="bbc", choices=["bbc", "tsc_teacher", "tsc_student"],
                    help="bbc = BASELINE configs 1 / 5 (default, the headline metric); tsc_teacher = config 3 (4096 envs per GPU); "
                         "tsc_student = config 4 (depth path, 2048 envs per GPU: 4096 envs on 2 GPUs)")
    ap.add_argument("--no-tsc", action="store_true", help="skip the TSC teacher / student sub-legs of the default line")
    ap.add_argument("--k2-bulk", type=int, default=2,
                    help="K2 variant: 2 = 8-env TMA tiles (default), 1 = warp-per-env + TMA row stores, 0 = warp stores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "bbc":
        run_tsc(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

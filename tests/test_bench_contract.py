"""bench.py contract pieces that run without a GPU: both arms name the same workload, and the reference arm's CPU leg (the
oracle port on host cores, a bounded sample) produces the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_both_arms_name_the_same_workload():
    import bench
    from qa_b200.pipeline import BbcIteration
    assert bench.WORKLOAD_NAME == BbcIteration.workload_name
    assert bench.METRIC == "env_steps_per_sec" and bench.UNIT == "env-steps/s"
    assert bench.K2_BYTES_PER_ENV == 11158 and bench.ENVS_PER_GPU == 4096 and bench.T_STEPS == 24


def test_cpu_sample_of_the_reference_port_runs_and_reports_its_sample():
    import torch
    import bench
    threads = torch.get_num_threads()
    try:
        full, r = bench.cpu_reference_sample(256, 2, 1, 2)
    finally:
        torch.set_num_threads(threads)
    assert full > 0 and all(k in r for k in ("t_rollout", "t_gae", "t_update"))
    line = bench._cpu_line(2, 2, 1, full, r)
    assert line["kind"] == "port" and line["cores"] == 2 and line["unit"] == "env-steps/s" and "2/24 rollout steps" in line["sample"]


def test_reference_arm_on_a_non_zero_rank_prints_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and not any(l.startswith("{") for l in out.stdout.splitlines()), out.stdout + out.stderr


def test_torch_baseline_leg_runs_the_full_iteration_and_never_raises():
    """The leg that times the reference's PyTorch path on the bench device (here: the host, 256 envs) -- same code, `device`
    is the only difference on the GPU box; a failure inside must come back as an `error` entry, not as an exception."""
    import torch
    import bench
    threads = torch.get_num_threads()
    try:
        out = bench.torch_gpu_baseline_sample(torch.device("cpu"), n_envs=256)
        assert "error" not in out, out
        assert out["value"] > 0 and out["kind"] == "port" and abs(out["ms_per_step"] - out["collection_ms"] - out["learning_ms"]) < 1e-6
        bad = bench.torch_gpu_baseline_sample("no-such-device", n_envs=256)
        assert set(bad) == {"error"}
    finally:
        torch.set_num_threads(threads)


def test_gemm_roofline_leg_never_raises():
    """Without a GPU the tcgen05 launch is refused; the leg must hand that back as an `error` entry (it is reported beside the
    K2 roofline: the tcgen05 GEMM is the kernel that dominates the step by time)."""
    import torch
    import bench
    out = bench.gemm_roofline_sample(torch.device("cpu"), {"bf16_tflops": 1600.0}, reps=1)
    assert set(out) == {"error"} and out["error"].startswith("RuntimeError")

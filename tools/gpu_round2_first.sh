#!/bin/bash
# (a 2-GPU follow-up: QA_SINGLE_ALLREDUCE=1 torchrun ... bench.py --gpus 2 against the default three collectives)
# First GPU call of round 2 (~3 min of box time): everything that was written at the end of round 1 without a GPU.
#   1. the whole GPU suite WITHOUT -x (tests/test_zz_runner_gpu.py holds the tests that have not run on a device yet, marked xfail(strict=False): look for XPASS)
#   2. the bench line (now with torch_gpu_baseline = the reference's PyTorch path on the same GPU) and the reference arm
#   3. the discriminator update with the shared forward pass (QA_DISC_BATCHED=1): compare disc_update_ms with run 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^XPASS|^XFAIL|passed|failed" gpurun_out/pytest_gpu.log | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
QA_DISC_BATCHED=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_disc_batched.json 2> gpurun_out/bench_disc_batched.err
echo "bench(disc batched) rc=$?"; python - <<'PY'
import json
for f in ("gpurun_out/bench.json", "gpurun_out/bench_disc_batched.json"):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "disc_update_ms", d.get("disc_update_ms"), "value", d.get("value"), "torch_gpu_baseline", d.get("torch_gpu_baseline"), "roofline_gemm", d.get("roofline_gemm"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
QA_SHARE_PRIV_LATENT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_share_priv.json 2> gpurun_out/bench_share_priv.err
echo "bench(shared priv latent) rc=$?"; python - <<'PY'
import json
for f in ("gpurun_out/bench.json", "gpurun_out/bench_share_priv.json"):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "value", d.get("value"), "learning_ms", d.get("learning_ms"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json

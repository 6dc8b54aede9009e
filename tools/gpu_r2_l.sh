#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ppo_plan_gpu.py tests/test_trainer_gpu.py tests/test_zz_runner_gpu.py -q > gpurun_out/pytest_plan.log 2>&1; echo "plan tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/pytest_plan.log | tail -8
for d in 1; do
QA_DEFER_REWARD_TAIL=$d timeout 900 python bench.py --steps 10 --warmup 3 --no-tsc --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value > gpurun_out/bench_l$d.json 2> gpurun_out/bench_l$d.err; echo "bench defer=$d rc=$?"; python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_l$d.json") if l.startswith("{")][-1]
print({k: d.get(k) for k in ("value", "ms_per_step", "collection_ms", "learning_ms", "gpu_launches")}, d["e2e"]["value"])
PY
done

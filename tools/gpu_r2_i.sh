#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ppo_plan_gpu.py tests/test_linear_gpu.py tests/test_trainer_gpu.py -q > gpurun_out/pytest_plan.log 2>&1; echo "plan tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/pytest_plan.log | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-tsc --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"; python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/bench_g.json") if l.startswith("{")][-1]
print({k: d.get(k) for k in ("value", "ms_per_step", "collection_ms", "learning_ms", "gpu_launches", "disc_update_ms")}, d["e2e"]["value"], {k: d["roofline_gemm"][k] for k in ("us_per_launch", "dx_us_per_launch", "dw_us_per_launch", "frac")})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value --no-tsc > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
python tools/launch_breakdown.py gpurun_out/launches.csv 40 > gpurun_out/launch_breakdown.txt 2>&1; sed -n '/one PPO/,$p' gpurun_out/launch_breakdown.txt | grep -v gemm | head -40

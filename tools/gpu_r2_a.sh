#!/bin/bash
# Round 2, GPU call A: new kernels first (each under its own timeout), then the whole suite, the bench line, the launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_ppo_plan_gpu.py -q -x > gpurun_out/pytest_plan.log 2>&1; echo "plan tests rc=$?"; tail -25 gpurun_out/pytest_plan.log
timeout 300 python -m pytest tests/test_advice_gpu.py "tests/test_env_gpu.py" -q > gpurun_out/pytest_env.log 2>&1; echo "env/advice tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_env.log | tail -12
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ppo_plan_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
python tools/launch_breakdown.py gpurun_out/launches.csv > gpurun_out/launch_breakdown.txt 2>&1; head -60 gpurun_out/launch_breakdown.txt

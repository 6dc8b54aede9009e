#!/bin/bash
# plan + trainer tests after the small-kernel rewrite, TSC legs, bench line, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ppo_plan_gpu.py tests/test_trainer_gpu.py -q -x > gpurun_out/pytest_plan.log 2>&1; echo "plan+trainer tests rc=$?"; tail -4 gpurun_out/pytest_plan.log
timeout 600 python bench.py --workload tsc_teacher --steps 3 --warmup 3 > gpurun_out/bench_tsc_teacher.json 2> gpurun_out/bench_tsc_teacher.err; echo "tsc_teacher rc=$?"; cat gpurun_out/bench_tsc_teacher.json; tail -3 gpurun_out/bench_tsc_teacher.err
timeout 600 python bench.py --workload tsc_student --steps 3 --warmup 3 > gpurun_out/bench_tsc_student.json 2> gpurun_out/bench_tsc_student.err; echo "tsc_student rc=$?"; cat gpurun_out/bench_tsc_student.json; tail -3 gpurun_out/bench_tsc_student.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-tsc > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/bench.json") if l.startswith("{")][-1]
print({k: d[k] for k in ("value", "ms_per_step", "collection_ms", "learning_ms", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["us_per_launch"], d.get("disc_update_ms"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value --no-tsc > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
python tools/launch_breakdown.py gpurun_out/launches.csv 30 > gpurun_out/launch_breakdown.txt 2>&1; sed -n '/one PPO/,$p' gpurun_out/launch_breakdown.txt | head -70

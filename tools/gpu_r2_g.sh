#!/bin/bash
# GEMM rework: parity first (own timeouts), then the micro-benchmark table with / without the L2 prefetch, then the bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_linear_gpu.py -q -x > gpurun_out/pytest_linear.log 2>&1; echo "linear tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/pytest_linear.log | tail -8
timeout 600 python -m pytest tests/test_ppo_plan_gpu.py tests/test_disc_plan_gpu.py -q > gpurun_out/pytest_plan.log 2>&1; echo "plan tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/pytest_plan.log | tail -8
timeout 300 python tools/bench_linear.py > gpurun_out/bench_linear_a.txt 2>&1; echo "bench_linear rc=$?"; grep -v "^ *4096" gpurun_out/bench_linear_a.txt
QA_TC_PREFETCH=0 timeout 300 python tools/bench_linear.py > gpurun_out/bench_linear_b.txt 2>&1; echo "bench_linear (no prefetch) rc=$?"; grep -v "^ *4096" gpurun_out/bench_linear_b.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-tsc --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"; python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/bench_g.json") if l.startswith("{")][-1]
print({k: d.get(k) for k in ("value", "ms_per_step", "collection_ms", "learning_ms", "gpu_launches", "disc_update_ms")}, d["e2e"]["value"], {k: d["roofline_gemm"][k] for k in ("us_per_launch", "dx_us_per_launch", "dw_us_per_launch", "frac")})
PY

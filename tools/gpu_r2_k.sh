#!/bin/bash
# K2 after moving the euler angles to the env warps + early history store: parity, phase trace, roofline legs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_gpu.py tests/test_advice_gpu.py tests/test_zz_runner_gpu.py -q > gpurun_out/pytest_env.log 2>&1; echo "env tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_env.log | tail -8
timeout 300 python tools/k2_trace.py --envs 4096 --reps 20 > gpurun_out/k2_trace_r2h.txt 2>&1; echo "trace rc=$?"; tail -26 gpurun_out/k2_trace_r2h.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-tsc --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"; python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/bench_k.json") if l.startswith("{")][-1]
print({k: d.get(k) for k in ("value", "ms_per_step", "collection_ms", "learning_ms")}, "k2 us", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "32768:", d["roofline_32768"]["us_per_launch"], d["roofline_32768"]["frac"])
PY

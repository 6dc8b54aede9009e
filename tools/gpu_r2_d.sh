#!/bin/bash
# K2 after the 4-role scalar split + PDL: parity, phase trace, roofline legs (with / without PDL)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_gpu.py tests/test_advice_gpu.py tests/test_zz_runner_gpu.py -q > gpurun_out/pytest_env.log 2>&1; echo "env tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_env.log | tail -8
timeout 300 python tools/k2_trace.py --envs 4096 > gpurun_out/k2_trace_r2.txt 2>&1; echo "trace rc=$?"; tail -26 gpurun_out/k2_trace_r2.txt
for pdl in 1 0; do
QA_K2_PDL=$pdl timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value --no-tsc > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"; python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_pdl$pdl.json") if l.startswith("{")][-1]
print("pdl=$pdl", {k: d[k] for k in ("value", "ms_per_step", "collection_ms", "learning_ms")}, "k2 us", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "32768:", d["roofline_32768"].get("us_per_launch"), d["roofline_32768"].get("frac"))
PY
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -x -q > gpurun_out/pytest_env.log 2>&1; echo "pytest-env rc=$?"; tail -15 gpurun_out/pytest_env.log
timeout 600 python tools/k2_trace.py --envs 4096 > gpurun_out/k2_trace_v3.txt 2>&1; echo "trace rc=$?"; tail -28 gpurun_out/k2_trace_v3.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_v3.json')); print(d['value'], d['e2e']['value'], d['collection_ms'], d['learning_ms'], d['roofline']['us_per_launch'], d['roofline']['frac'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_now.err; python -c "
import json; d=json.load(open('gpurun_out/bench_now.json')); print('value',d['value'],'e2e', d['e2e']['value'],'collect', d['collection_ms'],'learn', d['learning_ms'],'k2us', d['roofline']['us_per_launch'],'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trainer_gpu.py -x -q -m gpu -k "fused_rollout or act_and_disc" > gpurun_out/pytest_one.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_one.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_now.err; python -c "
import json; d=json.load(open('gpurun_out/bench_now.json')); print('value',d['value'],'e2e', d['e2e']['value'],'collect', d['collection_ms'],'learn', d['learning_ms'],'k2us', d['roofline']['us_per_launch'],'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_now.err; python -c "
import json; d=json.load(open('gpurun_out/bench_now.json')); print('value',d['value'],'e2e', d['e2e']['value'],'collect', d['collection_ms'],'learn', d['learning_ms'],'k2us', d['roofline']['us_per_launch'],'frac', d['roofline']['frac'], 'disc_ms', d.get('disc_update_ms'))"

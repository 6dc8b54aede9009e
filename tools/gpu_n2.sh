#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'collect', d['collection_ms'], 'learn', d['learning_ms'])"

#!/bin/bash
# 8 x B200: K31 vs NCCL at 8 ranks, BBC bench (config 5: 32 768 envs sharded 8 ways)
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
timeout 200 $TR tools/bench_peer.py 2>&1 | grep "us per call"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; grep '^{' gpurun_out/bench_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','collection_ms','learning_ms','n_gpus')}, d['e2e']['value'], d['config']['collectives_per_optimiser_step'])"
tail -2 gpurun_out/bench_n$N.err

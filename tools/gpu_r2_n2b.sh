#!/bin/bash
# 2 x B200 after the scheduling changes: NCCL results parity + BBC bench at 2 ranks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_nccl_gpu.py -q > gpurun_out/pytest_nccl.log 2>&1; echo "nccl tests rc=$?"; tail -3 gpurun_out/pytest_nccl.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_n2.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','collection_ms','learning_ms','n_gpus')}, d['e2e']['value'], d['config']['collectives_per_optimiser_step'])"
tail -3 gpurun_out/bench_n2.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/bench_ref_n2.json

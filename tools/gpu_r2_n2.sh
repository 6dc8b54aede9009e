#!/bin/bash
# 2 x B200: NCCL results parity, BBC bench at 2 ranks (single all-reduce default vs three), TSC student at 2 ranks (config 4)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_nccl_gpu.py -q > gpurun_out/pytest_nccl.log 2>&1; echo "nccl tests rc=$?"; tail -5 gpurun_out/pytest_nccl.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_n2.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','collection_ms','learning_ms','n_gpus')}, d['e2e']['value'], d['config']['collectives_per_optimiser_step'])"
QA_SINGLE_ALLREDUCE=0 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_three.json 2> gpurun_out/bench_n2_three.err; echo "bench n2 (3 collectives) rc=$?"; grep '^{' gpurun_out/bench_n2_three.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','collection_ms','learning_ms','n_gpus')}, d['e2e']['value'], d['config']['collectives_per_optimiser_step'])"
timeout 900 $TR bench.py --gpus 2 --workload tsc_student --steps 3 --warmup 2 > gpurun_out/bench_tsc_student_n2.json 2> gpurun_out/bench_tsc_student_n2.err; echo "tsc_student n2 rc=$?"; cat gpurun_out/bench_tsc_student_n2.json; tail -3 gpurun_out/bench_tsc_student_n2.err
timeout 600 $TR bench.py --gpus 2 --workload tsc_teacher --steps 3 --warmup 2 > gpurun_out/bench_tsc_teacher_n2.json 2> gpurun_out/bench_tsc_teacher_n2.err; echo "tsc_teacher n2 rc=$?"; cat gpurun_out/bench_tsc_teacher_n2.json; tail -3 gpurun_out/bench_tsc_teacher_n2.err

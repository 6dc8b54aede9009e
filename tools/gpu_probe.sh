#!/bin/bash
# K2 phase trace + launch list of one full iteration (rollout graph + PPO update)
mkdir -p gpurun_out
timeout 600 python tools/k2_trace.py --envs 4096 32768 > gpurun_out/k2_trace.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/k2_trace.txt | tail -40
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 -c 9000 --csv --log-file gpurun_out/launches_iter.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1; echo "ncu-list rc=$?"

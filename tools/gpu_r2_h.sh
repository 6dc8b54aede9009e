#!/bin/bash
# launch list of one iteration after the GEMM rework + ncu --set full of the pair GEMM (fwd / dX+act / dW) on two layer shapes
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value --no-tsc > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
python tools/launch_breakdown.py gpurun_out/launches.csv 40 > gpurun_out/launch_breakdown.txt 2>&1; sed -n '/one PPO/,$p' gpurun_out/launch_breakdown.txt | head -75
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32 --launch-skip 6 -c 3 -f -o gpurun_out/k7_r2_a python tools/bench_linear.py --only 24576,512,671 > gpurun_out/ncu_k7_a.log 2>&1; echo "ncu a rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32 --launch-skip 6 -c 3 -f -o gpurun_out/k7_r2_c python tools/bench_linear.py --only 24576,256,512 > gpurun_out/ncu_k7_c.log 2>&1; echo "ncu c rc=$?"

#!/bin/bash
# GEMM micro-benchmark table + full ncu captures (source-level) of the three contraction kinds on two layer shapes
mkdir -p gpurun_out
timeout 300 python tools/bench_linear.py > gpurun_out/bench_linear.txt 2>&1; echo "bench_linear rc=$?"; cat gpurun_out/bench_linear.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32 --launch-skip 6 -c 3 -f -o gpurun_out/k7_c2 python tools/bench_linear.py --only 24576,256,512 > gpurun_out/ncu_k7_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32 --launch-skip 6 -c 3 -f -o gpurun_out/k7_a1 python tools/bench_linear.py --only 24576,512,101 > gpurun_out/ncu_k7_a1.log 2>&1; echo "ncu a1 rc=$?"
ls -la gpurun_out/*.ncu-rep

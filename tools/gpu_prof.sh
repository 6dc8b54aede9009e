#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:k_gemm_tf32 --launch-skip 6 -c 3 -f -o gpurun_out/k7_full python tools/bench_linear.py --only 24576,512,671 > gpurun_out/ncu_k7.log 2>&1; echo "ncu-k7 rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"k_gae|k_gather|k_clip_adam|k_grad_sumsq" --launch-skip 0 -c 60 -f -o gpurun_out/misc_full python tools/prof_trainer_kernels.py > gpurun_out/ncu_misc.log 2>&1; echo "ncu-misc rc=$?"; tail -2 gpurun_out/ncu_misc.log

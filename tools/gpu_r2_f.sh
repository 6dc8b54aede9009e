#!/bin/bash
# round-2 evidence pass: smoke, whole GPU suite, bench (both arms), launch list of the bench command, ncu --set full of K2 and of the discriminator step
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value --no-tsc > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
python tools/launch_breakdown.py gpurun_out/launches.csv 30 > gpurun_out/launch_breakdown.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_post_physics_bbc_tiled --launch-skip 30 -c 2 \
  -f -o gpurun_out/k2_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-fp32-value --no-tsc > gpurun_out/ncu_k2.log 2>&1; echo "ncu-k2 rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt

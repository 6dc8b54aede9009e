#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list of the bench command, full ncu capture of K2.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_post_physics_bbc_tiled --launch-skip 30 -c 2 \
  -f -o gpurun_out/k2_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/ncu_k2.log 2>&1; echo "ncu-k2 rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt

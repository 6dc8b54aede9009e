#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -u -m pytest tests -m gpu -x -v -p no:cacheprovider --deselect tests/test_env_gpu.py --deselect tests/test_gae_gpu.py --deselect tests/test_linear_gpu.py --deselect tests/test_expert.py -k "dagger or tsc or depth" > gpurun_out/all2.log 2>&1; echo "all rc=$?"; tail -8 gpurun_out/all2.log

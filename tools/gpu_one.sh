#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_trainer_gpu.py -x -q -m gpu -k "dagger" > gpurun_out/pytest_one.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_one.log

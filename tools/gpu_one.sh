#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tsc_env.py -x -q -m gpu > gpurun_out/pytest_tsc.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_tsc.log

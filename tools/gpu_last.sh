#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -u -m pytest -m gpu -q -p no:cacheprovider tests/test_tsc_student.py "tests/test_trainer_gpu.py::test_runner_learn_books_episodes_logs_reference_tags_and_round_trips_checkpoint" "tests/test_tsc_env.py::test_tsc_runner_teacher_iteration_runs_and_rewards_match_torch_path" > gpurun_out/new.log 2>&1; echo "rc=$?"; grep -E "^E |passed|failed|Error" gpurun_out/new.log | head -40

#!/bin/bash
# discriminator schedule: kernel + step parity, then the bench line's disc_update_ms
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_disc_plan_gpu.py -q > gpurun_out/pytest_disc.log 2>&1; echo "disc tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|assert" gpurun_out/pytest_disc.log | tail -30
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_zz_runner_gpu.py -q > gpurun_out/pytest_tr.log 2>&1; echo "trainer tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_tr.log | tail -10
timeout 900 python bench.py --steps 10 --warmup 3 --no-tsc --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/bench.json") if l.startswith("{")][-1]
print({k: d.get(k) for k in ("value", "ms_per_step", "collection_ms", "learning_ms", "gpu_launches", "disc_update_ms")}, d["e2e"]["value"], d["roofline"]["us_per_launch"])
PY
tail -5 gpurun_out/bench.err

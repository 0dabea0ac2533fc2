#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
python scripts/one_step.py 4096 tf32 3 2>&1 | tail -1
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
    print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "share_of_step")}, "shares", d["roofline"]["kernel_share_ms_per_step"])
    print("cpu", d["cpu_baseline"], "sampling", round(d["sampling"]["value"]), "clocks", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY

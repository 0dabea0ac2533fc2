#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "tcgen05 or tf32" -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
grep -E "tcgen05 forward|passed|failed|Error|error" gpurun_out/pytest_tc.log | head -20
timeout 900 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for cfg in "tf32 4096" "fp32 4096"; do
  set -- $cfg
  timeout 900 python bench.py --steps 3 --warmup 3 --math $1 --batch $2 --cpu-sample 128 > gpurun_out/bench_$1_$2.log 2> gpurun_out/bench_$1_$2.err
  echo "exit $?" >> gpurun_out/bench_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$1_$2.log").read().strip().splitlines()[-1])
    print("$1 B=$2", round(d["value"]), "gestures/s", round(d["ms_per_step"],1), "ms/step", d["roofline"]["kernel_share_ms_per_step"], "sampling", round(d["sampling"]["value"]))
except Exception as e:
    print("bench $1 $2 failed", e); print(open("gpurun_out/bench_$1_$2.err").read()[-1500:])
PY
done

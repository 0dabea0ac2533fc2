#!/bin/bash
# bench (N=1 default flags) + reference arm + one `ncu --set full` capture of the dominant kernel
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
    r = d["roofline"]; print("roofline", {k: r[k] for k in ("bound", "kernel", "achieved", "peak", "frac", "traffic", "share_of_step")}, "tensor", r["tensor"]["frac"], "shares", r["kernel_share_ms_per_step"])
    print("cpu", d["cpu_baseline"], "sampling", round(d["sampling"]["value"]), "clocks", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; tail -c 400 gpurun_out/bench_ref.log
python scripts/one_step.py 4096 tf32 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_fwd_kernel -s 8 -c 3 -f -o gpurun_out/r01_lstm_tc_fwd \
  python scripts/one_step.py 4096 tf32 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -1 gpurun_out/plain.log

#!/bin/bash
# bench.py (plain) and, after it exits 0, the ncu launch list of the same command.
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; rc=$?; echo "bench exit $rc" >> gpurun_out/bench.err
tail -2 gpurun_out/bench.err
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
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; tail -c 600 gpurun_out/bench_ref.log
if [ $rc -eq 0 ]; then
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"
  python scripts/parse_ncu_list.py gpurun_out/launches_bench.csv gpurun_out/launch_summary_bench.txt "ncu launch list: bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph (18 training steps: 3 warm-up, 9 kernel-share probes, 2 timed, 2 bracketed, 2 end-to-end; + 12 sampling calls)" 18
fi
exit 0

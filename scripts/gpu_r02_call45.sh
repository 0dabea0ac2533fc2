#!/bin/bash
# wave-aligned batch sizes for the sweep lines: the critic phase's generator launches (5 B gestures) fit the 148 SMs in one wave
set -u
mkdir -p gpurun_out
run() { # hidden seq batch tag
timeout 1500 python bench.py --hidden $1 --seq $2 --batch $3 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H$1_T$2_B$3_$4.json 2> gpurun_out/r02_bench_H$1_T$2_B$3_$4.err
echo "H=$1 T=$2 B=$3 rc=$?"; tail -n 2 gpurun_out/r02_bench_H$1_T$2_B$3_$4.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H$1_T$2_B$3_$4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"
}
run 256 256 460 v13
run 512 256 228 v13
run 1024 256 100 v13

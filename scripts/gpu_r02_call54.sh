#!/bin/bash
# final sweep table of the round (all rungs, same build)
set -u
mkdir -p gpurun_out
run() { # hidden seq batch tag
timeout 1500 python bench.py --hidden $1 --seq $2 --batch $3 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H$1_T$2_B$3_$4.json 2> gpurun_out/r02_bench_H$1_T$2_B$3_$4.err
echo "H=$1 T=$2 B=$3 rc=$?"; tail -n 1 gpurun_out/r02_bench_H$1_T$2_B$3_$4.err | cut -c1-200
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H$1_T$2_B$3_$4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step']['achieved'], d['roofline']['whole_step']['frac'])"
}
run 128 256 1024 v17
run 128 256 2048 v17
run 256 256 512 v17
run 512 256 256 v17
run 1024 256 128 v17

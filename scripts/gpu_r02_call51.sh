#!/bin/bash
# bias-gradient column sums fused into the unchunk pass: parity + bench lines
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s > gpurun_out/r02_gpu_scaled_v18.log 2>&1; echo "scaled rc=$?"
grep -E "passed|failed|Error|assert|grads|^E " gpurun_out/r02_gpu_scaled_v18.log | cut -c1-200 | head -20
run() { # hidden seq batch tag
timeout 1500 python bench.py --hidden $1 --seq $2 --batch $3 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H$1_T$2_B$3_$4.json 2> gpurun_out/r02_bench_H$1_T$2_B$3_$4.err
echo "H=$1 T=$2 B=$3 rc=$?"; tail -n 2 gpurun_out/r02_bench_H$1_T$2_B$3_$4.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H$1_T$2_B$3_$4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"
}
run 128 256 1024 v15
run 512 256 256 v15

#!/bin/bash
# persistent tcgen05 recurrence for H = 128 (one launch per layer): parity, A/B against the per-timestep kernels, bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s -x > gpurun_out/r02_gpu_scaled_v8.log 2>&1; echo "scaled rc=$?"
grep -E "passed|failed|Error|assert|grads|^E " gpurun_out/r02_gpu_scaled_v8.log | cut -c1-220 | head -30
( for b in 4096 5120 9472; do timeout 300 python scripts/scaled_forward.py 128 256 $b; WGG_LSTM128_PERSIST=0 timeout 300 python scripts/scaled_forward.py 128 256 $b; done ) > gpurun_out/r02_scaled_forward_persist128.log 2>&1
cat gpurun_out/r02_scaled_forward_persist128.log | cut -c1-600
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v8.json 2> gpurun_out/r02_bench_H128_T256_B1024_v8.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_bench_H128_T256_B1024_v8.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H128_T256_B1024_v8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"

#!/bin/bash
# layer-0 projection v2 (register-resident weights, staged prototypes), head-forward kernel, 7-stage BPTT ring: parity + bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s > gpurun_out/r02_gpu_scaled_v15.log 2>&1; echo "scaled rc=$?"
grep -E "passed|failed|Error|assert|grads|^E " gpurun_out/r02_gpu_scaled_v15.log | cut -c1-220 | head -24
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v12.json 2> gpurun_out/r02_bench_H128_T256_B1024_v12.err
echo "rc=$?"; tail -n 2 gpurun_out/r02_bench_H128_T256_B1024_v12.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H128_T256_B1024_v12.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"
timeout 900 python scripts/prof_sites.py 1024 tf32 128 256 > gpurun_out/r02_prof_sites_H128_T256_B1024_v12.log 2>&1; echo "prof rc=$?"
sed -n '/filter kernel/,$p' gpurun_out/r02_prof_sites_H128_T256_B1024_v12.log | cut -c1-160 | head -12

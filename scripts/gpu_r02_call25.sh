#!/bin/bash
# scaled regime: tcgen05 split-K weight/input-gradient GEMMs + layer-0 projection kernel: parity, then the H=128 bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s > gpurun_out/r02_gpu_scaled_v6.log 2>&1; echo "scaled rc=$?"
grep -E "passed|failed|Error|assert|grads" gpurun_out/r02_gpu_scaled_v6.log | cut -c1-220 | head -30
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v6.json 2> gpurun_out/r02_bench_H128_T256_B1024_v6.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_bench_H128_T256_B1024_v6.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H128_T256_B1024_v6.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"
timeout 600 python scripts/prof_sites.py 1024 tf32 128 256 > gpurun_out/r02_prof_sites_H128_T256_B1024_v6.log 2>&1; echo "prof rc=$?"
sed -n '/filter kernel/,$p' gpurun_out/r02_prof_sites_H128_T256_B1024_v6.log | cut -c1-160 | head -40

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s > gpurun_out/r02_gpu_scaled_v17.log 2>&1; echo "scaled rc=$?"
grep -E "passed|failed|Error|assert|grads|^E " gpurun_out/r02_gpu_scaled_v17.log | cut -c1-220 | head -24
timeout 900 python scripts/prof_sites.py 1024 tf32 128 256 > gpurun_out/r02_prof_sites_H128_T256_B1024_v13.log 2>&1; echo "prof rc=$?"
sed -n '/filter kernel/,$p' gpurun_out/r02_prof_sites_H128_T256_B1024_v13.log | cut -c1-160 | head -8

#!/bin/bash
# re-entry validation of HEAD: full GPU suite; per-call-site breakdown of the scaled step (H=128, T=256, B=1024)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest_gpu_reentry.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_reentry.log
tail -4 gpurun_out/r02_pytest_gpu_reentry.log | cut -c1-400
timeout 600 python scripts/prof_sites.py 1024 tf32 128 256 > gpurun_out/r02_prof_sites_H128_T256_B1024.log 2>&1; echo "prof rc=$?"
cat gpurun_out/r02_prof_sites_H128_T256_B1024.log | cut -c1-160 | head -90

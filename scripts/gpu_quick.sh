#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-400
python scripts/one_step.py 4096 tf32 5 2>&1 | tail -1
timeout 600 python scripts/prof_sites.py 4096 tf32 > gpurun_out/prof_sites.log 2>&1; sed -n '/filter kernel/,$p' gpurun_out/prof_sites.log | head -12

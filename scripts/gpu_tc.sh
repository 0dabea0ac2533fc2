#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "tensor_core or tcgen05" -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
grep -E "tcgen05 forward|^tf32|passed|failed|Error|error|assert" gpurun_out/pytest_tc.log | cut -c1-1500 | head -30
timeout 900 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/prof_sites.py 4096 tf32 > gpurun_out/prof_sites.log 2>&1; tail -32 gpurun_out/prof_sites.log

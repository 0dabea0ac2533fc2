#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s > gpurun_out/r02_gpu_scaled.log 2>&1
echo "rc=$?" >> gpurun_out/r02_gpu_scaled.log
grep -E "tcgen05 GEMM|passed|failed|Error|assert|H=" gpurun_out/r02_gpu_scaled.log | head -30

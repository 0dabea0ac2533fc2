#!/bin/bash
# persistent H = 128 recurrence v3 (setmaxnreg: no spills): parity, timings, ncu
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -x > gpurun_out/r02_gpu_scaled_v10.log 2>&1; echo "scaled rc=$?"
grep -E "passed|failed|Error|assert|^E " gpurun_out/r02_gpu_scaled_v10.log | cut -c1-220 | head -10
( for b in 4096 5120 9472; do timeout 300 python scripts/scaled_forward.py 128 256 $b; done ) > gpurun_out/r02_scaled_forward_persist128_v3.log 2>&1
cat gpurun_out/r02_scaled_forward_persist128_v3.log | cut -c1-700
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm128_tc_fwd_kernel -s 1 -c 1 -f -o gpurun_out/r02_ncu_lstm128_v3 python scripts/scaled_forward.py 128 256 9472 nograph > gpurun_out/r02_ncu_lstm128_v3.log 2>&1
echo "ncu rc=$?"

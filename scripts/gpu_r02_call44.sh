#!/bin/bash
# cluster-pair TMA multicast of the weight slabs in the per-timestep forward kernel (WGG_STEP_CLUSTER=2): parity on small T first,
# then A/B timings at H = 256 / 512
set -u
mkdir -p gpurun_out
WGG_STEP_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s -x -k "large_batch" > gpurun_out/r02_gpu_scaled_cluster.log 2>&1; rc=$?; echo "scaled(cluster) rc=$rc"
grep -E "passed|failed|Error|assert|grads|^E " gpurun_out/r02_gpu_scaled_cluster.log | cut -c1-220 | head -12
if [ $rc -ne 0 ]; then exit 0; fi
( for c in 1 2; do for cfg in "256 256 2560" "512 256 1280"; do echo "CLUSTER=$c"; WGG_STEP_CLUSTER=$c timeout 200 python scripts/scaled_forward.py $cfg; done; done ) > gpurun_out/r02_scaled_forward_cluster.log 2>&1
cat gpurun_out/r02_scaled_forward_cluster.log | cut -c1-330

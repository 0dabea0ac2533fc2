#!/bin/bash
# round-2 GPU call 1: parity of the benchmarked (tf32 / tf32x3) path + the existing suite + compute-sanitizer
mkdir -p gpurun_out
nvidia-smi > gpurun_out/gpu.txt 2>&1
(free -g; nproc) > gpurun_out/host.txt 2>&1
timeout 2400 python -m pytest tests/test_gpu_parity_tc.py -m gpu -q --tb=short --durations=25 > gpurun_out/r02_tc_parity.log 2>&1
echo "rc=$?" >> gpurun_out/r02_tc_parity.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short > gpurun_out/r02_gpu_parity.log 2>&1
echo "rc=$?" >> gpurun_out/r02_gpu_parity.log
for tool in synccheck racecheck memcheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 30 python -m pytest "tests/test_gpu_parity.py::test_tensor_core_modes_parity" -m gpu -x -q \
     > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_sanitizer_$tool.log
done
tail -3 gpurun_out/r02_tc_parity.log gpurun_out/r02_gpu_parity.log

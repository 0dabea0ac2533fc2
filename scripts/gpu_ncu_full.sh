#!/bin/bash
# one `ncu --set full` capture of the dominant kernel (after the same command exited 0 without ncu)
set -u
mkdir -p gpurun_out
python scripts/one_step.py 4096 tf32 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_fwd_kernel -s 8 -c 3 -f -o gpurun_out/r01_lstm_tc_fwd \
  python scripts/one_step.py 4096 tf32 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep

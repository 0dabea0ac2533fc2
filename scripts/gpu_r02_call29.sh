#!/bin/bash
# ncu --set full of the persistent H = 128 recurrence (first version), one launch
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm128_tc_fwd_kernel -s 1 -c 1 -f -o gpurun_out/r02_ncu_lstm128_v1 python scripts/scaled_forward.py 128 256 4096 nograph > gpurun_out/r02_ncu_lstm128_v1.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_ncu_lstm128_v1.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep

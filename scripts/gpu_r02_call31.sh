#!/bin/bash
# ncu --set full of the persistent H = 128 recurrence v2 (chunked), one launch, full machine
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm128_tc_fwd_kernel -s 1 -c 1 -f -o gpurun_out/r02_ncu_lstm128_v2 python scripts/scaled_forward.py 128 256 9472 nograph > gpurun_out/r02_ncu_lstm128_v2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_ncu_lstm128_v2.log | cut -c1-300

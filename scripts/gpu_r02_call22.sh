#!/bin/bash
# fused forward step kernel: ring depth A/B (2 / 3 / 4 stages) and one ncu --set full capture of a mid-sequence launch
mkdir -p gpurun_out
( for n in 2 3 4; do echo "WGG_STEP_NST=$n"; WGG_STEP_NST=$n timeout 300 python scripts/scaled_forward.py 128 256 4096; WGG_STEP_NST=$n timeout 300 python scripts/scaled_forward.py 512 256 1024; done ) > gpurun_out/r02_scaled_forward_nst.log 2>&1
cat gpurun_out/r02_scaled_forward_nst.log
WGG_STEP_NST=3 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_lstm_fwd_kernel -s 300 -c 2 -f -o gpurun_out/r02_ncu_step_fwd_H128 python scripts/scaled_forward.py 128 256 4096 nograph > gpurun_out/r02_ncu_step_fwd_H128.log 2>&1
WGG_STEP_NST=3 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_lstm_fwd_kernel -s 300 -c 2 -f -o gpurun_out/r02_ncu_step_fwd_H512 python scripts/scaled_forward.py 512 256 1024 nograph > gpurun_out/r02_ncu_step_fwd_H512.log 2>&1
ls -la gpurun_out/*.ncu-rep

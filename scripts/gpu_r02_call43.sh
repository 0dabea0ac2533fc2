#!/bin/bash
# ncu tensor-pipe evidence: the tcgen05 GEMM (input projection) at H = 512 / 1024; conv smem-attribute change sanity (disc tests)
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_nt_kernel -s 3 -c 1 -f -o gpurun_out/r02_ncu_gemm_tc_H512 python scripts/scaled_forward.py 512 256 1024 nograph > gpurun_out/r02_ncu_gemm_tc_H512.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_nt_kernel -s 3 -c 1 -f -o gpurun_out/r02_ncu_gemm_tc_H1024 python scripts/scaled_forward.py 1024 128 1024 nograph > gpurun_out/r02_ncu_gemm_tc_H1024.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity_tc.py -m gpu -q --timeout 900 -k "discriminator" > gpurun_out/r02_pytest_disc_v3.log 2>&1; echo "disc rc=$?"; tail -2 gpurun_out/r02_pytest_disc_v3.log | cut -c1-200

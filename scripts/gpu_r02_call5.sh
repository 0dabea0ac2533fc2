#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_tc.py -m gpu -q --tb=short -k "resynchronised" > gpurun_out/r02_tc_parity_resync.log 2>&1
echo "rc=$?" >> gpurun_out/r02_tc_parity_resync.log
# ncu --set full of one launch of each forward-kernel flavour (layer 1, B = 18944)
WGG_FWD2=1 WGG_FWD2_VAR=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm_tc_fwd2_kernel -s 5 -c 1 \
   -o gpurun_out/r02_ncu_fwd2_pair python scripts/lstm_fwd_ab.py > gpurun_out/r02_ncu_fwd2.log 2>&1
WGG_FWD2=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm_tc_fwd_kernel -s 5 -c 1 \
   -o gpurun_out/r02_ncu_fwd_single python scripts/lstm_fwd_ab.py > gpurun_out/r02_ncu_fwd1.log 2>&1
timeout 1500 python scripts/acceptance_run.py --epochs 100 --seeds 0,1,2 --out gpurun_out/r02_acceptance.json > gpurun_out/r02_acceptance.log 2>&1
echo "rc=$?" >> gpurun_out/r02_acceptance.log
tail -n 3 gpurun_out/r02_tc_parity_resync.log; tail -n 40 gpurun_out/r02_acceptance.log; ls -la gpurun_out/*.ncu-rep

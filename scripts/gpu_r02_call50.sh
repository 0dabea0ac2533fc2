#!/bin/bash
# ncu launch list of one eager scaled step (H = 128, T = 256, B = 1024)
set -u
mkdir -p gpurun_out
python scripts/one_step.py 1024 tf32 1 128 256 > gpurun_out/one_step_H128.log 2>&1 && cat gpurun_out/one_step_H128.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_H128.csv python scripts/one_step.py 1024 tf32 1 128 256 > gpurun_out/ncu_list_H128.log 2>&1; echo "ncu rc=$?"
python scripts/parse_ncu_list.py gpurun_out/launches_H128.csv gpurun_out/r02_ncu_launch_list_step_H128_T256_B1024.txt "ncu launch list, one step H=128 T=256 B=1024 tf32 (final round-2 build)" > /dev/null
head -24 gpurun_out/r02_ncu_launch_list_step_H128_T256_B1024.txt | cut -c1-150

#!/bin/bash
# Speed-of-light / memory sections for one launch set of every heavy kernel class (after the same command exited 0
# without ncu).  Only the CSV summary travels back (the report itself is large).
set -u
mkdir -p gpurun_out
python scripts/ncu_targets.py > gpurun_out/plain_targets.log 2>&1 && \
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section ComputeWorkloadAnalysis \
  --clock-control none --profile-from-start off \
  -k regex:'lstm_tc_(fwd|bwd|dx|dw)_kernel|conv_tc_(fwd|wgrad3|wgrad)_kernel|clip_adam_dev|head_bwd_tc|head_tc_kernel' \
  -c 36 -f -o /tmp/r01_kernels python scripts/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/r01_kernels.ncu-rep --page raw --csv > gpurun_out/r01_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/r01_kernels_raw.csv; tail -2 gpurun_out/ncu_targets.log

#!/bin/bash
# round-2 GPU call 2: tc parity re-run, smoke, bench in both tensor-core modes, ncu launch list + DRAM traffic of one step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_tc.py -m gpu -q --tb=short > gpurun_out/r02_tc_parity.log 2>&1
echo "rc=$?" >> gpurun_out/r02_tc_parity.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r02_bench_tf32.json 2> gpurun_out/r02_bench_tf32.err
echo "rc=$?" >> gpurun_out/r02_bench_tf32.err
timeout 900 python bench.py --steps 8 --warmup 3 --math tf32x3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_tf32x3.json 2> gpurun_out/r02_bench_tf32x3.err
echo "rc=$?" >> gpurun_out/r02_bench_tf32x3.err
python scripts/one_step.py 4096 tf32 3 > gpurun_out/r02_one_step.log 2>&1
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/r02_step_launches.csv python scripts/one_step.py 4096 tf32 1 > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r02_ncu_list.log
python scripts/parse_ncu_list.py gpurun_out/r02_step_launches.csv gpurun_out/r02_launch_summary.txt "ncu launch list, one step B=4096 tf32 (round 2 baseline)" 1 > /dev/null 2>&1
python scripts/ncu_traffic.py gpurun_out/r02_step_launches.csv lstm_tc_fwd_kernel 4096 tf32 gpurun_out/r02_ncu_traffic_lstm_tc_fwd_kernel.json > /dev/null 2>&1
tail -n 3 gpurun_out/r02_tc_parity.log
tail -n 2 gpurun_out/r02_smoke.log

#!/bin/bash
mkdir -p gpurun_out
( WGG_FWD_NCH=3 timeout 300 python scripts/lstm_fwd_ab.py; WGG_FWD_NCH=2 timeout 300 python scripts/lstm_fwd_ab.py ) > gpurun_out/r02_fwd_nch_ab.log 2>&1
WGG_FWD_NCH=2 timeout 600 python -m pytest tests/test_gpu_parity_tc.py tests/test_gpu_parity.py -m gpu -q --tb=short -k "generator_multi_tile or tcgen05_generator_forward or sampling_properties or cycles_at_scale" > gpurun_out/r02_fwd_nch2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02_fwd_nch2_tests.log
timeout 900 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short > gpurun_out/r02_gpu_scaled.log 2>&1
echo "rc=$?" >> gpurun_out/r02_gpu_scaled.log
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_H128_T256_B1024.json 2> gpurun_out/r02_bench_H128_T256_B1024.err
echo "rc=$?" >> gpurun_out/r02_bench_H128_T256_B1024.err
cat gpurun_out/r02_fwd_nch_ab.log; tail -n 3 gpurun_out/r02_fwd_nch2_tests.log; tail -n 15 gpurun_out/r02_gpu_scaled.log; tail -n 5 gpurun_out/r02_bench_H128_T256_B1024.err; head -c 1500 gpurun_out/r02_bench_H128_T256_B1024.json

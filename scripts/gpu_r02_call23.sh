#!/bin/bash
# fused step kernels v4 (8-unit chunks, full sectors, 1 CTA/SM): parity, ring depth A/B, ncu durations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s > gpurun_out/r02_gpu_scaled_v4.log 2>&1; echo "scaled rc=$?"
grep -E "passed|failed|Error|assert" gpurun_out/r02_gpu_scaled_v4.log | head -10
( for n in 2 3 4; do echo "WGG_STEP_NST=$n"; WGG_STEP_NST=$n timeout 300 python scripts/scaled_forward.py 128 256 4096; WGG_STEP_NST=$n timeout 300 python scripts/scaled_forward.py 512 256 1024; done ) > gpurun_out/r02_scaled_forward_nst.log 2>&1
cat gpurun_out/r02_scaled_forward_nst.log
WGG_STEP_NST=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_lstm_fwd_kernel -s 300 -c 1 -f -o gpurun_out/r02_ncu_step_fwd_H128_v4 python scripts/scaled_forward.py 128 256 4096 nograph > gpurun_out/r02_ncu_step_fwd_H128.log 2>&1
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v5.json 2> gpurun_out/r02_bench_H128_T256_B1024_v5.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_bench_H128_T256_B1024_v5.err
python -c "
import json
for f in ('r02_bench_H128_T256_B1024_v5',):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"

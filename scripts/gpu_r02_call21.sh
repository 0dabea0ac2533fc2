#!/bin/bash
# fused step kernels v2: fast cell math, pipelined epilogue loads, 2 CTAs/SM + programmatic dependent launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s > gpurun_out/r02_gpu_scaled_v3.log 2>&1; echo "scaled rc=$?"
grep -E "tcgen05|passed|failed|Error|assert|H=" gpurun_out/r02_gpu_scaled_v3.log | head -30
( for pdl in 1 0; do echo "WGG_PDL=$pdl"; WGG_PDL=$pdl timeout 300 python scripts/scaled_forward.py 128 256 4096; WGG_PDL=$pdl timeout 300 python scripts/scaled_forward.py 512 256 1024; done
  timeout 300 python scripts/scaled_forward.py 256 256 2048; timeout 300 python scripts/scaled_forward.py 1024 128 1024 ) > gpurun_out/r02_scaled_forward_v3.log 2>&1
cat gpurun_out/r02_scaled_forward_v3.log
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v4.json 2> gpurun_out/r02_bench_H128_T256_B1024_v4.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_bench_H128_T256_B1024_v4.err
python -c "
import json
for f in ('r02_bench_H128_T256_B1024_v4',):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"

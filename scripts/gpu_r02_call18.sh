#!/bin/bash
mkdir -p gpurun_out
( python scripts/scaled_forward.py 128 256 4096; python scripts/scaled_forward.py 256 256 2048; python scripts/scaled_forward.py 512 256 1024; python scripts/scaled_forward.py 1024 128 1024 ) > gpurun_out/r02_scaled_forward.log 2>&1
cat gpurun_out/r02_scaled_forward.log
# ncu --set full of the layer-1 input projection (M = T*B = 65536, N = 2048, K = 1024) at H = 512: 63 per-step launches of layer 0 come first
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_nt_kernel -s 63 -c 1 -o gpurun_out/r02_ncu_gemm_tc_H512 python scripts/scaled_forward.py 512 64 1024 > gpurun_out/r02_ncu_gemm_tc.log 2>&1
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v2.json 2> gpurun_out/r02_bench_H128_T256_B1024_v2.err
echo "rc=$?"
timeout 1200 python bench.py --hidden 512 --seq 256 --batch 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_H512_T256_B256.json 2> gpurun_out/r02_bench_H512_T256_B256.err
echo "rc=$?"
python -c "
import json
for f in ('r02_bench_H128_T256_B1024_v2','r02_bench_H512_T256_B256'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d.get('reference_cuda'))"
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# fused latent head (encoder) + fused tcgen05 LSTM step kernels (scaled regime): parity, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s > gpurun_out/r02_gpu_scaled_v2.log 2>&1; echo "scaled rc=$?"
grep -E "tcgen05|passed|failed|Error|assert|H=" gpurun_out/r02_gpu_scaled_v2.log | head -30
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "encoder or golden or graph or resynchronised" > gpurun_out/r02_gpu_enc.log 2>&1; echo "enc rc=$?"; tail -n 5 gpurun_out/r02_gpu_enc.log
( timeout 300 python scripts/scaled_forward.py 128 256 4096; timeout 300 python scripts/scaled_forward.py 256 256 2048; timeout 300 python scripts/scaled_forward.py 512 256 1024; timeout 300 python scripts/scaled_forward.py 1024 128 1024 ) > gpurun_out/r02_scaled_forward_v2.log 2>&1
cat gpurun_out/r02_scaled_forward_v2.log
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v3.json 2> gpurun_out/r02_bench_H128_T256_B1024_v3.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_bench_H128_T256_B1024_v3.err
python -c "
import json
for f in ('r02_bench_H128_T256_B1024_v3',):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'])"

#!/bin/bash
# chunked stash for the grad-carrying passes at H = 128 (persistent forward with stash, chunked BPTT step kernel): parity + bench A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -x -s > gpurun_out/r02_gpu_scaled_v13.log 2>&1; echo "scaled rc=$?"
grep -E "passed|failed|Error|assert|grads|^E " gpurun_out/r02_gpu_scaled_v13.log | cut -c1-220 | head -24
for sc in 1 0; do
WGG_LSTM128_STASH_CHUNK=$sc timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v11_sc$sc.json 2> gpurun_out/r02_bench_H128_T256_B1024_v11_sc$sc.err
echo "stash_chunk=$sc rc=$?"; tail -n 2 gpurun_out/r02_bench_H128_T256_B1024_v11_sc$sc.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H128_T256_B1024_v11_sc$sc.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"
done

#!/bin/bash
# persistent cluster-pair BPTT for H = 128: parity (3 repetitions: the kernel synchronises two CTAs through global memory), bench A/B
set -u
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python -m pytest tests/test_gpu_scaled.py -m gpu -q --tb=short -s -x -k "128" > gpurun_out/r02_gpu_scaled_pbwd_$i.log 2>&1; rc=$?; echo "scaled H=128 run $i rc=$rc"
grep -E "passed|failed|Error|assert|grads|^E " gpurun_out/r02_gpu_scaled_pbwd_$i.log | cut -c1-200 | head -8
if [ $rc -ne 0 ]; then exit 0; fi
done
for pb in 1 0; do
WGG_BPTT128_PERSIST=$pb timeout 900 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v14_pb$pb.json 2> gpurun_out/r02_bench_H128_T256_B1024_v14_pb$pb.err
echo "persist_bwd=$pb rc=$?"; tail -n 2 gpurun_out/r02_bench_H128_T256_B1024_v14_pb$pb.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H128_T256_B1024_v14_pb$pb.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"
done

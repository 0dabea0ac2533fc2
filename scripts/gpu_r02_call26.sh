#!/bin/bash
# tcgen05 conv kernels generalised to T = 128 h (256-point gestures of configs[3]): full GPU suite, scaled + default bench lines
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_pytest_gpu_convT.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_convT.log
grep -E "passed|failed|Error|assert " gpurun_out/r02_pytest_gpu_convT.log | cut -c1-300 | head -20
tail -3 gpurun_out/r02_pytest_gpu_convT.log | cut -c1-300
timeout 1200 python bench.py --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H128_T256_B1024_v7.json 2> gpurun_out/r02_bench_H128_T256_B1024_v7.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_bench_H128_T256_B1024_v7.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H128_T256_B1024_v7.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_B4096_tf32_convT.json 2> gpurun_out/r02_bench_B4096_tf32_convT.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_bench_B4096_tf32_convT.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_B4096_tf32_convT.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_share_ms_per_step'])"

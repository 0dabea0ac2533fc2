#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_keyboard.py -m gpu -q --tb=short > gpurun_out/r02_gpu_keyboard.log 2>&1
echo "rc=$?" >> gpurun_out/r02_gpu_keyboard.log
timeout 1500 python -m pytest tests/ -m gpu -q --tb=short -x > gpurun_out/r02_gpu_all.log 2>&1
echo "rc=$?" >> gpurun_out/r02_gpu_all.log
timeout 600 python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_gradbuf.json 2> gpurun_out/r02_bench_gradbuf.err
tail -n 8 gpurun_out/r02_gpu_keyboard.log; tail -n 8 gpurun_out/r02_gpu_all.log; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_gradbuf.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['sampling']['value'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eval.py -m gpu -q --tb=short > gpurun_out/r02_gpu_eval.log 2>&1
echo "rc=$?" >> gpurun_out/r02_gpu_eval.log
timeout 1500 python -m pytest tests/ -m gpu -q --tb=short -x > gpurun_out/r02_gpu_all.log 2>&1
echo "rc=$?" >> gpurun_out/r02_gpu_all.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 900 python bench.py > gpurun_out/r02_bench_tf32_v2.json 2> gpurun_out/r02_bench_tf32_v2.err
echo "rc=$?" >> gpurun_out/r02_bench_tf32_v2.err
timeout 1500 python scripts/acceptance_run.py --epochs 100 --seeds 0,1,2,3,4 --out gpurun_out/r02_acceptance.json > gpurun_out/r02_acceptance.log 2>&1
echo "rc=$?" >> gpurun_out/r02_acceptance.log
tail -n 5 gpurun_out/r02_gpu_eval.log; tail -n 5 gpurun_out/r02_gpu_all.log; tail -c 600 gpurun_out/r02_bench_reference_arm.json; tail -n 3 gpurun_out/r02_bench_tf32_v2.err

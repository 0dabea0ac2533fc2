#!/bin/bash
# scaled regime under data parallelism: 2 ranks, H = 128 / T = 256 / 1024 gestures per GPU
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 > gpurun_out/r02_bench_H128_T256_B1024_dp2.json 2> gpurun_out/r02_bench_H128_T256_B1024_dp2.err; echo "rc=$?"
tail -2 gpurun_out/r02_bench_H128_T256_B1024_dp2.err | cut -c1-300
tail -1 gpurun_out/r02_bench_H128_T256_B1024_dp2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['n_gpus'], d['sampling']['value'], d['config'].get('gradient_exchange'))"

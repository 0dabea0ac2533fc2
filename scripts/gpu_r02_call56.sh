#!/bin/bash
# 8 ranks on one box: the default bench line and the scaled first rung, final build
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench_B4096_tf32_dp8_final.json 2> gpurun_out/r02_bench_dp8_final.err; echo "default rc=$?"
tail -1 gpurun_out/r02_bench_B4096_tf32_dp8_final.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['sampling'].get('configs4'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --hidden 128 --seq 256 --batch 1024 --steps 3 --warmup 3 > gpurun_out/r02_bench_H128_T256_B1024_dp8.json 2> gpurun_out/r02_bench_H128_T256_B1024_dp8.err; echo "scaled rc=$?"
tail -1 gpurun_out/r02_bench_H128_T256_B1024_dp8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['n_gpus'], d['sampling']['value'])"

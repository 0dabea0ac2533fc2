#!/bin/bash
# 2-GPU sanity of the final build: torchrun bench line + DP gradient parity
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_B4096_tf32_dp2_final.json 2> gpurun_out/r02_bench_dp2_final.err; echo "bench dp2 rc=$?"
tail -1 gpurun_out/r02_bench_B4096_tf32_dp2_final.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['config'].get('gradient_exchange'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_parity.py tf32 > gpurun_out/r02_dp2_gradient_parity_tf32_final.log 2>&1; echo "dp_parity rc=$?"; tail -4 gpurun_out/r02_dp2_gradient_parity_tf32_final.log | cut -c1-240

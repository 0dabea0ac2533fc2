#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_parity.py fp32 > gpurun_out/dp_parity.log 2>&1; echo "dp_parity exit $?"; grep -v "^W\|^\[W\|warn" gpurun_out/dp_parity.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err

#!/bin/bash
# 2-GPU checks: gradient parity of the data-parallel step (injected noise and the product path's own global-noise draw)
# in tf32 and fp32, then the bench exactly as the driver launches it at N = 2 (and the reference arm under torchrun)
set -u
mkdir -p gpurun_out
N=${1:-2}
for mode in tf32 fp32; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_parity.py $mode > gpurun_out/r02_dp_parity_$mode.log 2>&1; echo "dp_parity $mode exit $?"
  grep "DP_\|replicas" gpurun_out/r02_dp_parity_$mode.log
done
/usr/bin/time -f "bench wall %e s" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench exit $?"
tail -c 900 gpurun_out/r02_bench_n$N.json; tail -n 3 gpurun_out/r02_bench_n$N.err
/usr/bin/time -f "reference arm wall %e s" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err; echo "ref arm exit $?"
tail -c 300 gpurun_out/r02_bench_ref_n$N.json; tail -n 2 gpurun_out/r02_bench_ref_n$N.err

#!/bin/bash
# 2-GPU checks: gradient parity of the data-parallel step (injected noise and the product path's own global-noise draw)
# in tf32 and fp32, then the bench exactly as the driver launches it at N = 2 (and the reference arm under torchrun)
set -u
mkdir -p gpurun_out
N=${1:-2}
for mode in ${MODES:-tf32}; do
  WGG_P2P=${P2P_PARITY:-1} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_parity.py $mode > gpurun_out/r02_dp_parity_$mode.log 2>&1; echo "dp_parity $mode exit $?"
  grep "DP_\|replicas\|P2P\|exchange\|Error\|error" gpurun_out/r02_dp_parity_$mode.log | head -12
done
for p2p in 1 0; do
S0=$SECONDS; WGG_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 16 --warmup 3 > gpurun_out/r02_bench_n${N}_p2p$p2p.json 2> gpurun_out/r02_bench_n${N}_p2p$p2p.err; echo "bench (WGG_P2P=$p2p) exit $? wall $((SECONDS-S0)) s"
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n${N}_p2p$p2p.json').read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config'].get('gradient_exchange'))"
tail -n 2 gpurun_out/r02_bench_n${N}_p2p$p2p.err
done
S0=$SECONDS; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err; echo "ref arm exit $? wall $((SECONDS-S0)) s"
tail -c 300 gpurun_out/r02_bench_ref_n$N.json; tail -n 2 gpurun_out/r02_bench_ref_n$N.err

#!/bin/bash
# sweep completion (H = 1024; larger batches at H = 128) + ncu tensor-pipe evidence for the tcgen05 GEMM and the recurrence kernels
set -u
mkdir -p gpurun_out
run() { # hidden seq batch tag
timeout 1500 python bench.py --hidden $1 --seq $2 --batch $3 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_H$1_T$2_B$3_$4.json 2> gpurun_out/r02_bench_H$1_T$2_B$3_$4.err
echo "H=$1 T=$2 B=$3 rc=$?"; tail -n 2 gpurun_out/r02_bench_H$1_T$2_B$3_$4.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_H$1_T$2_B$3_$4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['sampling']['value'], d['roofline']['whole_step'], d['roofline']['kernel_share_ms_per_step'])"
}
run 128 256 2048 v13
run 128 256 1888 v13
run 1024 256 128 v13
for h in 128 512; do
b=$((h==128 ? 9472 : 1024))
timeout 600 ncu --set full --clock-control none -k regex:"gemm_tc_nt_kernel|lstm128_tc_fwd_kernel|gemm_tc_lstm_fwd_kernel" -s 8 -c 3 -f -o gpurun_out/r02_ncu_scaled_fwd_H$h python scripts/scaled_forward.py $h 256 $b nograph > gpurun_out/r02_ncu_scaled_fwd_H$h.log 2>&1
echo "ncu H=$h rc=$?"
done

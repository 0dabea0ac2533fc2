#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/lstm_fwd_ab.py > gpurun_out/r02_fwd_biasfold_ab.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity_tc.py tests/test_gpu_parity.py -m gpu -q --tb=short -k "generator_multi_tile or tcgen05_generator_forward or sampling_properties or cycles_at_scale or tensor_core_modes or train_batch_multi_tile" > gpurun_out/r02_fwd_biasfold_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02_fwd_biasfold_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_biasfold.json 2> gpurun_out/r02_bench_biasfold.err
cat gpurun_out/r02_fwd_biasfold_ab.log; tail -n 12 gpurun_out/r02_fwd_biasfold_tests.log; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_biasfold.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_ms_per_step'], d['sampling']['value'])"

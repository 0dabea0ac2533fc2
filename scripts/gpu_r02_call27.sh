#!/bin/bash
# conv kernels templated on MULTI (T = 128 folds to the round-1 code): discriminator tests at T = 128/256/384, scaled tests,
# default bench A/B against the previous build
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_tc.py tests/test_gpu_scaled.py -m gpu -q --timeout 900 -k "discriminator or scaled or h128 or large_batch or any_hidden" > gpurun_out/r02_pytest_gpu_convT2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_convT2.log
grep -E "passed|failed|^E  " gpurun_out/r02_pytest_gpu_convT2.log | cut -c1-300 | head -20
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_B4096_tf32_convT2.json 2> gpurun_out/r02_bench_B4096_tf32_convT2.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_bench_B4096_tf32_convT2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_B4096_tf32_convT2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_share_ms_per_step'], d['clocks'])"

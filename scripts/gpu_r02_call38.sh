#!/bin/bash
# full validation of the current build: GPU suite, smoke, default bench (with cpu baseline + reference_cuda), reference arm
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest_gpu_full_v2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_full_v2.log
tail -4 gpurun_out/r02_pytest_gpu_full_v2.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke_v2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_v2.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02_bench_B4096_tf32_v4.json 2> gpurun_out/r02_bench_B4096_tf32_v4.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_B4096_tf32_v4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sampling']['value'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"

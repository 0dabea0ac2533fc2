#!/bin/bash
# A/B: critic-phase generator as two concurrent calls (WGG_SPLIT_GEN=1, default) against the single stacked call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "graph or golden" > gpurun_out/r02_split_tests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_split_tests.log
for s in 0 1 0 1; do
WGG_SPLIT_GEN=$s timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_split$s.json 2> gpurun_out/r02_bench_split$s.err; echo "split=$s rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_split$s.json').read().strip().splitlines()[-1]); print('split $s', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
done

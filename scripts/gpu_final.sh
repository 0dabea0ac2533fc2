#!/bin/bash
# what the driver runs at round end, in one call: build check, pytest -m gpu, smoke(), bench (ours + reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x --tb=short > gpurun_out/r02_gpu_all.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r02_smoke.log
S0=$SECONDS; timeout 900 python bench.py --impl reference --gpus 1 --steps 8 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference arm rc=$? wall $((SECONDS-S0)) s"
S0=$SECONDS; timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$? wall $((SECONDS-S0)) s"
tail -n 3 gpurun_out/r02_gpu_all.log; tail -n 2 gpurun_out/r02_smoke.log
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/r02_bench_reference_arm.json').read().strip().splitlines()[-1])
print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'ref', r['value'], 'ratio e2e', d['e2e']['value']/r['value'])
print('refcuda', d['reference_cuda']); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'])"

#!/bin/bash
# final validation of the round-2 build: what the driver runs (pytest -m gpu, smoke, both bench arms) + ncu launch lists
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -m gpu -q --tb=short > gpurun_out/r02_gpu_all_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_gpu_all_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r02_smoke_final.log
S0=$SECONDS; timeout 900 python bench.py --impl reference --gpus 1 --steps 8 --warmup 3 > gpurun_out/r02_bench_reference_arm_final.json 2> gpurun_out/r02_bench_reference_arm_final.err; echo "reference arm rc=$? wall $((SECONDS-S0)) s"
S0=$SECONDS; timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$? wall $((SECONDS-S0)) s"
tail -n 3 gpurun_out/r02_gpu_all_final.log | cut -c1-300; tail -n 2 gpurun_out/r02_smoke_final.log | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/r02_bench_reference_arm_final.json').read().strip().splitlines()[-1])
print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'ref', r['value'], 'ratio e2e', d['e2e']['value']/r['value'])
print('refcuda', d['reference_cuda']); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'], 'launches', d['gpu_launches'])"
# ncu launch lists (per-launch durations, cold-cache and serialised): default step, scaled step
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_default.csv python scripts/one_step.py 4096 tf32 1 > gpurun_out/ncu_list_default.log 2>&1; echo "ncu default rc=$?"
python scripts/parse_ncu_list.py gpurun_out/launches_default.csv gpurun_out/r02_ncu_launch_list_step_B4096_tf32_final.txt "ncu launch list, one step B=4096 tf32 (final round-2 build)" > /dev/null
head -12 gpurun_out/r02_ncu_launch_list_step_B4096_tf32_final.txt | cut -c1-140

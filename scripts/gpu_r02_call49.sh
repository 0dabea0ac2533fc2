#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/prof_sites.py 4096 tf32 > gpurun_out/r02_prof_sites_default_final.log 2>&1; echo "prof rc=$?"
sed -n '/filter kernel/,$p' gpurun_out/r02_prof_sites_default_final.log | cut -c1-150 | head -60

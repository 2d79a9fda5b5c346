#!/bin/bash
# Runs ON the GPU box: one full ncu capture (with source) of k_render<0> on the c3 slice for the CURRENT build.
#   gpurun --timeout 900 -- 'bash tools/render_ncu_source.sh TAG'
set -u
TAG=${1:-s}
mkdir -p gpurun_out
CMD="python bench.py --workload c3s --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_renderILi0 -s 3 -c 1 -f -o gpurun_out/${TAG}_render $CMD > gpurun_out/${TAG}_render.log 2>&1
ncu -i gpurun_out/${TAG}_render.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*

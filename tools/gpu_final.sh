#!/bin/bash
# Runs ON the GPU box (through gpurun): the round's record -- parity suite, launch list + full ncu capture of the render kernel,
# default bench line (c3, e2e, cpu baseline), optionally the other single-GPU configurations and the reference arm.
#   gpurun --timeout 2400 -- 'bash tools/gpu_final.sh r01 [all]'
set -u
TAG=${1:-r01}
ALL=${2:-}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
timeout 900 bash tools/profile_gpu.sh ${TAG} > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c3_1gpu.json 2> gpurun_out/${TAG}_bench_c3_1gpu.err
timeout 600 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2_1gpu.json 2> /dev/null
CFG="c3 c2"
if [ -n "$ALL" ]; then
  timeout 900 python bench.py --workload c4 --steps 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4_1gpu.json 2> /dev/null
  timeout 900 python bench.py --workload c5 --steps 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5_1gpu.json 2> /dev/null
  timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> /dev/null
  CFG="c3 c2 c4 c5"
fi
for f in $CFG; do python -c "
import json
j=json.loads(open('gpurun_out/${TAG}_bench_${f}_1gpu.json').read().strip().splitlines()[-1])
print('$f', round(j['value']), 'ms', round(j['ms_per_step'],2), j['stage_ms'], 'e2e', j.get('e2e',{}).get('value'), 'frac', round(j['roofline']['frac'],4), 'cpu', j.get('cpu_baseline',{}).get('value'))"; done

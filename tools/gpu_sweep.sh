#!/bin/bash
# development: the c3 slice under a sweep of one environment switch:  gpu_sweep.sh TAG VAR v1 v2 ...
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload c3s --render-profile > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python -c "
import json
j=json.loads(open('gpurun_out/${TAG}_$v.json').read().strip().splitlines()[-1]); print('$VAR','$v','render ms',round(j['stage_ms']['render'],3),'step',round(j['ms_per_step'],3))"
  tail -1 gpurun_out/${TAG}_$v.err
done

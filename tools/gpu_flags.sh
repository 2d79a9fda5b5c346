#!/bin/bash
# development: the c3 slice under every combination of the render switches
TAG=${1:-x}
mkdir -p gpurun_out
for g in 0 1; do for f in 0 1 2 3; do
  MH_RENDER_GRED=$g MH_RENDER_FLAGS=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload c3s > gpurun_out/${TAG}_g${g}f${f}.json 2> gpurun_out/${TAG}_g${g}f${f}.err
  python -c "
import json,sys
j=json.loads(open('gpurun_out/${TAG}_g${g}f${f}.json').read().strip().splitlines()[-1]); print('gred',$g,'flags',$f,'render ms',round(j['stage_ms']['render'],3),'step',round(j['ms_per_step'],3))"
done; done

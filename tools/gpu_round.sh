#!/bin/bash
# Runs ON the GPU box (through gpurun): the c3 profiling slice with the render phase shares, the parity suite, the full c3 line.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
set -u
TAG=${1:-x}
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B --workload c3s --render-profile > gpurun_out/${TAG}_c3s.json 2> gpurun_out/${TAG}_c3s.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
for f in gpurun_out/${TAG}_c3s gpurun_out/${TAG}_c3; do
  python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    j = json.loads(open(f + '.json').read().strip().splitlines()[-1])
    print(f, 'value', round(j['value']), 'ms', round(j['ms_per_step'], 2), 'render', round(j['stage_ms']['render'], 2), 'e2e', j.get('e2e', {}).get('value'))
except Exception as e:
    print(f, 'FAILED', e)
print(open(f + '.err').read()[-400:])
PY
done

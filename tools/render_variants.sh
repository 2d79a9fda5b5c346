#!/bin/bash
# Runs ON the GPU box (through gpurun): A/B of compile-time variants of the render kernel on the c3 profiling slice.
#   gpurun --timeout 1500 -- 'bash tools/render_variants.sh TAG "<XFLAGS variant 1>" "<XFLAGS variant 2>" ...'
# The LAST variant listed stays built (list the default last), then the parity suite runs on it.
set -u
TAG=${1:-v}; shift
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload c3s"
i=0
for XF in "$@"; do
  touch scene-aware-3d-multi-human_b200/csrc/mh_render.cu
  make -s -C scene-aware-3d-multi-human_b200/csrc XFLAGS="$XF" > gpurun_out/${TAG}_build$i.log 2>&1 || { echo "variant $i [$XF] BUILD FAILED"; tail -5 gpurun_out/${TAG}_build$i.log; i=$((i+1)); continue; }
  timeout 300 $B > gpurun_out/${TAG}_v$i.json 2> gpurun_out/${TAG}_v$i.err
  python - "$i" "$XF" gpurun_out/${TAG}_v$i <<'PY'
import json, sys
i, xf, f = sys.argv[1:4]
try:
    j = json.loads(open(f + '.json').read().strip().splitlines()[-1])
    print('variant', i, '[' + xf + ']', 'value', round(j['value']), 'ms', round(j['ms_per_step'], 3), 'render', round(j['stage_ms']['render'], 3), 'sil', j.get('loss_check'))
except Exception as e:
    print('variant', i, '[' + xf + ']', 'FAILED', e, open(f + '.err').read()[-600:])
PY
  i=$((i+1))
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log

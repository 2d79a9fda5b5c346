#!/bin/bash
# instrumented build (make XFLAGS=-DMH_RSTATS): pair statistics + phase shares on the c3 slice, block-level prune off / on
TAG=${1:-x}
mkdir -p gpurun_out
for f in 0 1; do
MH_RENDER_FLAGS=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workload c3s --render-profile > gpurun_out/${TAG}_stats$f.json 2> gpurun_out/${TAG}_stats$f.err
tail -5 gpurun_out/${TAG}_stats$f.err
done

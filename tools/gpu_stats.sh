#!/bin/bash
# instrumented build (make XFLAGS=-DMH_RSTATS): pair statistics + phase shares on the c3 slice
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workload c3s --render-profile "$@" > gpurun_out/${TAG}_stats.json 2> gpurun_out/${TAG}_stats.err
cat gpurun_out/${TAG}_stats.err | tail -5

#!/bin/bash
# instrumented build (make -C scene-aware-3d-multi-human_b200/csrc XFLAGS=-DMH_RSTATS, after touching mh_render.cu): pair statistics +
# phase shares of the render kernel on the c3 slice.  Rebuild without XFLAGS afterwards.
#   gpurun --timeout 600 -- 'bash tools/gpu_stats.sh TAG'
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workload c3s --render-profile > gpurun_out/${TAG}_stats.json 2> gpurun_out/${TAG}_stats.err
tail -5 gpurun_out/${TAG}_stats.err

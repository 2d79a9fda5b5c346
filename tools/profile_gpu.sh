#!/bin/bash
# Runs ON the GPU box (through gpurun): launch list of one bench command, one full ncu capture of the render kernel.
#   gpurun --timeout 1500 -- 'bash tools/profile_gpu.sh r01'
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --workload c3s --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
# every launch with its device time (cold-cache, serialised: compare SHARES); skip the input-synthesis launches
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
# the dominant kernel, full set, with source
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_renderILi0 -s 3 -c 1 -o gpurun_out/${TAG}_render $CMD > gpurun_out/${TAG}_render.log 2>&1
# the two tensor-core contractions (one launch each, after the warm-up cycles)
ncu --set full --clock-control none --kernel-name-base function -k regex:k_gemm_.*_tc -s 6 -c 2 -o gpurun_out/${TAG}_gemm_tc $CMD > gpurun_out/${TAG}_gemm_tc.log 2>&1
# the same command without a profiler, for the stage shares measured with CUDA events
$CMD > gpurun_out/${TAG}_bench_c3s.json 2> /dev/null
tail -c 600 gpurun_out/${TAG}_bench_c3s.json

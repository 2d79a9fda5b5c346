#!/bin/bash
# Runs ON a multi-GPU box (gpurun --gpus N): sharded == single over real NCCL, explicit and deferred sharding.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/dist_check_all.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
python tools/dist_check.py > gpurun_out/dist_check_run.log 2>&1
for W in $(seq 2 $N); do
  case $W in 2|4|8) ;; *) continue;; esac
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py >> gpurun_out/dist_check_run.log 2>&1
  DIST_CHECK_DEFERRED=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py >> gpurun_out/dist_check_run.log 2>&1
done
grep "^world" gpurun_out/dist_check_run.log
: > gpurun_out/dist_check_compare.txt
for W in $(seq 2 $N); do
  case $W in 2|4|8) ;; *) continue;; esac
  python tools/dist_check.py compare 1 $W >> gpurun_out/dist_check_compare.txt 2>&1
  python tools/dist_check.py compare 1 $W deferred_ >> gpurun_out/dist_check_compare.txt 2>&1
done
grep -E "^---|DIST CHECK|MISMATCH" gpurun_out/dist_check_compare.txt

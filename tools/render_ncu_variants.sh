#!/bin/bash
# Runs ON the GPU box: per compile-time variant of the render kernel, a short list of ncu metrics of one launch (c3 slice).
#   gpurun --timeout 1500 -- 'bash tools/render_ncu_variants.sh TAG "<XFLAGS 1>" "<XFLAGS 2>" ...'
set -u
TAG=${1:-n}; shift
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__occupancy_limit_warps,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__shared_mem_config_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed_op_shared_atom.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum
CMD="python bench.py --workload c3s --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
i=0
for XF in "$@"; do
  touch scene-aware-3d-multi-human_b200/csrc/mh_render.cu
  make -s -C scene-aware-3d-multi-human_b200/csrc XFLAGS="$XF" > gpurun_out/${TAG}_build$i.log 2>&1 || { echo "variant $i BUILD FAILED"; i=$((i+1)); continue; }
  ncu --metrics $M --clock-control none --kernel-name-base mangled -k regex:k_renderILi0 -s 3 -c 1 --csv --log-file gpurun_out/${TAG}_ncu$i.csv $CMD > gpurun_out/${TAG}_ncu$i.log 2>&1
  echo "== variant $i [$XF]"
  python - gpurun_out/${TAG}_ncu$i.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ni, ui, vi = h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value')
for r in rows[1:]:
    print(f'  {r[ni]:90s} {r[vi]:>16s} {r[ui]}')
PY
  i=$((i+1))
done

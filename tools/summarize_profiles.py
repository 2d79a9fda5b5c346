"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
G = os.path.join(ROOT, 'gpurun_out')
Pd = os.path.join(ROOT, 'profiles')
os.makedirs(Pd, exist_ok=True)

# ---- launch list: per-kernel totals over the LAST bench cycle pair (skip set-up launches: keep launches after the last k_rank_count-3..)
rows = []
with open(os.path.join(G, f'{tag}_launches.csv')) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
for r in rd:
    if len(r) == len(hdr):
        v = float(r[vi].replace(',', ''))
        v = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
        rows.append((r[ki].split('(')[0], v))
# the timed region = the last 2 cycles: from the second-to-last k_gather onwards
gidx = [i for i, (k, _) in enumerate(rows) if 'k_gather' in k]
ridx = [i for i, (k, _) in enumerate(rows) if 'k_render<0>' in k or 'k_renderILi0' in k]
if len(ridx) >= 2:
    # the last two fit cycles: from the k_gather before the second-to-last render launch to the RMSprop step after the last one
    # (bench.py times hot loop A afterwards: its launches are not part of the cycle)
    start = max(i for i in gidx if i < ridx[-2])
    end = next(i for i in range(ridx[-1], len(rows)) if 'k_rmsprop' in rows[i][0]) + 1
else:
    start, end = (gidx[-2] if len(gidx) >= 2 else 0), len(rows)
cyc = rows[start:end]
tot = collections.OrderedDict()
for k, v in cyc:
    tot[k] = tot.get(k, 0.0) + v
total = sum(tot.values())
with open(os.path.join(Pd, f'{tag}_launch_list.md'), 'w') as f:
    f.write(f'# Launch list of the last 2 fit cycles (`ncu --metrics gpu__time_duration.sum --clock-control none`, workload c3s = 8 persons x 64 frames x 1280x720)\n\n')
    f.write('Per-launch times under ncu are cold-cache and serialised: compare SHARES with `stage_ms` of the un-profiled run.\n\n')
    f.write(f'{len(cyc)} launches, {total / 1e3:.2f} ms total ({len(rows)} launches in the whole process incl. input synthesis)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n')
    cnt = collections.Counter(k for k, _ in cyc)
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f'| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / total:.2f} % |\n')
    bj = os.path.join(G, f'{tag}_bench_c3s.json')
    if os.path.exists(bj):
        d = json.loads(open(bj).read().strip().splitlines()[-1])
        f.write('\nUn-profiled run of the same command (CUDA events on the launch stream): `stage_ms` = ' + json.dumps(d['stage_ms']) + f", ms_per_step = {d['ms_per_step']:.3f}\n")
        sm = d['stage_ms']
        f.write(f"\nrender share: ncu {100 * tot.get('k_render<0>', tot.get('void k_render<0>', 0)) / total:.1f} % vs CUDA events {100 * sm['render'] / sum(sm.values()):.1f} %\n")

# ---- full capture of the render kernel
rep = os.path.join(G, f'{tag}_render.ncu-rep')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'sm__inst_executed_pipe_tma.sum', 'lts__t_sector_hit_rate.pct']
vals = {}
with open(os.path.join(Pd, f'{tag}_render_ncu.md'), 'w') as f:
    f.write('# `k_render<0>`: one launch, `ncu --set full --clock-control none --import-source on` (workload c3s: 512 person-frames at 1280x720)\n\n| metric | unit | value |\n|---|---|---|\n')
    for a, b, c in zip(h, u, v):
        if a in keep or ('issue_stalled' in a and a.endswith('per_issue_active.ratio')):
            f.write(f'| {a} | {b} | {c} |\n')
            vals[a] = (b, c)
    f.write('\nTop source lines by stall samples (`tools/ncu_lines.py`):\n\n```\n')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
    tmp = '/tmp/_src.csv'
    open(tmp, 'w').write(src)
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), tmp, '25'], capture_output=True, text=True).stdout)
    f.write('```\n')


def num(key):
    b, c = vals[key]
    x = float(c.replace(',', ''))
    return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(b, 1)


units = 512
traffic = num('dram__bytes_read.sum') + num('dram__bytes_write.sum')
json.dump({'workload': 'c3s', 'n_gpus': 1, 'dram_bytes_per_launch': traffic, 'person_frames_per_launch': units, 'dram_bytes_per_person_frame': traffic / units,
           'source': f'profiles/{tag}_render_ncu.md'}, open(os.path.join(Pd, 'render_traffic.json'), 'w'), indent=1)
print('wrote profiles for', tag, 'traffic per person-frame', traffic / units)


# ---- the tensor-core contractions (optional capture)
rep2 = os.path.join(G, f'{tag}_gemm_tc.ncu-rep')
if os.path.exists(rep2):
    raw = subprocess.run(['ncu', '-i', rep2, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u, rows2 = rr[0], rr[1], rr[2:]
    keep2 = ('Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
             'launch__shared_mem_per_block_dynamic', 'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
             'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
             'smsp__mem_tensor_reads_op_utcmma_matrix_c.sum.pct_of_peak_sustained_elapsed', 'smsp__mem_tensor_writes_op_utcmma.sum.pct_of_peak_sustained_elapsed',
             'smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
             'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
             'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')
    ki = h.index('Kernel Name')
    with open(os.path.join(Pd, f'{tag}_gemm_tc_ncu.md'), 'w') as f:
        f.write('# `k_gemm_fwd_tc` / `k_gemm_bwd_tc`: one launch each, `ncu --set full --clock-control none` (workload c3s: 528 bodies incl. halo slots, 1280x720)\n\n')
        f.write('Tensor-core contractions of `mh_gemm_tc.cu` (tcgen05.mma kind::tf32, 3 x TF32 split, accumulator in TMEM, basis by TMA bulk copies).\n\n')
        f.write('| metric | unit | ' + ' | '.join(r[ki].split('(')[0] for r in rows2) + ' |\n|---|---|' + '---|' * len(rows2) + '\n')
        for i, (k, uu) in enumerate(zip(h, u)):
            if k in keep2:
                f.write(f'| {k} | {uu} | ' + ' | '.join(r[i] for r in rows2) + ' |\n')
    print('wrote', f'{tag}_gemm_tc_ncu.md')

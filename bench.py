"""Benchmark of the hot path: person-frame optimizer-iters/sec of one fit() cycle (hot loop B, all terms on, frozen scene
cloud, filters present -- SURVEY.md section 8d) on the synthetic sequence BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM, CUDA-event timed, max over ranks.  `e2e`: the same cycles
through the public `SMPLDepthSequenceOptimizer.fit()` with HOST buffers (ingest H2D + result D2H inside the timed region).
`roofline`: the render kernel (dominant) against the measured HBM peak.  `cpu_baseline` / `--impl reference`: the CPU oracle
port of the reference optimiser (oracle/fit_ref.py; the reference itself needs PyTorch3D, which cannot be installed) on a
bounded sample of the same workload, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on; fits one GPU (~25 GB resident)
    'c3': dict(name='8 persons x 512 frames x 1280x720, 200k-pt scene cloud, batch 8', N=8, T=512, W=1280, H=720, M=200000, B=8),
    'c2': dict(name='3 persons x 200 frames x 512x512, batch 10', N=3, T=200, W=512, H=512, M=200000, B=10),
    'c4': dict(name='4 persons x 1000 frames x 1920x1080, batch 8', N=4, T=1000, W=1920, H=1080, M=200000, B=8),
    # BASELINE.json configs[4]: the scene-contact KNN sweep (512x512 planes as SURVEY.md 8d sizes it); 8 GPUs in BASELINE, fits one
    'c5': dict(name='6 persons x 2000 frames x 512x512, 1M-pt scene cloud, batch 8', N=6, T=2000, W=512, H=512, M=1000000, B=8),
    'c3s': dict(name='profiling slice of c3: 8 persons x 64 frames x 1280x720, batch 8', N=8, T=64, W=1280, H=720, M=200000, B=8),
    # accuracy leg (both arms, outside the timed region): small enough for 30 cycles of the CPU port
    'acc': dict(name='accuracy slice: 2 persons x 4 frames x 256x192, 20k-pt cloud, batch 2', N=2, T=4, W=256, H=192, M=20000, B=2),
    'small': dict(name='2 persons x 16 frames x 320x240 (plumbing)', N=2, T=16, W=320, H=240, M=20000, B=4),
}
COEFS = dict(proj2d_loss_coef=1.0, depth_loss_coef=0.05, silhouette_loss_coef=0.1, reg_velocity_coef=0.05,
             reg_verts_filter_coef=0.002, reg_poses_coef=0.002, reg_scales_coef=1e-4, reg_contact_coef=0.001,
             reg_foot_sliding_coef=0.01)          # configs/predict_mupots.yml:17-25
CPU_SAMPLE_T = 2                                   # frames of the CPU sample (all N persons, full resolution, full cloud)
ACC_CYCLES = 30                                    # cycles of the accuracy leg


def algorithmic_bytes_per_pf(w):
    """4 H W (1 + 1/N) [one f32 mask plane per person + the frame's f32 disparity plane shared by N persons]
    + 12 V [filtered-vertex target] + ~2.5 KB [parameters, references, gradients, optimiser state]."""
    return 4.0 * w['H'] * w['W'] * (1.0 + 1.0 / w['N']) + 12.0 * 6890 + 2560.0


# ------------------------------------------------------------------------------------------------- inputs
def model_dir():
    import synthdata
    d = os.path.join(tempfile.gettempdir(), 'mh_bench_model')
    if not os.path.exists(os.path.join(d, 'SMPL_NEUTRAL.pkl')):
        tmp = d + f'.{os.getpid()}'
        synthdata.write_model_dir(tmp, seed=0)
        try:
            os.rename(tmp, d)
        except OSError:
            pass
    return d


def project(K, P):
    x = P[..., 0] / P[..., 2]
    y = P[..., 1] / P[..., 2]
    return np.stack([K[0, 0] * x + K[0, 1] * y + K[0, 2], K[1, 0] * x + K[1, 1] * y + K[1, 2]], -1)


def build_problem(pkg, w, device, scene_update=False):
    """Optimiser with the synthetic sequence resident on the device.  The instance masks / disparity planes are rendered
    on the device from the ground-truth motion (hard z-buffer, nearest person wins); everything else is numpy."""
    import synthdata
    L = sys.modules[pkg.__name__ + '._lib']
    N, T, W, H, M, B = w['N'], w['T'], w['W'], w['H'], w['M'], w['B']
    cam_K = synthdata.camera_for(W, H)
    opt = pkg.SMPLDepthSequenceOptimizer(image_size=(W, H), num_frames=T, cam_K=cam_K, device=device,
                                         smpl_model_parameters_path=model_dir(), scene_update=scene_update, max_scene_points=M, **COEFS)
    opt._make_context(T, N, B)
    opt.batch_size = B
    opt.optim_scale_factor = True
    ctx, st = opt.ctx, opt._stream()
    sl = slice(opt.t0, opt.t1)
    Tl = opt.T_local
    motion = synthdata.make_motion(N, T, seed=1)
    # the noise is drawn for the WHOLE sequence and sliced: a frame gets the same inputs whatever rank owns it, so the losses of a
    # sharded run are comparable with the single-GPU run (`loss_check`)
    rng = np.random.default_rng(101)
    low_conf = rng.random((T, N, 17, 1)) < 0.05
    uv_noise = rng.normal(0, 0.3, (T, N, 17, 2))
    theta_noise = rng.normal(0, 0.05, (T, N, 72))
    # ground truth on the device -> planes
    ctx.set_param(L.P_XSCALE, np.zeros(N, np.float32), st)
    ctx.set_param(L.P_POSES_T, motion['trans'][sl], st)
    ctx.set_param(L.P_POSES_SMPL, motion['theta'][sl], st)
    ctx.set_param(L.P_BETAS, motion['beta'], st)
    ctx.call('mh_synth_planes', 1.0, 10.0, st)
    # 2D poses: projected ground-truth joints + noise, 5 % low-confidence joints
    _, j17 = opt.smpl_forward(np.tile(motion['beta'], (Tl, 1, 1)), motion['theta'][sl], want_verts=False)
    j3d = j17.reshape(Tl, N, 17, 3) + motion['trans'][sl][:, :, None, :]
    uv = project(cam_K, j3d)
    conf = np.where(low_conf[sl], 0.1, 0.9)
    pose2d = np.concatenate([uv + uv_noise[sl], conf], -1).astype(np.float32)
    theta_ref = (motion['theta'][sl] + theta_noise[sl]).astype(np.float32)
    theta_ref[..., 66:] = 0
    valid = np.ones((Tl, N), np.float32)
    ctx.call('mh_ingest_frames', 0, Tl, None, None, L.ptr(pose2d), L.ptr(theta_ref), L.ptr(valid), st)
    ctx.call('mh_finalize_ingest', st)
    opt.valid_smpl = np.ones((T, N, 1), np.float32)
    opt._ingested = True
    start = start_params(w, motion, opt.t0, opt.t1)
    set_start(opt, L, start)
    cloud = synthdata.scene_cloud(M, seed=2)
    opt.set_scene_pcd(cloud)
    return opt, dict(motion=motion, pose2d=pose2d, theta_ref=theta_ref, cam_K=cam_K, cloud=cloud, start=start)


def start_params(w, motion, t0, t1):
    """The state the timed cycles start from: ground truth perturbed like an early fit() cycle (same on every arm)."""
    rng = np.random.default_rng(7)
    N, T = w['N'], w['T']
    pT = (motion['trans'] + rng.normal(0, 0.03, (T, N, 3))).astype(np.float32)
    th = (motion['theta'] + rng.normal(0, 0.05, (T, N, 72))).astype(np.float32)
    th[..., 66:] = 0
    be = (motion['beta'] + rng.normal(0, 0.05, (1, N, 10))).astype(np.float32)
    max_z = np.clip(np.max(pT[..., 2], axis=1), 2, None).astype(np.float32)
    return dict(poses_T=pT[t0:t1], poses_smpl=th[t0:t1], betas=be, zmin_lin=np.ones_like(max_z)[t0:t1], zmax_lin=(2.0 * max_z)[t0:t1],
                xscale=np.zeros(N, np.float32))


def set_start(opt, L, s):
    ctx, st = opt.ctx, opt._stream()
    ctx.set_param(L.P_POSES_T, s['poses_T'], st); ctx.set_param(L.P_POSES_SMPL, s['poses_smpl'], st)
    ctx.set_param(L.P_BETAS, s['betas'], st); ctx.set_param(L.P_BETAS_REF, s['betas'], st)
    ctx.set_param(L.P_ZMIN_LIN, s['zmin_lin'], st); ctx.set_param(L.P_ZMAX_LIN, s['zmax_lin'], st)
    ctx.set_param(L.P_XSCALE, s['xscale'], st)
    ctx.call('mh_reset_optimizer', st)


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(',')]
            if len(p) < 6 or not (t_begin - 0.05 <= ts <= t_end + 0.15):
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], p[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_problem(w, Ts=None):
    """Bounded CPU sample of the workload: the first Ts (default CPU_SAMPLE_T) frames, all N persons, full resolution, full scene
    cloud.  The planes come from the oracle's own rasteriser so that this leg needs no GPU."""
    import torch
    import synthdata
    from oracle import fit_ref, raster, refmath as rm
    N, W, H, M = w['N'], w['W'], w['H'], w['M']
    Ts = CPU_SAMPLE_T if Ts is None else Ts
    md = model_dir()
    model = synthdata.load_model_tensors(md)
    cam_K = synthdata.camera_for(W, H)
    motion_full = synthdata.make_motion(N, w['T'], seed=1)
    motion = {k: (v[:Ts] if k != 'beta' else v) for k, v in motion_full.items()}
    mt = {a: (torch.from_numpy(v) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v) for a, v in model.items()}
    mt['parents'] = [int(p) for p in model['parents']]
    faces = torch.from_numpy(model['faces'].astype(np.int64))
    with torch.no_grad():
        out = rm.smpl_forward(mt, torch.from_numpy(np.tile(motion['beta'], (Ts, 1, 1)).reshape(-1, 10)), torch.from_numpy(motion['theta'].reshape(-1, 72)))
        verts = out['verts'].view(Ts, N, -1, 3) + torch.from_numpy(motion['trans']).view(Ts, N, 1, 3)
        j17 = rm.regress_joints(mt['J_regressor_alphapose'], verts.view(Ts * N, -1, 3))
        uv = rm.camera_projection(j17, torch.from_numpy(cam_K)[None].expand(Ts * N, 3, 3)).view(Ts, N, 17, 2).numpy()
        Kndc = torch.from_numpy(rm.compute_calibration_matrix(1.0, 100.0, cam_K, (W, H)))
        zb = np.zeros((Ts, N, H, W), np.float32)
        for t in range(Ts):
            for n in range(N):
                zb[t, n] = raster.rasterize(raster.world_to_ndc(verts[t, n], Kndc), faces, H, W, 0.0, 1)['zbuf'][..., 0].numpy()
    data = synthdata.assemble_inputs(zb, motion, uv, cam_K, W, H, seed=1)
    coefs = dict(proj2d=1.0, depth=0.05, silhouette=0.1, reg_velocity=0.05, reg_verts_filter=0.002, reg_poses=0.002,
                 reg_scales=1e-4, reg_contact=0.001, reg_foot_sliding=0.01)
    fr = fit_ref.FitRef(model, (W, H), Ts, cam_K, coefs)
    s = start_params(w, motion_full, 0, Ts)
    fr.set_variables(s['poses_T'].reshape(Ts, N, 1, 3), s['poses_smpl'], s['betas'], data['valid_smpl'], s['zmin_lin'].reshape(Ts, 1, 1),
                     s['zmax_lin'].reshape(Ts, 1, 1), np.zeros((1, N, 1, 1), np.float32))
    fr.set_scene_pcd(synthdata.scene_cloud(M, seed=2))
    fr.refresh_filters()
    opt = torch.optim.RMSprop(fr.leaves(), lr=0.01, alpha=0.5, momentum=0.9)
    B = min(w['B'], Ts)
    batches = [np.arange(b, min(b + B, Ts)) for b in range(0, Ts, B)]

    def step(lr=0.01):
        for grp in opt.param_groups:
            grp['lr'] = lr
        fr.cycle_grads(data, batches)
        opt.step()
    step.fit_ref, step.data, step.batches, step.cam_K, step.start = fr, data, batches, cam_K, s      # for tests/test_gpu_fullsize.py
    step.motion, step.cloud = motion, fr.scene_pcd[0, 0].numpy()
    return step, Ts * N


def run_cpu(w, steps, warmup):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, pf = cpu_problem(w)
    for _ in range(warmup):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = time.time() - t0
    return pf * steps / dt, dt / steps, cores, pf


def reference_arm(args, w, rank):
    if rank != 0:
        return
    value, sec, cores, pf = run_cpu(w, args.steps, args.warmup)
    sample = f'{CPU_SAMPLE_T} frames x {w["N"]} persons x {w["W"]}x{w["H"]}, {w["M"]}-pt cloud, all terms; one cycle per step'
    print(json.dumps({
        'impl': 'reference', 'metric': 'person_frame_optimizer_iters_per_sec', 'value': value, 'unit': 'person-frame-iters/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': w['name']},
        'cpu_baseline': {'value': value, 'unit': 'person-frame-iters/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'person-frame-iters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'CPU oracle port of mhmocap/optimizer.py (torch autograd, CPU rasteriser restatement): the unmodified reference '
                'imports PyTorch3D, which is neither vendored nor installable offline',
    }), flush=True)


# ------------------------------------------------------------------------------------------------- accuracy
def run_accuracy(pkg, device='cuda:0', cycles=ACC_CYCLES, wname='acc'):
    """"MPJPE vs ref" half of the BASELINE metric, outside every timed region: ACC_CYCLES steady-state cycles (all terms, frozen
    cloud, filters refreshed once at the start) from IDENTICAL starts and IDENTICAL host inputs on both arms -- the CUDA path and
    the CPU port of the reference optimiser -- then the reference's own evaluation (evaluate.py:180-296, eval_mupots.py:18-42:
    MPJPE over the first 14 MuPoTS joints, absolute, in mm) of each arm against the synthetic ground truth and of one against the
    other.  Part of the CPU-baseline leg (the only place bench.py may execute oracle/)."""
    import torch
    L = sys.modules[pkg.__name__ + '._lib']
    ev = sys.modules.get(pkg.__name__ + '.evaluation')
    if ev is None:
        import importlib
        ev = importlib.import_module(pkg.__name__ + '.evaluation')
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import gpu_harness as gh
    w = WORKLOADS[wname]
    N, T, W, H, M, B = w['N'], w['T'], w['W'], w['H'], w['M'], w['B']
    torch.set_num_threads(os.cpu_count() or 1)
    step, _ = cpu_problem(w, Ts=T)
    fr, data, cam_K, start, motion = step.fit_ref, step.data, step.cam_K, step.start, step.motion
    # ---- CUDA arm (first: the port's leaves still hold the start state, the filters are refreshed from it on both arms)
    opt = pkg.SMPLDepthSequenceOptimizer(image_size=(W, H), num_frames=T, cam_K=cam_K, device=device, smpl_model_parameters_path=model_dir(),
                                         scene_update=False, max_scene_points=M, **COEFS)
    opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=0, batch_size=B)
    opt._ingest(gh.ListLoader(data, B))
    set_start(opt, L, start)
    opt.set_scene_pcd(step.cloud)
    opt._refresh_filters(0.01, 0.02, 0.001, 0.5)
    lr = 0.01
    for _ in range(cycles):
        opt.step_device_only(lr); lr *= 0.99
    ours = opt.get_optimized_variables()
    # ---- CPU port
    t0 = time.time()
    lr = 0.01
    for _ in range(cycles):
        step(lr); lr *= 0.99
    port_s = time.time() - t0
    port = {'poses_T': fr.poses_T.detach().numpy(), 'poses_smpl': fr.poses_smpl.detach().numpy(), 'betas_smpl': fr.betas.detach().numpy(),
            'scale_factor': np.power(np.float32(1.1), fr.xscale.detach().numpy()).astype(np.float32)}
    first = {'poses_T': start['poses_T'].reshape(T, N, 1, 3), 'poses_smpl': start['poses_smpl'], 'betas_smpl': start['betas'],
             'scale_factor': np.ones((1, N, 1, 1), np.float32)}
    # ---- evaluation (SMPL + MuPoTS joint regression on the device, matching / averages on the host)
    sj = ev.SMPLJoints(opt, {'mupots': np.load(os.path.join(model_dir(), 'SMPL_MuPoTs_Regressor_v1.npy'))})
    vis = np.ones((T, N, 17, 1), np.float32)

    def joints_of(v):
        be = np.tile(np.asarray(v['betas_smpl']).reshape(1, N, 10), (T, 1, 1))
        j = sj(be.reshape(-1, 10), np.asarray(v['poses_smpl']).reshape(-1, 72), 'mupots').reshape(T, N, 17, 3)
        sc = np.asarray(v['scale_factor']).reshape(1, N, 1, 1)
        return sc * j + np.asarray(v['poses_T']).reshape(T, N, 1, 3)

    def mpjpe(v, ref):
        vv = dict(v)
        vv['betas_smpl'] = np.tile(np.asarray(v['betas_smpl']).reshape(1, N, 10), (T, 1, 1))
        return float(ev.compute_mm_pck_results(vv, ref, vis, sj, cam_K)['mm_abs_error'])

    gt = joints_of({'poses_T': motion['trans'][:T], 'poses_smpl': motion['theta'][:T], 'betas_smpl': motion['beta'],
                    'scale_factor': np.ones((1, N, 1, 1), np.float32)})
    res = {'cycles': cycles, 'slice': w['name'], 'metric': 'MPJPE over the first 14 MuPoTS joints, absolute, mm (eval_mupots.py:18-42)',
           'start_vs_gt_mm': mpjpe(first, gt),
           'ours': {'mpjpe_vs_gt_mm': mpjpe(ours, gt), 'mpjpe_vs_port_mm': mpjpe(ours, joints_of(port))},
           'port': {'mpjpe_vs_gt_mm': mpjpe(port, gt), 'seconds': port_s},
           'max_abs_diff_ours_vs_port': {'poses_T_m': float(np.abs(np.asarray(ours['poses_T']) - port['poses_T']).max()),
                                         'poses_smpl_rad': float(np.abs(np.asarray(ours['poses_smpl']) - port['poses_smpl']).max()),
                                         'betas': float(np.abs(np.asarray(ours['betas_smpl']) - port['betas_smpl']).max())}}
    opt.ctx.close()
    return res


# ------------------------------------------------------------------------------------------------- hot loop A
def run_init_loop(opt, aux, w, L, iters=50, warmup=5):
    """Hot loop A (optimizer.py:740-761, Adam on the translations) on the bench workload, CUDA-event timed: person-frame
    iterations per second of `mh_init_grads` + exchange + `mh_init_update` after the one-time SMPL evaluation of `mh_init_begin`."""
    import torch
    import torch.distributed as dist
    ctx, st = opt.ctx, opt._stream()
    Tl, N = opt.T_local, w['N']
    betas = np.ascontiguousarray(np.tile(aux['motion']['beta'].reshape(1, N, 10), (Tl, 1, 1)), np.float32)
    ctx.call('mh_init_begin', L.ptr(aux['pose2d']), L.ptr(aux['theta_ref']), L.ptr(betas), 0.15, st)
    sh = sys.modules[type(opt).__module__.rsplit('.', 1)[0] + '.sharding']

    def it(k, lr):                                   # as SMPLDepthSequenceOptimizer.__init_global_poses, without the loss readback
        if opt._lib_comm or not opt._dist:
            ctx.call('mh_init_cycle', lr, k + 1, st)
            return
        hp, hn = opt._exchange_halo()
        ctx.call('mh_init_grads', hp, hn, st)
        sh.allreduce_shared(opt._view(L.BUF_SHARED), opt.group)
        ctx.call('mh_init_update', lr, k + 1, st)

    lr = 0.5
    for k in range(warmup):
        it(k, lr); lr *= 0.95
    torch.cuda.synchronize()
    if opt._dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(warmup, warmup + iters):
        it(k, lr); lr *= 0.95
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if opt._dist:
        tt = torch.tensor([ms], device=opt.device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return {'value': w['N'] * w['T'] * iters / (ms * 1e-3), 'unit': 'person-frame-iters/s', 'ms_per_iter': ms / iters, 'iters': iters,
            'what': 'hot loop A: 2-D reprojection + velocity gradients of the translations, fused Adam step (SMPL joints cached once)'}


# ------------------------------------------------------------------------------------------------- e2e
class HostLoader(object):
    """Re-iterable over pinned HOST tensors with the keys of the reference dataset (datautils.py:630-641)."""
    def __init__(self, arrays, B):
        self.a, self.B = arrays, B
        self.T = arrays['idxs'].shape[0]

    def __iter__(self):
        for s in range(0, self.T, self.B):
            yield {k: v[s:s + self.B] for k, v in self.a.items()}


def run_e2e(pkg, w, device, opt_dev, aux, steps):
    """The same cycles through the public API from HOST buffers: a fresh optimiser ingests the whole (local) sequence from
    pinned host memory, runs `steps` cycles of fit() and reads the optimised variables back -- all inside the timed region."""
    import torch
    import torch.distributed as dist
    L = sys.modules[pkg.__name__ + '._lib']
    N, T, W, H, M, B = w['N'], w['T'], w['W'], w['H'], w['M'], w['B']
    Tl, t0 = opt_dev.T_local, opt_dev.t0
    depths = np.empty((Tl, H, W), np.float32)
    seg = np.empty((Tl, N, H, W), np.float32)
    chunk = 8
    for s in range(0, Tl, chunk):
        c = min(chunk, Tl - s)
        opt_dev.ctx.call('mh_read_planes', s, c, L.ptr(depths[s:s + c]), L.ptr(seg[s:s + c]))
    # every rank iterates the whole loader in the reference API; here each rank's loader holds its own frames and the
    # other frames are marked as seen through a tiny index-only pass
    arrays = {'depths': torch.from_numpy(depths).pin_memory(), 'seg_mask': torch.from_numpy(seg).pin_memory(),
              'pose2d': torch.from_numpy(aux['pose2d']).pin_memory(), 'poses_smpl': torch.from_numpy(aux['theta_ref']).pin_memory(),
              'idxs': torch.arange(t0, t0 + Tl, dtype=torch.int64)}
    del depths, seg
    # whole-sequence arrays of init_optimized_variables (optimizer.py:262): every rank passes the full shapes, only its own frames
    # are read (the 2-D poses of the other ranks' frames are left zero here instead of being recomputed on every rank)
    pose2d = np.zeros((T, N, 17, 3), np.float32); pose2d[t0:t0 + Tl] = aux['pose2d']
    theta_ref = np.zeros((T, N, 72), np.float32); theta_ref[t0:t0 + Tl] = aux['theta_ref']
    betas = np.ascontiguousarray(np.tile(aux['motion']['beta'].reshape(1, N, 10), (T, 1, 1)), np.float32)
    valid = np.ones((T, N, 1), np.float32)
    loader = HostLoader(arrays, B)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    # the device-timed optimiser is done: its buffers go back to the library's pool, as between two sequences of one job
    # (predict.py:315-357 builds one optimiser per sequence); the fresh optimiser below is built from them
    opt_dev.ctx.close()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    t_begin = time.perf_counter()
    # the calls a user of the reference makes (predict.py:290-344): construct, init (hot loop A, 100 iterations), fit, read back.
    # `set_scene_pcd` freezes the scene cloud and `start_cycle=50` resumes the schedule in the steady-state regime (filters
    # refreshed at the first cycle, every term on) -- the regime `value` is measured in
    opt = pkg.SMPLDepthSequenceOptimizer(image_size=(W, H), num_frames=T, cam_K=aux['cam_K'], device=device,
                                         smpl_model_parameters_path=model_dir(), scene_update=False, max_scene_points=M,
                                         allow_partial_loader=True, **COEFS)
    t_ctor = time.perf_counter()
    opt.init_optimized_variables(pose2d, theta_ref, betas, valid, num_iter=100, batch_size=B)
    t_iv = time.perf_counter()
    opt.set_scene_pcd(aux['cloud'])
    t_init = time.perf_counter()
    log = opt.fit(loader, num_iter=50 + steps, start_cycle=50)
    t_fit = time.perf_counter()
    out = opt.get_optimized_variables()
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t_begin
    if world > 1:
        tt = torch.tensor([dt], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    print(f'e2e phases (rank {opt.rank}): construct {t_ctor - t_begin:.3f} s + init {t_iv - t_ctor:.3f} s + scene cloud {t_init - t_iv:.3f} s | fit (ingest {getattr(opt, "ingest_seconds", 0.0):.3f} s + '
          f'{steps} cycles) {t_fit - t_init:.3f} s | read back {time.perf_counter() - t_fit:.3f} s', file=sys.stderr)
    d2h = sum(v.nbytes for v in out.values() if isinstance(v, np.ndarray)) + (steps + 100) * 16 * 4
    h2d = opt.h2d_bytes + pose2d.nbytes + theta_ref.nbytes + betas.nbytes + aux['cloud'].nbytes
    opt.ctx.close()
    return N * T * steps / dt, h2d / steps, d2h / steps, float(log[-1]['loss_silhouette'])


# ------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--render-profile', action='store_true', help='development: per-phase cycle counters of the render kernel')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        reference_arm(args, w, rank)
        return
    assert args.warmup >= 3, 'timing rules: at least 3 warm-up steps'
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    L = sys.modules[pkg.__name__ + '._lib']

    opt, aux = build_problem(pkg, w, device)
    opt._refresh_filters(0.01, 0.02, 0.001, 0.5)
    ctx = opt.ctx
    lr = 0.01
    for _ in range(args.warmup):
        opt.step_device_only(lr); lr *= 0.99
    losses0 = ctx.read_losses(opt._stream())
    ctx.call('mh_set_timing', 1)
    if args.render_profile:
        ctx.call('mh_render_profile', 1, None)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    ev0.record()
    for _ in range(args.steps):
        opt.step_device_only(lr); lr *= 0.99
    ev1.record()
    torch.cuda.synchronize(device)
    t_end = time.time()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches() - launches0
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    if world > 1:
        tt = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    if args.render_profile:
        prof = np.zeros(32, np.int64)
        ctx.call('mh_render_profile', 0, L.ptr(prof))
        if prof[8:].any():
            pf = float(opt.T_local * w['N'] * args.steps)
            sn = ['items', 'slots', 'warp-passes', 'lanes past prune', 'warp-passes past prune', 'lanes past key test', 'warp-passes past key test',
                  'depth atomics', 'sil inserts', 'not-inside distance evals', 'items without a survivor', '-']
            print('pair statistics per person-frame:', {n_: round(float(v) / pf, 1) for n_, v in zip(sn, prof[8:20])}, file=sys.stderr)
        prof = prof[:8]
        names = ['load+ndc', 'binning', 'staging', 'pairs', 'per-pixel', 'sums+depth-bwd', 'chain', '-']
        print('render phases (% of CTA cycles):', {n_: round(100.0 * float(v) / max(float(prof.sum()), 1.0), 1) for n_, v in zip(names, prof)}, file=sys.stderr)
    stage = ctx.read_timing(min(args.steps, 64))
    ctx.call('mh_set_timing', 0)
    losses1 = ctx.read_losses(opt._stream())
    value = w['N'] * w['T'] * args.steps / (ms * 1e-3)
    render_ms = float(stage[:, 3].mean())
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    units = opt.T_local * w['N']
    achieved = algorithmic_bytes_per_pf(w) * units / (render_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'render_traffic.json')
    if os.path.exists(tp):
        # dram__bytes_read + dram__bytes_write of ONE k_render<0> launch from the committed `ncu --set full` capture (workload c3s:
        # the first 64 frames of c3, same persons / resolution / cloud), scaled by the person-frames of this launch
        tj = json.load(open(tp))
        if args.workload in ('c3', 'c3s'):
            traffic = tj['dram_bytes_per_person_frame'] * units
    line = {
        'metric': 'person_frame_optimizer_iters_per_sec', 'value': value, 'unit': 'person-frame-iters/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': w['name'], 'persons': w['N'], 'frames': w['T'], 'image': [w['W'], w['H']], 'scene_points': w['M'],
                   'batch': w['B'], 'terms': 'proj2d+depth+silhouette+contact+foot+priors+velocity+filtered-verts, RMSprop',
                   'frames_per_gpu': opt.T_local, 'l2': 'inputs larger than L2 (>= 1.7 GB of per-cycle vertex arrays + planes per GPU)'},
        'gpu_launches': int(launches),
        'stage_ms': {k: float(stage[:, i].mean()) for i, k in enumerate(['smpl_forward', 'terms', 'order_prepass', 'render', 'smpl_backward', 'post'])},
        'roofline': {'bound': 'hbm', 'kernel': 'k_render<0>', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': algorithmic_bytes_per_pf(w) * units,
                     'launch_ms': render_ms},
        'clocks': clocks,
        'loss_check': {'silhouette_first': float(losses0[L.L_SILHOUETTE]), 'silhouette_last': float(losses1[L.L_SILHOUETTE]),
                       'pose2d_first': float(losses0[L.L_POSE2D]), 'pose2d_last': float(losses1[L.L_POSE2D])},
    }
    # sharded == single: the first-cycle losses of this run against the committed single-GPU values of the same workload
    ref_path = os.path.join(ROOT, 'profiles', 'loss_check_n1.json')
    if os.path.exists(ref_path):
        n1 = json.load(open(ref_path)).get(args.workload)
        if n1:
            rel = max(abs(line['loss_check'][k] - n1[k]) / max(abs(n1[k]), 1e-12) for k in ('silhouette_first', 'pose2d_first'))
            line['loss_check']['n1_reference'] = {k: n1[k] for k in ('silhouette_first', 'pose2d_first')}
            line['loss_check']['max_rel_diff_vs_n1'] = rel
            if rel > 1e-3:
                raise SystemExit(f'loss_check: the {world}-GPU losses differ from the single-GPU reference by {rel:.2e} (> 1e-3): {line["loss_check"]}')
    line['init_loop'] = run_init_loop(opt, aux, w, L)
    if not args.no_e2e:
        e2e, h2d, d2h, _ = run_e2e(pkg, w, device, opt, aux, args.steps)           # closes `opt`
        line['e2e'] = {'value': e2e, 'unit': 'person-frame-iters/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                       'what': f'public API from pinned host buffers (reference dtypes): init_optimized_variables (100 iterations) + fit() of {args.steps} '
                               f'steady-state cycles incl. one-time ingest, filter refresh, per-cycle loss readback and get_optimized_variables(); device buffers recycled from the '
                               f'library pool (second optimiser of the process)'}
    opt.ctx.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores, pf = run_cpu(w, 1, 1)
        line['cpu_baseline'] = {'value': v, 'unit': 'person-frame-iters/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{CPU_SAMPLE_T} frames x {w["N"]} persons x {w["W"]}x{w["H"]}, {w["M"]}-pt cloud, all terms; 1 warm + 1 timed cycle ({sec:.1f} s)'}
        line['accuracy'] = run_accuracy(pkg, device)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

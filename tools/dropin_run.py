"""A whole drop-in run on a MuPoTS-shaped synthetic sequence (BASELINE config C2: 3 persons x 200 frames x 512x512, batch 10):
init_optimized_variables (100 Adam iterations) + fit (scene updates from cycle 30, filter refreshes every 25 cycles) from
HOST buffers through the public API, with wall-clock per phase."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.argv = [sys.argv[0]] + sys.argv[1:]
import torch
import bench
import __graft_entry__ as ge

pkg = ge.load_package()
L = sys.modules[pkg.__name__ + '._lib']
w = dict(bench.WORKLOADS[os.environ.get('WORKLOAD', 'c2')])
num_iter = int(os.environ.get('NUM_ITER', '100'))
device = torch.device('cuda', 0)
torch.cuda.set_device(device)
dev_opt, aux = bench.build_problem(pkg, w, device)
N, T, W, H, B = w['N'], w['T'], w['W'], w['H'], w['B']
depths = np.empty((T, H, W), np.float32); seg = np.empty((T, N, H, W), np.float32)
for s in range(0, T, 8):
    c = min(8, T - s)
    dev_opt.ctx.call('mh_read_planes', s, c, L.ptr(depths[s:s + c]), L.ptr(seg[s:s + c]))
dev_opt.ctx.close()
rng = np.random.default_rng(0)
arrays = {'depths': torch.from_numpy(depths), 'seg_mask': torch.from_numpy(seg), 'pose2d': torch.from_numpy(aux['pose2d']),
          'poses_smpl': torch.from_numpy(aux['theta_ref']), 'idxs': torch.arange(T, dtype=torch.int64),
          'images': torch.from_numpy(rng.integers(0, 255, (T, H, W, 3), dtype=np.uint8)),
          'backmasks': torch.from_numpy((seg.sum(1) == 0).astype(np.uint8))}
betas = np.tile(aux['motion']['beta'], (T, 1, 1)) + rng.normal(0, 0.05, (T, N, 10)).astype(np.float32)
opt = pkg.SMPLDepthSequenceOptimizer(image_size=(W, H), num_frames=T, cam_K=aux['cam_K'], device=device,
                                     smpl_model_parameters_path=bench.model_dir(), **bench.COEFS)
t0 = time.perf_counter()
init_log = opt.init_optimized_variables(aux['pose2d'], aux['theta_ref'], betas.astype(np.float32), np.ones((T, N, 1), np.float32), num_iter=100, batch_size=B)
torch.cuda.synchronize()
t1 = time.perf_counter()
loader = bench.HostLoader(arrays, B)
opt._ingest(loader)
torch.cuda.synchronize()
t2 = time.perf_counter()
# fit in two segments so that the cycles with the per-cycle scene update (>= 30: device median -> device post-processing -> device
# point cloud) are timed apart from the cycles without it
ta = time.perf_counter()
log = opt.fit(loader, num_iter=30)
torch.cuda.synchronize()
tb = time.perf_counter()
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
log += opt.fit(loader, num_iter=num_iter, start_cycle=30)
torch.cuda.synchronize()
pr.disable()
t3 = time.perf_counter()
print(f'cycles 0-29 (no scene update): {(tb - ta) / 30 * 1e3:.2f} ms/cycle | cycles 30-{num_iter - 1} (scene update every cycle, filter refresh '
      f'every 25): {(t3 - tb) / max(num_iter - 30, 1) * 1e3:.2f} ms/cycle')
v = opt.get_optimized_variables()
gt = aux['motion']['trans']
print(f'{w["name"]}: init (100 it) {t1 - t0:.2f} s | ingest {t2 - t1:.2f} s | fit ({num_iter} cycles) {t3 - t2:.2f} s = {(t3 - t2) / num_iter * 1e3:.1f} ms/cycle '
      f'| person-frame-iters/s over fit {N * T * num_iter / (t3 - t2):.0f}')
print('init loss_2d', float(init_log[0]['loss_2d']), '->', float(init_log[-1]['loss_2d']), '| translation error after init (m, mean)',
      float(np.abs(opt.ctx.get_param(L.P_POSES_T, (T, N, 3)) - gt).mean()))
print('fit log first/last:', {k: (round(log[0][k], 5), round(log[-1][k], 5)) for k in log[0]})
print('scene points', v['scene_depth'].shape if v['scene_depth'] is not None else None, 'translation error after fit (m, mean)', float(np.abs(v['poses_T'][:, :, 0] - gt).mean()))
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)

# the per-cycle scene update taken apart (CUDA events on the launch stream, mean of 5)
st = opt._stream()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = np.zeros(3)
for _ in range(5):
    evs[0].record()
    opt._median_passes(0)
    evs[1].record()
    opt.ctx.call('mh_scene_update_from_median', 1, 7, None, st)
    evs[2].record()
    opt.ctx.call('mh_fit_cycle_grads', st)
    evs[3].record()
    torch.cuda.synchronize()
    acc += [evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3])]
acc /= 5
print(f'scene update: temporal median (10 radix passes over {T} frames) {acc[0]:.2f} ms | post-processing + point cloud + contact grids {acc[1]:.2f} ms '
      f'| gradients of the cycle {acc[2]:.2f} ms')

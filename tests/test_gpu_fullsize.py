"""Parity at the image sizes of the BASELINE configurations: a teacher-forced cycle with every term on, CUDA path vs CPU oracle on
identical inputs and parameters; all nine losses and all six gradient tensors must agree.
  c3       2 frames x 8 persons x 1280x720, 200 000-point cloud (the shape bench.py's CPU baseline times): hundreds of raster tiles per
           body, persons from 3 m to 10 m, occlusion order over 8 persons, top-32 selection over 200 k points;
  c2       4 frames x 3 persons x 512x512 in batches of 2 (MuPoTS-shaped): the SQUARE-image branch of the NDC convention, per-batch
           weights of the priors over two batches;
  c4       2 frames x 4 persons x 1920x1080: 2040 tiles of 32x32 pixels in the image, larger than the tile-bin table of a body;
  portrait 2 frames x 2 persons x 480x640: the H > W branch of the NDC convention.
Needs a B200."""
import os
import sys

import numpy as np
import pytest

import gpu_harness as gh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


SHAPES = {
    'c3': dict(name='c3-shaped', N=8, T=2, W=1280, H=720, M=200000, B=2),
    'c2': dict(name='c2-shaped', N=3, T=4, W=512, H=512, M=200000, B=2),
    'c4': dict(name='c4-shaped', N=4, T=2, W=1920, H=1080, M=200000, B=2),
    'portrait': dict(name='portrait', N=2, T=2, W=480, H=640, M=50000, B=2),
}


@pytest.mark.parametrize('shape', sorted(SHAPES))
def test_baseline_shaped_cycle_matches_the_oracle(shape):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    sys.argv = ['bench.py']
    sys.path.insert(0, ROOT)
    import bench
    import __graft_entry__ as ge
    pkg = ge.load_package()
    L = sys.modules[pkg.__name__ + '._lib']
    sh = sys.modules[pkg.__name__ + '.sharding']
    w = SHAPES[shape]
    torch.set_num_threads(os.cpu_count() or 1)
    N, W, H, M, T, B = w['N'], w['W'], w['H'], w['M'], w['T'], w['B']
    step, pf = bench.cpu_problem(w, Ts=T)
    fr, data, cam_K, start = step.fit_ref, step.data, step.cam_K, step.start
    olog, _ = fr.cycle_grads(data, step.batches)
    ograds = {nm: p.grad.numpy().copy() for nm, p in zip(gh.NAMES, fr.leaves())}

    opt = pkg.SMPLDepthSequenceOptimizer(image_size=(W, H), num_frames=T, cam_K=cam_K, device='cuda:0', smpl_model_parameters_path=bench.model_dir(),
                                         scene_update=False, max_scene_points=M, **bench.COEFS)
    opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=0, batch_size=B)
    opt._ingest(gh.ListLoader(data, B))
    ctx, st = opt.ctx, opt._stream()
    ctx.set_param(L.P_POSES_T, start['poses_T'], st); ctx.set_param(L.P_POSES_SMPL, start['poses_smpl'], st)
    ctx.set_param(L.P_BETAS, start['betas'], st); ctx.set_param(L.P_BETAS_REF, start['betas'], st)
    ctx.set_param(L.P_ZMIN_LIN, start['zmin_lin'], st); ctx.set_param(L.P_ZMAX_LIN, start['zmax_lin'], st)
    ctx.set_param(L.P_XSCALE, start['xscale'], st)
    opt.set_scene_pcd(fr.scene_pcd[0, 0].numpy())
    gh.set_filtered(opt, fr.verts_filtered.numpy())
    ctx.call('mh_fit_grads', 0, 0, st)
    log = sh.log_from_loss_block(ctx.read_losses(st), len(step.batches))
    for k, v in log.items():
        assert abs(v - olog[k]) <= 1e-4 * abs(olog[k]) + 1e-9, (shape, k, v, olog[k])
    assert log['loss_depth'] > 0 and log['loss_silhouette'] > 0 and log['reg_contact'] > 0            # the terms are really on
    grads = {'poses_T': ctx.get_grad(L.P_POSES_T, (T, N, 1, 3)), 'poses_smpl': ctx.get_grad(L.P_POSES_SMPL, (T, N, 72)),
             'betas': ctx.get_grad(L.P_BETAS, (1, N, 10)), 'zmin_lin': ctx.get_grad(L.P_ZMIN_LIN, (T, 1, 1)),
             'zmax_lin': ctx.get_grad(L.P_ZMAX_LIN, (T, 1, 1)), 'xscale': ctx.get_grad(L.P_XSCALE, (1, N, 1, 1))}
    for nm, gr in grads.items():
        ref = ograds[nm].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, (shape, nm, np.abs(gr - ref).max(), np.abs(ref).max())
    opt.ctx.close()

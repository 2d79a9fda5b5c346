"""Parity at the benchmark's full image size: 2 frames x 8 persons x 1280x720 with the 200 000-point scene cloud and every
term on (the shape bench.py's CPU baseline times).  The CUDA path and the CPU oracle get identical inputs and parameters; all
nine losses and all six gradient tensors must agree.  Exercises what the small golden cases cannot: hundreds of raster tiles
per body, persons from 3 m to 10 m, occlusion order over 8 persons, top-32 selection over 200 k points.  Needs a B200."""
import os
import sys

import numpy as np
import pytest

import gpu_harness as gh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c3_shaped_cycle_matches_the_oracle():
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
    w = bench.WORKLOADS['c3']
    torch.set_num_threads(os.cpu_count() or 1)
    step, pf = bench.cpu_problem(w)
    fr, data, cam_K, start = step.fit_ref, step.data, step.cam_K, step.start
    N, W, H, M, T = w['N'], w['W'], w['H'], w['M'], bench.CPU_SAMPLE_T
    olog, _ = fr.cycle_grads(data, step.batches)
    ograds = {nm: p.grad.numpy().copy() for nm, p in zip(gh.NAMES, fr.leaves())}

    opt = pkg.SMPLDepthSequenceOptimizer(image_size=(W, H), num_frames=T, cam_K=cam_K, device='cuda:0', smpl_model_parameters_path=bench.model_dir(),
                                         scene_update=False, max_scene_points=M, **bench.COEFS)
    opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=0, batch_size=T)
    opt._ingest(gh.ListLoader(data, T))
    ctx, st = opt.ctx, opt._stream()
    ctx.set_param(L.P_POSES_T, start['poses_T'], st); ctx.set_param(L.P_POSES_SMPL, start['poses_smpl'], st)
    ctx.set_param(L.P_BETAS, start['betas'], st); ctx.set_param(L.P_BETAS_REF, start['betas'], st)
    ctx.set_param(L.P_ZMIN_LIN, start['zmin_lin'], st); ctx.set_param(L.P_ZMAX_LIN, start['zmax_lin'], st)
    ctx.set_param(L.P_XSCALE, start['xscale'], st)
    opt.set_scene_pcd(fr.scene_pcd[0, 0].numpy())
    gh.set_filtered(opt, fr.verts_filtered.numpy())
    ctx.call('mh_fit_grads', 0, 0, st)
    log = sh.log_from_loss_block(ctx.read_losses(st), 1)
    for k, v in log.items():
        assert abs(v - olog[k]) <= 1e-4 * abs(olog[k]) + 1e-9, (k, v, olog[k])
    grads = {'poses_T': ctx.get_grad(L.P_POSES_T, (T, N, 1, 3)), 'poses_smpl': ctx.get_grad(L.P_POSES_SMPL, (T, N, 72)),
             'betas': ctx.get_grad(L.P_BETAS, (1, N, 10)), 'zmin_lin': ctx.get_grad(L.P_ZMIN_LIN, (T, 1, 1)),
             'zmax_lin': ctx.get_grad(L.P_ZMAX_LIN, (T, 1, 1)), 'xscale': ctx.get_grad(L.P_XSCALE, (1, N, 1, 1))}
    for nm, gr in grads.items():
        ref = ograds[nm].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, (nm, np.abs(gr - ref).max(), np.abs(ref).max())
    opt.ctx.close()

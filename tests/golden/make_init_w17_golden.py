"""Golden vector of hot loop A with NON-UNIFORM ``pose17j_weights`` (``optimizer.py:108-130, 754-756``): the UNMODIFIED reference's
``init_optimized_variables`` on the inputs of ``fit_n2.npz``.  Build container only (needs ``/root/reference``).

Usage:  python tests/golden/make_init_w17_golden.py        -> tests/golden/init_w17.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets sys.path for the reference and the PyTorch3D stand-in)

W17 = [3.0, 1.0, 1.0, 0.5, 0.5, 2.0, 2.0, 1.0, 1.0, 0.25, 0.25, 2.5, 2.5, 1.0, 1.0, 4.0, 4.0]

if __name__ == '__main__':
    assert os.path.isdir(mg.REF), 'needs the reference mounted at /root/reference'
    mg.synth.write_model_dir(mg.MODEL_DIR, seed=0)
    ropt = mg.import_reference()[0]
    g = np.load(os.path.join(HERE, 'fit_n2.npz'))
    N, T, W, H, batch, num_iter, init_iter = [int(v) for v in g['meta_NTWH_batch']]
    opt = ropt.SMPLDepthSequenceOptimizer(image_size=(W, H), num_frames=T, cam_K=g['cam_K'], device='cpu',
                                          smpl_model_parameters_path=mg.MODEL_DIR, pose17j_weights=W17,
                                          proj2d_loss_coef=1.0, reg_velocity_coef=0.05)          # as make_golden.run_fit
    torch.manual_seed(0)
    log = opt.init_optimized_variables(g['in_pose2d'], g['in_poses_smpl'], g['in_betas_smpl'], g['in_valid_smpl'], num_iter=init_iter)
    out = {'w17': np.array(W17, np.float32), 'init_iter': np.int64(init_iter),
           'init_loss_2d': np.array([float(l['loss_2d']) for l in log], np.float32),
           'init_poses_T': opt.poses_T.detach().numpy().copy(), 'init_zmax_lin': opt.zmax_lin.detach().numpy().copy()}
    np.savez_compressed(os.path.join(HERE, 'init_w17.npz'), **out)
    print('wrote init_w17.npz', out['init_loss_2d'][[0, -1]], out['init_poses_T'][0, :, 0])

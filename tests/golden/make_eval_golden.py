"""Golden vectors of the reference's evaluation path: ``mhmocap/evaluate.py`` +
``mhmocap/eval_mupots.py:compute_mm_pck_results`` run UNMODIFIED on a synthetic sequence with the synthetic SMPL model.
Build-container only (needs /root/reference).  Output: tests/golden/eval_kat.npz.

Usage:  python tests/golden/make_eval_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)

from oracle import synth  # noqa: E402

MODEL_DIR = '/tmp/mh_golden_model'


def main():
    sys.argv = [sys.argv[0]]
    import mhmocap.smpl as rsmpl
    import mhmocap.evaluate as rev
    if not os.path.exists(os.path.join(MODEL_DIR, 'SMPL_NEUTRAL.pkl')):
        synth.write_model_dir(MODEL_DIR, seed=0)
    smpl = rsmpl.SMPL(MODEL_DIR,
                      J_reg_extra9_path=os.path.join(MODEL_DIR, 'J_regressor_extra.npy'),
                      J_reg_h36m17_path=os.path.join(MODEL_DIR, 'J_regressor_h36m.npy'),
                      J_reg_alphapose_path=os.path.join(MODEL_DIR, 'SMPL_AlphaPose_Regressor_RMSprop_6.npy'),
                      J_reg_mupots_path=os.path.join(MODEL_DIR, 'SMPL_MuPoTs_Regressor_v1.npy'))

    def SMPLPY(betas, poses):
        with torch.no_grad():
            return smpl(betas=torch.from_numpy(np.asarray(betas, np.float32)), poses=torch.from_numpy(np.asarray(poses, np.float32)))

    rng = np.random.default_rng(21)
    T, N, K = 6, 3, 4                                    # 4 annotated persons, 3 predictions: the assignment leaves one GT unmatched
    poses = (rng.normal(0, 0.25, (T, N, 72)) + rng.normal(0, 0.02, (T, 1, 72)).cumsum(0)).astype(np.float32)
    poses[..., 0] += np.pi                               # global orientation: Y down
    betas = np.repeat(rng.normal(0, 0.5, (1, N, 10)), T, axis=0).astype(np.float32)
    trans = np.stack([np.array([-1.0 + n, 0.1 * n, 3.5 + 0.8 * n]) + 0.02 * np.arange(T)[:, None] for n in range(N)], 1).astype(np.float32)[:, :, None, :]
    scale = (1.0 + 0.05 * rng.normal(0, 1, (1, N, 1, 1))).astype(np.float32)
    out_data = {'poses_T': trans, 'poses_smpl': poses, 'betas_smpl': betas, 'scale_factor': scale, 'valid_smpl': np.ones((T, N, 1), np.float32)}
    cam_K = np.array([[1100, 0, 640], [0, 1100, 360], [0, 0, 1]], np.float32)
    res = SMPLPY(betas.reshape(-1, 10), poses.reshape(-1, 72))
    jm = res['joints_mupots'].numpy().reshape(T, N, 17, 3)
    ja = res['joints_alphapose'].numpy().reshape(T, N, 17, 3)
    pred_abs = scale * jm + trans                        # (T, N, 17, 3)
    # ground truth: the predictions of a permutation of the persons + noise, one extra far-away person, some joints invisible
    perm = [2, 0, 1]
    gt = np.zeros((T, K, 17, 3), np.float32)
    for k, n in enumerate(perm):
        gt[:, k] = pred_abs[:, n] + rng.normal(0, 0.03, (T, 17, 3))
    gt[:, 3] = pred_abs[:, 0] + np.array([2.5, 0.0, 4.0], np.float32) + rng.normal(0, 0.03, (T, 17, 3))
    vis = (rng.random((T, K, 17, 1)) > 0.15).astype(np.float32)
    vis[2, 1, 14, 0] = 0                                 # an invisible root
    m17 = rev.compute_smpl_pred_error_3dproj(out_data, gt.copy(), vis.copy(), SMPLPY, cam_K)
    Kd = np.array([0.05, -0.01, 0.001, -0.002, 0.0005], np.float32)
    m17kd = rev.compute_smpl_pred_error_3dproj(out_data, gt.copy(), vis.copy(), SMPLPY, cam_K, Kd=Kd)
    # CMU-Panoptic layout (19 joints): random but consistent annotation
    gt19 = (pred_abs[:, [0, 1, 2, 0]].mean(2, keepdims=True) + rng.normal(0, 0.2, (T, K, 19, 3))).astype(np.float32)
    vis19 = (rng.random((T, K, 19, 1)) > 0.1).astype(np.float32)
    m19 = rev.compute_smpl_pred_error_3dproj(out_data, gt19.copy(), vis19.copy(), SMPLPY, cam_K)
    out = dict(poses=poses, betas=betas, trans=trans, scale=scale, cam_K=cam_K, Kd=Kd, joints_mupots=jm, joints_alphapose=ja,
               gt17=gt, vis17=vis, gt19=gt19, vis19=vis19)
    for tag, m in (('m17', m17), ('m17kd', m17kd), ('m19', m19)):
        for k, v in m.items():
            out[f'{tag}_{k}'] = np.asarray(v)
    # the scalar metrics of eval_mupots.compute_mm_pck_results (eval_mupots.py:24-40), computed with the reference's functions
    out['mm'] = np.array([1000 * rev.masked_average_error(m17['abs_dist'], m17['valid_joints']),
                          1000 * rev.masked_average_error(m17['rel_dist'], m17['valid_joints']),
                          1000 * rev.masked_average_error(m17['abs_root_pos_err'], m17['valid_root']),
                          100 * rev.masked_average_pck(m17['rel_dist'], m17['valid_joints'], 0.15),
                          100 * rev.masked_average_pck(m17['abs_root_pos_err'], m17['valid_root'], 0.25),
                          1000 * rev.masked_average_error(m17['abs_jitter'], m17['valid_joints'])], np.float64)
    np.savez_compressed(os.path.join(HERE, 'eval_kat.npz'), **out)
    print('wrote eval_kat.npz', {k: float(v) for k, v in zip(('mm_abs', 'mm_rel', 'mrpe', 'pck_rel', 'ap25', 'jitter'), out['mm'])})


if __name__ == '__main__':
    main()

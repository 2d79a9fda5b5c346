"""Generate the golden vectors under ``tests/golden/`` by running the UNMODIFIED reference.

Runs only in the build container (needs ``/root/reference``; nothing here is used
at test time on the GPU box).  The reference's ``mhmocap.optimizer`` is imported
as-is with two stand-ins ahead of it on ``sys.path``: ``oracle/pytorch3d_shim``
(the reference's PyTorch3D dependency is un-vendored and not installable -- the
rasteriser part of every vector is therefore "parity unpinned") and a synthetic
``SMPL_NEUTRAL.pkl`` (``oracle.synth``; the real model is licence-gated).

Outputs (all small, committed):
  kat_functions.npz      known-answer vectors of the reference's own functions
                         (rodrigues, SMPL forward, projection, calibration, erosion,
                         losses, one-euro, softplus, inverse projection)
  fit_c1.npz             config C1 (1 person x 4 frames, 96x64): inputs, init result,
                         teacher-forcing snapshots (params, state, losses, gradients)
                         at chosen cycles and the final variables of a 52-cycle fit
  fit_n2.npz             same for 2 persons x 4 frames (occlusion order exercised)

Usage:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'pytorch3d_shim'))
sys.path.insert(1, REF)

from oracle import synth  # noqa: E402

MODEL_DIR = '/tmp/mh_golden_model'
SNAP_CYCLES = (0, 1, 30, 31, 50, 51)


def import_reference():
    sys.argv = [sys.argv[0]]
    import mhmocap.optimizer as ropt
    import mhmocap.smpl as rsmpl
    import mhmocap.transforms as rtr
    import mhmocap.losses as rlo
    import mhmocap.morphology as rmo
    import mhmocap.one_euro_filter as roe
    return ropt, rsmpl, rtr, rlo, rmo, roe


def make_kats(ropt, rsmpl, rtr, rlo, rmo, roe):
    rng = np.random.default_rng(7)
    out = {}
    rv = np.concatenate([rng.normal(0, 1.0, (13, 3)), np.zeros((1, 3)), [[3, 0, 0]], [[0.1, -0.2, 0.3]]]).astype(np.float32)
    out['rodrigues_in'] = rv
    out['rodrigues_out'] = rsmpl.batch_rodrigues(torch.from_numpy(rv)).numpy()
    smpl = rsmpl.SMPL(MODEL_DIR,
                      J_reg_extra9_path=os.path.join(MODEL_DIR, 'J_regressor_extra.npy'),
                      J_reg_h36m17_path=os.path.join(MODEL_DIR, 'J_regressor_h36m.npy'),
                      J_reg_alphapose_path=os.path.join(MODEL_DIR, 'SMPL_AlphaPose_Regressor_RMSprop_6.npy'),
                      J_reg_mupots_path=os.path.join(MODEL_DIR, 'SMPL_MuPoTs_Regressor_v1.npy'))
    betas = rng.normal(0, 0.7, (5, 10)).astype(np.float32)
    poses = rng.normal(0, 0.4, (5, 72)).astype(np.float32)
    poses[0] = 0
    res = smpl(betas=torch.from_numpy(betas), poses=torch.from_numpy(poses))
    out['smpl_betas'] = betas; out['smpl_poses'] = poses
    out['smpl_verts'] = res['verts'].detach().numpy()
    out['smpl_joints24'] = res['joints_smpl24'].detach().numpy()
    out['smpl_joints_alphapose'] = res['joints_alphapose'].detach().numpy()
    out['smpl_joints_mupots'] = res['joints_mupots'].detach().numpy()
    # autograd gradients of a fixed linear functional of verts / joints
    gv = rng.normal(0, 1, out['smpl_verts'].shape).astype(np.float32)
    gj = rng.normal(0, 1, out['smpl_joints_alphapose'].shape).astype(np.float32)
    bt = torch.from_numpy(betas).requires_grad_(True); pt = torch.from_numpy(poses).requires_grad_(True)
    res = smpl(betas=bt, poses=pt)
    (torch.sum(res['verts'] * torch.from_numpy(gv)) + torch.sum(res['joints_alphapose'] * torch.from_numpy(gj))).backward()
    out['smpl_gverts'] = gv; out['smpl_gjoints'] = gj
    out['smpl_gbetas'] = bt.grad.numpy(); out['smpl_gposes'] = pt.grad.numpy()
    K = np.array([[1000, 0, 640], [0, 1000, 360], [0, 0, 1]], np.float32)
    p = np.array([[[0.5, -0.25, 4], [-1, 1, 2]]], np.float32)
    out['proj_K'] = K; out['proj_p'] = p
    out['proj_out'] = rtr.camera_projection_torch(torch.from_numpy(p), torch.from_numpy(K)[None]).numpy()
    Kd = np.array([0.1, 0.01, 0.001, 0.002, 0.0001], np.float32)
    out['proj_Kd'] = Kd
    out['proj_out_kd'] = rtr.camera_projection_torch(torch.from_numpy(p), torch.from_numpy(K)[None], Kd=Kd).numpy()
    out['calib_land'] = rtr.compute_calibration_matrix(1, 100, K, (1280, 720))
    K2 = np.array([[600, 0, 250], [0, 610, 260], [0, 0, 1]], np.float32)
    out['calib_K2'] = K2
    out['calib_square'] = rtr.compute_calibration_matrix(1, 100, K2, (512, 512))
    out['calib_port'] = rtr.compute_calibration_matrix(1, 100, K2, (480, 640))
    uvd = np.array([[[765, 297.5, 4], [140, 860, 2]]], np.float32)
    out['invproj_in'] = uvd
    out['invproj_out'] = rtr.camera_inverse_projection_torch(torch.from_numpy(uvd), torch.from_numpy(K)[None]).numpy()
    out['focal_720_60'] = np.float64(rtr.get_focal(720, 60))
    out['softplus_in'] = np.array([-3, 0, 1, 7.5], np.float32)
    out['softplus_out'] = rtr.softplus(torch.from_numpy(out['softplus_in'])).numpy()
    seg = (rng.random((2, 1, 24, 31)) > 0.25).astype(np.float32)
    seg[:, :, 6:20, 8:25] = 1
    er = torch.nn.Sequential(rmo.Erode2D(kernel_size=3), rmo.Erode2D(kernel_size=3))
    out['erode_in'] = seg
    out['erode_out'] = er(torch.from_numpy(seg)).numpy()
    yp = rng.random((2, 3, 8, 9)).astype(np.float32) + 0.01
    yt = rng.random((2, 1, 8, 9)).astype(np.float32) + 0.01
    mk = (rng.random((2, 3, 8, 9)) > 0.5).astype(np.float32)
    out['loss_yp'] = yp; out['loss_yt'] = yt; out['loss_mk'] = mk
    out['loss_avg_depth'] = rlo.build_avg_depth_loss_fn()(torch.from_numpy(yp), torch.from_numpy(yt), torch.from_numpy(mk)).numpy()
    out['loss_masked_mse'] = rlo.build_masked_mse_loss_fn()(torch.from_numpy(yp[0, 0]), torch.from_numpy(yt[0, 0]), torch.from_numpy(mk[0, 0])).numpy()
    # One-Euro with the optimiser's cumulative-time quirk (optimizer.py:664-675).  NB: on a CPU tensor the
    # reference filters IN PLACE (x.cpu().detach().numpy() aliases x, :665) -- pass copies here.
    class _O(object):
        device = 'cpu'
    y = np.array([0, 1, 0.5, 2, 1.5], np.float32)
    out['oef_in'] = y
    out['oef_out'] = ropt.SMPLDepthSequenceOptimizer.one_euro_filter(_O(), torch.from_numpy(y.copy()), min_cutoff=0.01, beta=0.02).numpy()
    y2 = rng.normal(0, 1, (9, 4, 3)).astype(np.float32).cumsum(0)
    out['oef2_in'] = y2
    out['oef2_out'] = ropt.SMPLDepthSequenceOptimizer.one_euro_filter(_O(), torch.from_numpy(y2.copy()), min_cutoff=0.001, beta=0.5).numpy()
    np.savez_compressed(os.path.join(HERE, 'kat_functions.npz'), **out)
    print('wrote kat_functions.npz', len(out), 'arrays')


class ListLoader(object):
    """Re-iterable yielding the reference's batch dicts (contiguous frames, shuffle=False)."""
    def __init__(self, inputs, batch):
        self.inputs, self.batch = inputs, batch
        self.T = len(inputs['idxs'])

    def __iter__(self):
        for s in range(0, self.T, self.batch):
            yield {k: torch.from_numpy(v[s:s + self.batch]) for k, v in self.inputs.items()}


def run_fit(ropt, name, N, T, W, H, batch, num_iter=52, init_iter=30, seed=1):
    inputs, cam_K, motion = synth.make_sequence(MODEL_DIR, N, T, W, H, seed)
    coefs = dict(proj2d_loss_coef=1.0, depth_loss_coef=0.05, silhouette_loss_coef=0.1, reg_velocity_coef=0.05,
                 reg_verts_filter_coef=0.002, reg_poses_coef=0.002, reg_scales_coef=1e-4,
                 reg_contact_coef=0.001, reg_foot_sliding_coef=0.01)     # configs/predict_mupots.yml:17-25
    opt = ropt.SMPLDepthSequenceOptimizer(image_size=(W, H), num_frames=T, cam_K=cam_K, device='cpu',
                                          smpl_model_parameters_path=MODEL_DIR, **coefs)
    torch.manual_seed(0)
    init_log = opt.init_optimized_variables(inputs['pose2d'], inputs['poses_smpl'], inputs['betas_smpl'],
                                            inputs['valid_smpl'], num_iter=init_iter)
    out = {'meta_NTWH_batch': np.array([N, T, W, H, batch, num_iter, init_iter]), 'cam_K': cam_K}
    for k, v in inputs.items():
        out['in_' + k] = v
    out['init_loss_2d'] = np.array([float(l['loss_2d']) for l in init_log], np.float32)
    out['init_poses_T'] = opt.poses_T.detach().numpy().copy()
    out['init_zmax_lin'] = opt.zmax_lin.detach().numpy().copy()
    out['init_betas'] = opt.betas_smpl.detach().numpy().copy()

    names = ['poses_T', 'poses_smpl', 'betas', 'zmin_lin', 'zmax_lin', 'xscale']
    state = {'cycle': 0}
    orig_step = torch.optim.RMSprop.step

    def step(self_, *a, **kw):
        c = state['cycle']
        if c in SNAP_CYCLES:
            ps = self_.param_groups[0]['params']
            for nm, p in zip(names, ps):
                out[f'c{c}_p_{nm}'] = p.detach().numpy().copy()
                out[f'c{c}_g_{nm}'] = p.grad.detach().numpy().copy()
            # state the losses of this cycle were computed with (scene update of cycle c happens
            # before step(), so snapshot what the batch loop saw via the pre-cycle hook below)
        state['cycle'] += 1
        return orig_step(self_, *a, **kw)

    # snapshot scene / filter state at the START of each cycle: zero_grad is the first call of a cycle
    orig_zero = torch.optim.RMSprop.zero_grad

    def zero_grad(self_, *a, **kw):
        c = state['cycle']
        if c in SNAP_CYCLES:
            out[f'c{c}_scene_pcd'] = (opt.scene_pcd[0, 0].numpy().copy() if opt.scene_pcd is not None
                                      else np.zeros((0, 3), np.float32))
        return orig_zero(self_, *a, **kw)

    torch.optim.RMSprop.step = step
    torch.optim.RMSprop.zero_grad = zero_grad
    try:
        log = opt.fit(ListLoader(inputs, batch), num_iter=num_iter, verbose=False)
    finally:
        torch.optim.RMSprop.step = orig_step
        torch.optim.RMSprop.zero_grad = orig_zero
    for c in SNAP_CYCLES:
        if c < num_iter:
            for k, v in log[c].items():
                out[f'c{c}_log_{k}'] = np.float32(v)
    # filtered vertices exist from cycle 50 on and are constant until cycle 75
    if opt.verts_filtered is not None:
        out['verts_filtered'] = opt.verts_filtered.numpy().copy()
    for k in log[0].keys():
        out['log_' + k] = np.array([float(l[k]) for l in log], np.float32)
    fv = opt.get_optimized_variables()
    for k, v in fv.items():
        if v is not None:
            out['final_' + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print('wrote', name, {k: float(log[-1][k]) for k in log[-1]})


if __name__ == '__main__':
    assert os.path.isdir(REF), 'needs the reference mounted at /root/reference'
    synth.write_model_dir(MODEL_DIR, seed=0)
    mods = import_reference()
    make_kats(*mods)
    run_fit(mods[0], 'fit_c1.npz', N=1, T=4, W=96, H=64, batch=2)
    run_fit(mods[0], 'fit_n2.npz', N=2, T=4, W=96, H=64, batch=2, seed=3)

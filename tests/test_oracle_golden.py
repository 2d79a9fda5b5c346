"""The CPU oracle against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, COEFS
from oracle import refmath as rm, fit_ref, synth


@pytest.fixture(scope='module')
def kat():
    return np.load(os.path.join(GOLDEN, 'kat_functions.npz'))


def _mt(model):
    mt = {a: (torch.from_numpy(v) if v.dtype == np.float32 else v) for a, v in model.items()}
    mt['parents'] = [int(p) for p in model['parents']]
    return mt


def test_model_is_the_golden_model(model, kat):
    # the golden vectors are only meaningful if the seeded model regenerates identically
    o = rm.smpl_forward(_mt(model), torch.from_numpy(kat['smpl_betas'][:1]), torch.from_numpy(kat['smpl_poses'][:1]))
    assert np.abs(o['verts'].numpy() - kat['smpl_verts'][:1]).max() < 1e-6


def test_rodrigues(kat):
    out = rm.rodrigues(torch.from_numpy(kat['rodrigues_in'])).numpy()
    assert np.abs(out - kat['rodrigues_out']).max() < 1e-6
    # SURVEY.md section 4 known answers
    r = rm.rodrigues(torch.tensor([[0.1, -0.2, 0.3], [3.0, 0, 0], [0, 0, 0]])).numpy()
    assert np.allclose(r[0], [[0.93575484, -0.30293274, -0.18054008], [0.28316498, 0.95058066, -0.12733456],
                              [0.21019170, 0.06803133, 0.97529030]], atol=1e-6)
    assert np.allclose(r[1], [[1, 0, 0], [0, -0.98999250, -0.14112000], [0, 0.14112000, -0.98999250]], atol=1e-6)
    assert np.array_equal(r[2], np.eye(3, dtype=np.float32))


def test_smpl_forward_and_grad(model, kat):
    mt = _mt(model)
    b = torch.from_numpy(kat['smpl_betas']).requires_grad_(True)
    p = torch.from_numpy(kat['smpl_poses']).requires_grad_(True)
    o = rm.smpl_forward(mt, b, p)
    j17 = rm.regress_joints(mt['J_regressor_alphapose'], o['verts'])
    assert np.abs(o['verts'].detach().numpy() - kat['smpl_verts']).max() < 1e-5
    assert np.abs(o['joints24'].detach().numpy() - kat['smpl_joints24']).max() < 1e-5
    assert np.abs(j17.detach().numpy() - kat['smpl_joints_alphapose']).max() < 1e-5
    jm = rm.regress_joints(mt['J_regressor_mupots'], o['verts'])
    assert np.abs(jm.detach().numpy() - kat['smpl_joints_mupots']).max() < 1e-5
    (torch.sum(o['verts'] * torch.from_numpy(kat['smpl_gverts'])) + torch.sum(j17 * torch.from_numpy(kat['smpl_gjoints']))).backward()
    for got, ref in ((b.grad.numpy(), kat['smpl_gbetas']), (p.grad.numpy(), kat['smpl_gposes'])):
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
    # verts are exactly invariant to the two hand joints (smpl.py:544-546)
    assert np.abs(p.grad.numpy()[:, 66:]).max() == 0


def test_camera_functions(kat):
    K = torch.from_numpy(kat['proj_K'])[None]
    p = torch.from_numpy(kat['proj_p'])
    assert np.allclose(rm.camera_projection(p, K).numpy(), kat['proj_out'], atol=1e-4)
    assert np.allclose(rm.camera_projection(p, K).numpy(), [[[765, 297.5], [140, 860]]], atol=1e-4)
    assert np.allclose(rm.camera_projection(p, K, kat['proj_Kd']).numpy(), kat['proj_out_kd'], atol=1e-4)
    assert np.allclose(rm.camera_inverse_projection(torch.from_numpy(kat['invproj_in']), K).numpy(), kat['invproj_out'], atol=1e-5)
    assert np.array_equal(rm.compute_calibration_matrix(1, 100, kat['proj_K'], (1280, 720)), kat['calib_land'])
    assert np.array_equal(rm.compute_calibration_matrix(1, 100, kat['calib_K2'], (512, 512)), kat['calib_square'])
    assert np.array_equal(rm.compute_calibration_matrix(1, 100, kat['calib_K2'], (480, 640)), kat['calib_port'])
    assert abs(rm.get_focal(720, 60) - float(kat['focal_720_60'])) < 1e-9
    assert np.allclose(rm.softplus(torch.from_numpy(kat['softplus_in'])).numpy(), kat['softplus_out'], atol=1e-6)


def test_losses_and_erosion(kat):
    assert np.array_equal(rm.erode5_twice3(torch.from_numpy(kat['erode_in'])).numpy(), kat['erode_out'])
    yp, yt, mk = (torch.from_numpy(kat[k]) for k in ('loss_yp', 'loss_yt', 'loss_mk'))
    assert np.allclose(rm.avg_depth_loss(yp, yt, mk).numpy(), kat['loss_avg_depth'], rtol=1e-6)
    assert np.allclose(rm.masked_mse_loss(yp[0, 0], yt[0, 0], mk[0, 0]).numpy(), kat['loss_masked_mse'], rtol=1e-6)


def test_one_euro(kat):
    assert np.allclose(rm.one_euro_filter_sequence(kat['oef_in'], 0.01, 0.02), kat['oef_out'], atol=1e-7)
    assert np.allclose(kat['oef_out'], [0, 0.02700325, 0.05314534, 0.31941577, 0.50352543], atol=1e-7)
    assert np.allclose(rm.one_euro_filter_sequence(kat['oef2_in'], 0.001, 0.5), kat['oef2_out'], atol=1e-6)


def _load_fit(name):
    g = np.load(os.path.join(GOLDEN, name))
    N, T, W, H, batch, num_iter, init_iter = [int(x) for x in g['meta_NTWH_batch']]
    data = {k[3:]: g[k] for k in g.files if k.startswith('in_')}
    return g, data, (N, T, W, H, batch, num_iter, init_iter)


@pytest.mark.parametrize('name', ['fit_c1.npz', 'fit_n2.npz'])
def test_synthetic_sequence_regenerates(model_dir, name):
    # bench/test inputs come from oracle.synth: it must reproduce what the golden run consumed
    g, data, (N, T, W, H, *_r) = _load_fit(name)
    seed = 1 if name == 'fit_c1.npz' else 3
    inputs, cam_K, _ = synth.make_sequence(model_dir, N, T, W, H, seed)
    assert np.array_equal(cam_K, g['cam_K'])
    for k in ('depths', 'seg_mask', 'pose2d', 'poses_smpl', 'betas_smpl'):
        assert np.allclose(inputs[k], data[k], atol=1e-6), k


def test_init_stage_with_joint_weights(model):
    """Non-uniform ``pose17j_weights`` weigh the 2-D residual of hot loop A too (``optimizer.py:754-756``): golden vector of the
    unmodified reference (``tests/golden/make_init_w17_golden.py``)."""
    g, data, (N, T, W, H, batch, num_iter, init_iter) = _load_fit('fit_n2.npz')
    k = np.load(os.path.join(GOLDEN, 'init_w17.npz'))
    fr = fit_ref.FitRef(model, (W, H), T, g['cam_K'], COEFS, pose17j_weights=k['w17'])
    log = fr.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=int(k['init_iter']))
    assert np.abs(fr.poses_T.detach().numpy() - k['init_poses_T']).max() < 1e-4          # metres
    l2d = np.array([float(l['loss_2d']) for l in log])
    assert np.abs(l2d - k['init_loss_2d']).max() <= 1e-5 * k['init_loss_2d'].max()
    assert np.abs(k['init_poses_T'] - g['init_poses_T']).max() > 1e-2                    # the weights do change the answer


@pytest.mark.parametrize('name', ['fit_c1.npz', 'fit_n2.npz'])
def test_init_stage(model, name):
    g, data, (N, T, W, H, batch, num_iter, init_iter) = _load_fit(name)
    fr = fit_ref.FitRef(model, (W, H), T, g['cam_K'], COEFS)
    log = fr.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'],
                                      num_iter=init_iter)
    assert np.abs(fr.poses_T.detach().numpy() - g['init_poses_T']).max() < 1e-4          # metres
    l2d = np.array([float(l['loss_2d']) for l in log])
    assert np.abs(l2d - g['init_loss_2d']).max() <= 1e-5 * g['init_loss_2d'].max()
    assert np.allclose(fr.zmax_lin.detach().numpy(), g['init_zmax_lin'], atol=1e-4)


@pytest.mark.parametrize('name', ['fit_c1.npz', 'fit_n2.npz'])
@pytest.mark.parametrize('cycle', [0, 1, 30, 31, 50, 51])
def test_teacher_forced_cycle(model, name, cycle):
    """Same parameters/state in -> same 9 losses and 6 gradient tensors out as the reference."""
    g, data, (N, T, W, H, batch, num_iter, init_iter) = _load_fit(name)
    fr = fit_ref.FitRef(model, (W, H), T, g['cam_K'], COEFS)
    c = cycle
    fr.set_variables(g[f'c{c}_p_poses_T'], g[f'c{c}_p_poses_smpl'], g[f'c{c}_p_betas'], data['valid_smpl'],
                     g[f'c{c}_p_zmin_lin'], g[f'c{c}_p_zmax_lin'], g[f'c{c}_p_xscale'])
    fr.betas_ref = torch.from_numpy(g['init_betas'])
    if len(g[f'c{c}_scene_pcd']):
        fr.set_scene_pcd(g[f'c{c}_scene_pcd'])
    if c >= 50:
        fr.verts_filtered = torch.from_numpy(g['verts_filtered'])
        fr.poses_T_filtered = True
    batches = [np.arange(s, min(s + batch, T)) for s in range(0, T, batch)]
    log, _ = fr.cycle_grads(data, batches)
    for k, v in log.items():
        ref = float(g[f'c{c}_log_{k}'])
        assert abs(v - ref) <= 1e-4 * abs(ref) + 1e-9, (k, v, ref)
    names = ['poses_T', 'poses_smpl', 'betas', 'zmin_lin', 'zmax_lin', 'xscale']
    for nm, p in zip(names, fr.leaves()):
        ref = g[f'c{c}_g_{nm}']
        assert np.abs(p.grad.numpy() - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, nm

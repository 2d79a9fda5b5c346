"""Evaluation path against golden vectors produced by the UNMODIFIED reference
(`tests/golden/make_eval_golden.py`: mhmocap/evaluate.py + eval_mupots.py on a synthetic sequence).

CPU part: the host logic (projection, Hungarian matching, per-pair distances, masked averages) with the reference's own SMPL
joints fed in -> bit-level agreement.  GPU part: the joints come from the device (`mh_smpl_regress`)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden', 'eval_kat.npz')
KEYS = ('abs_dist', 'rel_dist', 'valid_joints', 'abs_root_pos_err', 'valid_root', 'abs_jitter')


def _load_eval_module():
    """evaluation.py has no dependency on libmhopt.so: load it on its own (the package import would need the library)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('mh_evaluation', os.path.join(ROOT, 'scene-aware-3d-multi-human_b200', 'evaluation.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _optvar(g):
    return {'poses_T': g['trans'], 'poses_smpl': g['poses'], 'betas_smpl': g['betas'], 'scale_factor': g['scale']}


def _golden_joints(g):
    def fn(betas, poses, which):
        assert np.array_equal(np.asarray(poses).reshape(g['poses'].shape), g['poses'])
        return g['joints_' + which].reshape(-1, 17, 3)
    return fn


@pytest.mark.parametrize('case', ['m17', 'm17kd', 'm19'])
def test_host_metrics_match_the_reference(case):
    ev = _load_eval_module()
    g = np.load(GOLD)
    gt, vis = (g['gt19'], g['vis19']) if case == 'm19' else (g['gt17'], g['vis17'])
    m = ev.compute_smpl_pred_error_3dproj(_optvar(g), gt, vis, _golden_joints(g), g['cam_K'], Kd=g['Kd'] if case == 'm17kd' else None)
    for k in KEYS:
        ref = g[f'{case}_{k}']
        assert m[k].shape == ref.shape and m[k].dtype == ref.dtype, k
        assert np.array_equal(m[k], ref), (case, k, np.abs(m[k] - ref).max())
    assert m['valid_joints'].sum() > 100 and m['valid_root'].sum() > 10          # the case is not degenerate
    assert not m['valid_joints'][:, 3].any()                                     # 4 annotated persons, 3 predictions: the last row stays empty


def test_sequence_metrics_match_the_reference():
    ev = _load_eval_module()
    g = np.load(GOLD)
    r = ev.compute_mm_pck_results(_optvar(g), g['gt17'], g['vis17'], _golden_joints(g), g['cam_K'])
    got = np.array([r[k] for k in ('mm_abs_error', 'mm_rel_error', 'mm_mrpe', 'pck_rel', 'ap25_root', 'abs_jitter')], np.float64)
    assert np.array_equal(got, g['mm']), (got, g['mm'])


def test_masked_averages_edge_cases():
    ev = _load_eval_module()
    d = np.array([[0.1, 0.2], [0.3, 0.4]], np.float32)
    assert ev.masked_average_error(d, np.zeros_like(d)) == 0.0                    # nothing visible: 0 / max(0, 1)
    assert ev.masked_average_pck(d, np.ones_like(d), 0.2) == 0.5                  # <= threshold
    with pytest.raises(ValueError):
        ev.masked_average_error(d, np.ones(3, np.float32))
    with pytest.raises(ValueError):
        ev.compute_smpl_pred_error_3dproj({'poses_T': np.zeros((1, 1, 1, 3)), 'scale_factor': np.ones((1, 1, 1, 1)), 'poses_smpl': np.zeros((1, 1, 72)),
                                           'betas_smpl': np.zeros((1, 1, 10))}, np.zeros((1, 1, 16, 3)), np.zeros((1, 1, 16, 1)), None, np.eye(3))


@pytest.mark.gpu
def test_device_joints_and_metrics():
    """SMPL + sparse joint regression on the GPU vs the reference's joints (1e-5 m), then the whole evaluation through the device."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import __graft_entry__ as ge
    import gpu_harness as gh
    pkg = ge.load_package()
    ev = pkg.evaluation
    g = np.load(GOLD)
    fit, data, meta = gh.load_fit('fit_c1.npz')
    opt = gh.make_optimizer(pkg, fit, data, meta)
    gh.prepare(opt, fit, data, meta, ingest=False)
    md = gh.model_dir()
    joints = ev.SMPLJoints(opt, {'mupots': np.load(os.path.join(md, 'SMPL_MuPoTs_Regressor_v1.npy')),
                                 'alphapose': np.load(os.path.join(md, 'SMPL_AlphaPose_Regressor_RMSprop_6.npy'))})
    for which in ('mupots', 'alphapose'):
        j = joints(g['betas'].reshape(-1, 10), g['poses'].reshape(-1, 72), which)
        assert np.abs(j - g['joints_' + which].reshape(-1, 17, 3)).max() < 1e-5
    kat = np.load(os.path.join(gh.GOLDEN, 'kat_functions.npz'))
    assert np.abs(joints(kat['smpl_betas'], kat['smpl_poses'], 'mupots') - kat['smpl_joints_mupots']).max() < 1e-5
    r = ev.compute_mm_pck_results(_optvar(g), g['gt17'], g['vis17'], joints, g['cam_K'])
    got = np.array([r[k] for k in ('mm_abs_error', 'mm_rel_error', 'mm_mrpe', 'pck_rel', 'ap25_root', 'abs_jitter')], np.float64)
    assert np.abs(got - g['mm']).max() <= 0.02, (got, g['mm'])                   # mm / percent: 1e-5 m of joint noise


@pytest.mark.skipif(not os.path.isdir('/root/reference/mhmocap'), reason='differential check against the reference needs /root/reference (build container only)')
@pytest.mark.parametrize('seed,T,N,K,J', [(0, 3, 2, 2, 17), (1, 4, 3, 2, 17), (2, 2, 1, 3, 19), (3, 5, 4, 4, 17), (4, 3, 2, 5, 19)])
def test_differential_against_the_reference(seed, T, N, K, J):
    """Random sequences (more / fewer predictions than annotations, invisible joints and roots, both joint layouts): the mirror and the
    UNMODIFIED reference, fed the same joints, agree bit for bit."""
    sys.path.insert(0, '/root/reference')
    argv, sys.argv = sys.argv, sys.argv[:1]
    try:
        import mhmocap.evaluate as rev
    finally:
        sys.argv = argv
        sys.path.remove('/root/reference')
    import torch
    ev = _load_eval_module()
    rng = np.random.default_rng(100 + seed)
    jm = rng.normal(0, 0.3, (T * N, 17, 3)).astype(np.float32)
    ja = rng.normal(0, 0.3, (T * N, 17, 3)).astype(np.float32)
    optvar = {'poses_T': (rng.normal(0, 0.5, (T, N, 1, 3)) + [0, 0, 4]).astype(np.float32), 'poses_smpl': rng.normal(0, 0.2, (T, N, 72)).astype(np.float32),
              'betas_smpl': rng.normal(0, 0.5, (T, N, 10)).astype(np.float32), 'scale_factor': (1 + 0.1 * rng.normal(0, 1, (1, N, 1, 1))).astype(np.float32),
              'valid_smpl': np.ones((T, N, 1), np.float32)}
    gt = (rng.normal(0, 0.5, (T, K, J, 3)) + [0, 0, 4]).astype(np.float32)
    vis = (rng.random((T, K, J, 1)) > 0.3).astype(np.float32)
    cam_K = np.array([[900, 0, 500], [0, 950, 300], [0, 0, 1]], np.float32)

    def SMPLPY(betas, poses):
        return {'joints_mupots': torch.from_numpy(jm), 'joints_alphapose': torch.from_numpy(ja)}

    ref = rev.compute_smpl_pred_error_3dproj(optvar, gt.copy(), vis.copy(), SMPLPY, cam_K)
    got = ev.compute_smpl_pred_error_3dproj(optvar, gt, vis, lambda b, p, which: jm if which == 'mupots' else ja, cam_K)
    for k in KEYS:
        assert np.array_equal(np.asarray(ref[k]), got[k]), (k, seed)
    assert ev.masked_average_error(got['abs_dist'], got['valid_joints']) == rev.masked_average_error(ref['abs_dist'], ref['valid_joints'])
    assert ev.masked_average_pck(got['rel_dist'], got['valid_joints'], 0.15) == rev.masked_average_pck(ref['rel_dist'], ref['valid_joints'], 0.15)

"""Parity of the CUDA path (through the package / C ABI of libmhopt.so) with the golden vectors of the UNMODIFIED
reference (tests/golden, produced by tests/golden/make_golden.py) and with the CPU oracle, plus size-independent
properties at larger sizes.  Needs a B200: run with `-m gpu`.

Tolerances (float32, stated per test): SMPL vertices / joints 1e-5 m absolute; every loss 1e-4 relative; every gradient
1e-3 of its largest reference entry; optimiser updates 1e-6 relative."""
import os
import sys

import numpy as np
import pytest

import gpu_harness as gh

pytestmark = pytest.mark.gpu

CYCLES = [0, 1, 30, 31, 50, 51]


@pytest.fixture(scope='module')
def pkg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope='module')
def L(pkg):
    return sys.modules[pkg.__name__ + '._lib']


@pytest.fixture(scope='module')
def kat():
    return np.load(os.path.join(gh.GOLDEN, 'kat_functions.npz'))


@pytest.fixture(scope='module')
def c1(pkg):
    g, data, meta = gh.load_fit('fit_c1.npz')
    opt = gh.make_optimizer(pkg, g, data, meta)
    gh.prepare(opt, g, data, meta)
    return opt, g, data, meta


@pytest.fixture(scope='module')
def n2(pkg):
    g, data, meta = gh.load_fit('fit_n2.npz')
    opt = gh.make_optimizer(pkg, g, data, meta)
    gh.prepare(opt, g, data, meta)
    return opt, g, data, meta


def test_smpl_forward_kat(c1, kat):
    opt = c1[0]
    v, j = opt.smpl_forward(kat['smpl_betas'], kat['smpl_poses'])
    assert np.abs(v - kat['smpl_verts']).max() < 1e-5                        # metres
    assert np.abs(j - kat['smpl_joints_alphapose']).max() < 1e-5
    # ragged batch sizes go through the same chunked path
    v1, j1 = opt.smpl_forward(kat['smpl_betas'][:1], kat['smpl_poses'][:1])
    assert np.array_equal(v1, v[:1]) and np.array_equal(j1, j[:1])


def test_pose_corrective_contraction_on_tensor_cores(c1, L):
    """`v_posed - v_shaped = pose_feature . posedirs` (smpl.py:549-553): the tcgen05 / TMEM kernel (3 x TF32 split) against float64
    and against the FP32 SIMT kernel, on a row count that is not a multiple of the 128-row tile (the last column tile is 192 wide)."""
    ctx = c1[0].ctx
    eye = np.eye(192, dtype=np.float32)
    basis = np.zeros((192, L.LD3V), np.float32)
    ctx.call('mh_debug_gemm_fwd', L.ptr(eye), L.ptr(basis), 192, 0)             # exact: one product per output
    assert np.abs(basis).max() > 0
    via_tc = np.zeros_like(basis)
    ctx.call('mh_debug_gemm_fwd', L.ptr(eye), L.ptr(via_tc), 192, 1)
    assert np.abs(via_tc - basis).max() <= 1e-6 * np.abs(basis).max()           # hi + lo reproduces every entry (lo . lo dropped)
    rng = np.random.default_rng(11)
    M = 300
    A = rng.normal(0, 0.2, (M, 192)).astype(np.float32)
    ref = A.astype(np.float64) @ basis.astype(np.float64)
    out = {}
    for use_tc in (0, 1):
        C = np.full((M, L.LD3V), np.nan, np.float32)
        ctx.call('mh_debug_gemm_fwd', L.ptr(A), L.ptr(C), M, use_tc)
        out[use_tc] = C
    scale = np.abs(ref).max()
    assert np.abs(out[0] - ref).max() <= 1e-6 * scale                           # FP32 FMA chain
    assert np.abs(out[1] - ref).max() <= 3e-6 * scale                           # tensor cores, fp32 accumulation in TMEM
    assert np.array_equal(out[1][:, L.LD3V - 2:], out[0][:, L.LD3V - 2:])       # the padding columns stay exactly zero


def test_backward_contraction_on_tensor_cores(c1, L):
    """`dL/dpose_feature = dL/dv_posed . posedirs^T` plus the shape-blend rows (K = 20672 in 19 split-K ranges): tcgen05 / TMEM kernel
    against float64 and against the FP32 SIMT kernel."""
    ctx = c1[0].ctx
    eye = np.eye(192, dtype=np.float32)
    basis = np.zeros((192, L.LD3V), np.float32)
    ctx.call('mh_debug_gemm_fwd', L.ptr(eye), L.ptr(basis), 192, 0)
    rng = np.random.default_rng(12)
    M = 200                                                                     # not a multiple of the 128-row tile
    E = rng.normal(0, 1e-3, (M, L.LD3V)).astype(np.float32)
    E[:, 3 * L.V:] = 0
    ref = E.astype(np.float64) @ basis.astype(np.float64).T                     # (M, 192)
    out = {}
    for use_tc in (0, 1):
        D = np.full((M, 208), np.nan, np.float32)
        ctx.call('mh_debug_gemm_bwd', L.ptr(E), L.ptr(D), M, use_tc)
        out[use_tc] = D
    scale = np.abs(ref).max()
    e0, e1 = np.abs(out[0][:, :192] - ref).max() / scale, np.abs(out[1][:, :192] - ref).max() / scale
    assert e0 <= 2e-5, ('simt', e0)                                             # 1088 sequential FP32 additions per split
    assert e1 <= 2e-5, ('tensor cores', e1)
    s2 = np.abs(out[0][:, 192:]).max()
    e2 = np.abs(out[1][:, 192:] - out[0][:, 192:]).max() / s2
    assert s2 > 0 and e2 <= 4e-5, ('shape-blend rows', e2)                      # shape-blend rows (+ zero padding)


def test_one_euro_kat(c1, kat):
    opt = c1[0]
    assert np.array_equal(opt.one_euro_filter(kat['oef_in'], 0.01, 0.02).cpu().numpy(), kat['oef_out'])
    assert np.abs(opt.one_euro_filter(kat['oef2_in'], 0.001, 0.5).cpu().numpy() - kat['oef2_out']).max() <= 1e-6
    one = opt.one_euro_filter(kat['oef2_in'][:1], 0.001, 0.5).cpu().numpy()     # a single frame passes through
    assert np.array_equal(one, kat['oef2_in'][:1])


@pytest.mark.parametrize('which', ['c1', 'n2'])
def test_render_planes_vs_oracle(which, c1, n2, L):
    """zbuf[...,0] of the depth raster and the soft-silhouette alpha of every person-frame vs oracle.raster.

    The rasterisers are compared on IDENTICAL vertices (the device's own SMPL output, read back): coverage, the blur-radius
    tests and the 4 nearest fragments are discrete in the vertex positions, so a 1e-8 m difference between two correct SMPL
    evaluations (FP32 FMA chain vs tensor-core accumulation vs torch's CPU sgemm) can flip one fragment and move alpha by
    1e-3.  The SMPL forward itself is pinned by `test_smpl_forward_kat` (1e-5 m) and checked here to 1e-6 m."""
    import torch
    from oracle import raster, refmath as rm, synth
    opt, g, data, meta = c1 if which == 'c1' else n2
    N, T, W, H = meta[:4]
    c = 31
    ctx, st = opt.ctx, opt._stream()
    ctx.set_param(L.P_POSES_T, g[f'c{c}_p_poses_T'], st); ctx.set_param(L.P_POSES_SMPL, g[f'c{c}_p_poses_smpl'], st)
    ctx.set_param(L.P_BETAS, g[f'c{c}_p_betas'], st); ctx.set_param(L.P_XSCALE, g[f'c{c}_p_xscale'], st)
    model = synth.load_model_tensors(gh.model_dir())
    mt = {a: (torch.from_numpy(v) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v) for a, v in model.items()}
    mt['parents'] = [int(p) for p in model['parents']]
    faces = torch.from_numpy(model['faces'].astype(np.int64))
    Kndc = torch.from_numpy(rm.compute_calibration_matrix(1.0, 100.0, g['cam_K'], (W, H)))
    for t in range(T):
        for n in range(N):
            zb = np.zeros((H, W), np.float32); al = np.zeros((H, W), np.float32)
            ctx.call('mh_debug_render', t, n, L.ptr(zb), L.ptr(al))
            dev_verts = opt._view(L.BUF_VERTS).view(T + 2, N, L.LD3V)[1 + t, n, :3 * L.V].cpu().view(L.V, 3)
            with torch.no_grad():
                out = rm.smpl_forward(mt, torch.from_numpy(g[f'c{c}_p_betas'][0, n:n + 1]), torch.from_numpy(g[f'c{c}_p_poses_smpl'][t, n:n + 1]))
                s = np.float32(1.1) ** g[f'c{c}_p_xscale'].reshape(-1)[n]
                va = float(s) * out['verts'][0] + torch.from_numpy(g[f'c{c}_p_poses_T'][t, n])
                assert float((dev_verts - va).abs().max()) < 1e-6             # metres: the two SMPL evaluations agree
                z0, a0 = raster.render_person(dev_verts.clone(), faces, Kndc, H, W)
            z0, a0 = z0.numpy(), a0.numpy()
            assert (z0 > 0).sum() > 100
            assert np.array_equal(z0 > 0, zb > 0), (t, n)                     # identical coverage
            assert np.abs(zb - z0).max() < 1e-5                               # metres
            assert np.abs(al - a0).max() < 1e-4


@pytest.mark.parametrize('cycle', CYCLES)
@pytest.mark.parametrize('which', ['c1', 'n2'])
def test_teacher_forced_cycle_vs_reference(which, cycle, c1, n2):
    """Same parameters / scene cloud / filtered vertices in -> the reference's 9 logged losses and 6 gradient tensors out."""
    opt, g, data, meta = c1 if which == 'c1' else n2
    log, grads = gh.teacher_forced_cycle(opt, g, data, meta, cycle)
    for k, v in log.items():
        ref = float(g[f'c{cycle}_log_{k}'])
        assert abs(v - ref) <= 1e-4 * abs(ref) + 1e-9, (k, v, ref)
    for nm, gr in grads.items():
        ref = g[f'c{cycle}_g_{nm}'].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, nm
    if cycle >= 1:
        assert np.all(grads['poses_smpl'][..., 66:] == -0.002 * np.sign(g[f'c{cycle}_p_poses_smpl'] - data['poses_smpl'])[..., 66:])  # hands: prior only


def test_teacher_forced_cycle_vs_oracle(n2):
    opt, g, data, meta = n2
    log, grads = gh.teacher_forced_cycle(opt, g, data, meta, 51)
    olog, ograds = gh.oracle_cycle(g, data, meta, 51)
    for k, v in log.items():
        assert abs(v - olog[k]) <= 1e-4 * abs(olog[k]) + 1e-9, (k, v, olog[k])
    for nm, gr in grads.items():
        ref = ograds[nm].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, nm


def test_cycle_is_repeatable(n2):
    """Run-to-run BITWISE determinism of a cycle: the raster gradients are summed in 64-bit fixed point, the
    loss partials and the shared-leaf gradients in a fixed order, the foot-sliding scatter is a gather -- no float atomics remain."""
    opt, g, data, meta = n2
    runs = [gh.teacher_forced_cycle(opt, g, data, meta, 51) for _ in range(3)]
    for r in runs[1:]:
        for k in runs[0][0]:
            assert r[0][k] == runs[0][0][k], (k, r[0][k], runs[0][0][k])
        for nm in runs[0][1]:
            assert np.array_equal(r[1][nm], runs[0][1][nm]), nm


@pytest.mark.parametrize('name', ['fit_c1.npz', 'fit_n2.npz'])
def test_init_stage(pkg, L, name):
    """Hot loop A (Adam on the translations) vs the reference's result and loss curve."""
    g, data, meta = gh.load_fit(name)
    N, T, W, H, batch, num_iter, init_iter = meta
    opt = gh.make_optimizer(pkg, g, data, meta)
    log = opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=init_iter,
                                       batch_size=batch)
    pT = opt.ctx.get_param(L.P_POSES_T, (T, N, 1, 3))
    assert np.abs(pT - g['init_poses_T']).max() < 1e-4                        # metres
    l2d = np.array([float(l['loss_2d']) for l in log])
    assert np.abs(l2d - g['init_loss_2d']).max() <= 1e-4 * g['init_loss_2d'].max()
    assert np.allclose(opt.ctx.get_param(L.P_ZMAX_LIN, (T, 1, 1)), g['init_zmax_lin'], atol=1e-4)
    assert np.allclose(opt.ctx.get_param(L.P_BETAS, (1, N, 10)), g['init_betas'], atol=1e-7)


def test_init_stage_with_joint_weights(pkg, L):
    """``pose17j_weights`` in hot loop A (``optimizer.py:754-756``) vs the unmodified reference (``tests/golden/init_w17.npz``)."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch, num_iter, init_iter = meta
    k = np.load(os.path.join(gh.GOLDEN, 'init_w17.npz'))
    opt = gh.make_optimizer(pkg, g, data, meta, pose17j_weights=[float(v) for v in k['w17']])
    log = opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=int(k['init_iter']),
                                       batch_size=batch)
    pT = opt.ctx.get_param(L.P_POSES_T, (T, N, 1, 3))
    assert np.abs(pT - k['init_poses_T']).max() < 1e-4                        # metres
    l2d = np.array([float(l['loss_2d']) for l in log])
    assert np.abs(l2d - k['init_loss_2d']).max() <= 1e-4 * k['init_loss_2d'].max()
    assert np.allclose(opt.ctx.get_param(L.P_ZMAX_LIN, (T, 1, 1)), k['init_zmax_lin'], atol=1e-4)
    opt.ctx.close()


def test_regressor_rows_that_do_not_sum_to_one(pkg, L, tmp_path):
    """``J17 = R17 . V + T (1 - rowsum)``: with a 17-joint regressor whose rows do NOT sum to 1 (the shipped ones do) part of dL/dT
    bypasses the vertices.  Teacher-forced cycle vs the CPU oracle (autograd) with the same scaled regressor."""
    import torch
    from oracle import fit_ref, synth
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch, num_iter, init_iter = meta
    reg = np.load(os.path.join(gh.model_dir(), 'SMPL_AlphaPose_Regressor_RMSprop_6.npy'))            # (V, 17) on disk
    scale = np.linspace(0.8, 1.2, 17).astype(np.float32)
    path = str(tmp_path / 'scaled_regressor.npy')
    np.save(path, (reg * scale[None, :]).astype(np.float32))
    opt = gh.make_optimizer(pkg, g, data, meta, smpl_J_reg_alphapose_path=path)
    log, grads = gh.teacher_forced_cycle(opt, g, data, meta, 31)
    model = dict(synth.load_model_tensors(gh.model_dir()))
    model['J_regressor_alphapose'] = np.ascontiguousarray((reg * scale[None, :]).T.astype(np.float32))
    fr = fit_ref.FitRef(model, (W, H), T, g['cam_K'], gh.COEFS)
    c = 31
    fr.set_variables(g[f'c{c}_p_poses_T'], g[f'c{c}_p_poses_smpl'], g[f'c{c}_p_betas'], data['valid_smpl'],
                     g[f'c{c}_p_zmin_lin'], g[f'c{c}_p_zmax_lin'], g[f'c{c}_p_xscale'])
    fr.betas_ref = torch.from_numpy(g['init_betas'])
    fr.set_scene_pcd(g[f'c{c}_scene_pcd'])
    batches = [np.arange(s, min(s + batch, T)) for s in range(0, T, batch)]
    olog, _ = fr.cycle_grads(data, batches)
    ograds = {nm: p.grad.numpy().copy() for nm, p in zip(gh.NAMES, fr.leaves())}
    assert abs(log['loss_pose24j'] - olog['loss_pose24j']) <= 1e-4 * abs(olog['loss_pose24j'])
    assert olog['loss_pose24j'] > 3 * float(g[f'c{c}_log_loss_pose24j'])             # the scaled joints are off the 2-D poses: the term is large
    for nm in ('poses_T', 'poses_smpl', 'betas', 'xscale'):
        ref = ograds[nm].reshape(grads[nm].shape)
        assert np.abs(grads[nm] - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, (nm, np.abs(grads[nm] - ref).max(), np.abs(ref).max())
    opt.ctx.close()


def test_optimizer_updates_match_torch(c1, L):
    """Fused RMSprop / Adam steps vs torch.optim with identical gradients."""
    import torch
    opt, g, data, meta = c1
    N, T = meta[0], meta[1]
    ctx, st = opt.ctx, opt._stream()
    rng = np.random.default_rng(5)
    p0 = rng.normal(0, 1, (T, N, 72)).astype(np.float32)
    ctx.set_param(L.P_POSES_SMPL, p0, st)
    ctx.call('mh_reset_optimizer', st)
    tp = torch.tensor(p0.copy(), requires_grad=True)
    ropt = torch.optim.RMSprop([tp], lr=0.01, alpha=0.5, momentum=0.9)
    view = opt._view(L.BUF_GRADS)
    off = T * N * 3
    lr = 0.01
    for it in range(5):
        gr = (rng.normal(0, 1, p0.shape) * 10.0 ** rng.integers(-4, 1)).astype(np.float32)
        view.zero_()
        view[off:off + gr.size].copy_(torch.from_numpy(gr.reshape(-1)).to(view.device))
        ctx.call('mh_fit_update', lr, st)
        tp.grad = torch.from_numpy(gr.copy())
        for grp in ropt.param_groups:
            grp['lr'] = lr
        ropt.step()
        lr *= 0.99
        ours = ctx.get_param(L.P_POSES_SMPL, p0.shape)
        assert np.abs(ours - tp.detach().numpy()).max() <= 1e-6 * np.abs(tp.detach().numpy()).max()
    # Adam(lr .5, betas (.5, .5), eps 1e-6) on the translations
    t0 = rng.normal(0, 1, (T, N, 3)).astype(np.float32)
    ctx.set_param(L.P_POSES_T, t0, st)
    ctx.call('mh_init_begin', L.ptr(L.f32(data['pose2d'])), L.ptr(L.f32(data['poses_smpl'])), L.ptr(L.f32(data['betas_smpl'])), 0.15, st)
    ctx.set_param(L.P_POSES_T, t0, st)
    tt = torch.tensor(t0.copy(), requires_grad=True)
    aopt = torch.optim.Adam([tt], lr=0.5, betas=(0.5, 0.5), eps=1e-6)
    lr = 0.5
    for it in range(5):
        gr = rng.normal(0, 1e-2, t0.shape).astype(np.float32)
        view.zero_()
        view[:gr.size].copy_(torch.from_numpy(gr.reshape(-1)).to(view.device))
        ctx.call('mh_init_update', lr, it + 1, st)
        tt.grad = torch.from_numpy(gr.copy())
        for grp in aopt.param_groups:
            grp['lr'] = lr
        aopt.step()
        lr *= 0.95
        ours = ctx.get_param(L.P_POSES_T, t0.shape)
        assert np.abs(ours - tt.detach().numpy()).max() <= 2e-6 * np.abs(tt.detach().numpy()).max()


@pytest.mark.parametrize('knn_q', [1, 2, 8])
def test_contact_knn_against_brute_force(pkg, L, knn_q, monkeypatch):
    """Streaming exact top-32 (no distance matrix) vs numpy on a 50k-point cloud with duplicated points (ties); 1, 2 and 8
    person-frames per CTA (the grid-filling rule picks 8 only at benchmark sizes: MH_KNN_Q forces it here)."""
    monkeypatch.setenv('MH_KNN_Q', str(knn_q))
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch = meta[:5]
    opt = gh.make_optimizer(pkg, g, data, meta, coefs=dict(gh.COEFS, depth=0.0, silhouette=0.0), max_scene_points=60000)
    gh.prepare(opt, g, data, meta)
    rng = np.random.default_rng(3)
    cloud = np.stack([rng.uniform(-3, 3, 50000), rng.uniform(0.8, 1.2, 50000), rng.uniform(1, 8, 50000)], -1).astype(np.float32)
    cloud[1000:1040] = cloud[1000]                                             # 40 identical points
    c = 31
    ctx, st = opt.ctx, opt._stream()
    for which, key in ((L.P_POSES_T, 'poses_T'), (L.P_POSES_SMPL, 'poses_smpl'), (L.P_BETAS, 'betas'), (L.P_XSCALE, 'xscale')):
        ctx.set_param(which, g[f'c{c}_p_{key}'], st)
    ctx.set_param(L.P_ZMIN_LIN, g[f'c{c}_p_zmin_lin'], st); ctx.set_param(L.P_ZMAX_LIN, g[f'c{c}_p_zmax_lin'], st)
    opt.set_scene_pcd(cloud)
    ctx.call('mh_fit_grads', 0, 0, st)
    losses = ctx.read_losses(st)
    verts = opt._view(L.BUF_VERTS).view(T + 2, N, L.LD3V)[1:T + 1, :, :3 * L.V].cpu().numpy().reshape(T, N, L.V, 3)
    low = verts[np.arange(T)[:, None], np.arange(N)[None], np.argmax(verts[..., 1], axis=2)]           # (T, N, 3)
    ref = 0.0
    for t in range(T):
        for n in range(N):
            d2 = ((cloud - low[t, n]) ** 2).sum(1)
            nn = np.argsort(d2, kind='stable')[:32]
            cdv = cloud[nn, 1].mean() - low[t, n, 1]
            ref += abs(cdv + 0.02)
    assert abs(float(losses[L.L_CONTACT]) - ref) <= 1e-5 * ref


def test_contact_knn_one_million_points_with_ties(pkg, L):
    """BASELINE config C5's cloud size: 1 000 000 scene points, with MORE duplicates than the candidate list holds (700 > KNN_CAP = 512)
    planted exactly at the 32-nearest boundary of one body -- the bound must then be bisected on (distance, index) -- and 10 duplicates
    nearer than them; vs numpy brute force with the reference's arithmetic (``optimizer.py:492-500``)."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch = meta[:5]
    M = 1000000
    opt = gh.make_optimizer(pkg, g, data, meta, coefs=dict(gh.COEFS, depth=0.0, silhouette=0.0), max_scene_points=M)
    gh.prepare(opt, g, data, meta)
    c = 31
    ctx, st = opt.ctx, opt._stream()
    for which, key in ((L.P_POSES_T, 'poses_T'), (L.P_POSES_SMPL, 'poses_smpl'), (L.P_BETAS, 'betas'), (L.P_XSCALE, 'xscale')):
        ctx.set_param(which, g[f'c{c}_p_{key}'], st)
    ctx.set_param(L.P_ZMIN_LIN, g[f'c{c}_p_zmin_lin'], st); ctx.set_param(L.P_ZMAX_LIN, g[f'c{c}_p_zmax_lin'], st)
    rng = np.random.default_rng(11)
    cloud = np.stack([rng.uniform(-6, 6, M), rng.uniform(0.8, 1.2, M), rng.uniform(1, 10, M)], -1).astype(np.float32)
    opt.set_scene_pcd(cloud)
    ctx.call('mh_fit_grads', 0, 0, st)                                            # first: where the lowest vertices are
    verts = opt._view(L.BUF_VERTS).view(T + 2, N, L.LD3V)[1:T + 1, :, :3 * L.V].cpu().numpy().reshape(T, N, L.V, 3)
    low = verts[np.arange(T)[:, None], np.arange(N)[None], np.argmax(verts[..., 1], axis=2)]           # (T, N, 3)
    q = low[1, 0]
    idx = rng.permutation(M)
    cloud[idx[:10]] = q + np.array([0.0005, 0.0030, 0.0], np.float32)             # 10 copies, nearer than every random point
    cloud[idx[10:710]] = q + np.array([0.0, -0.0040, 0.0010], np.float32)         # 700 copies at the 32-nearest boundary (different y)
    opt.set_scene_pcd(cloud)
    ctx.call('mh_fit_grads', 0, 0, st)
    losses = ctx.read_losses(st)
    ref = 0.0
    for t in range(T):
        for n in range(N):
            d2 = ((cloud - low[t, n]) ** 2).sum(1)
            nn = np.argsort(d2, kind='stable')[:32]
            if t == 1 and n == 0:
                assert np.isin(nn, idx[:710]).all() and np.isin(nn, idx[:10]).sum() == 10                # the planted points are the neighbours
            cdv = cloud[nn, 1].mean() - low[t, n, 1]
            ref += abs(cdv + 0.02)
    assert abs(float(losses[L.L_CONTACT]) - ref) <= 1e-5 * ref
    opt.ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize('kind', ['floor', 'deep', 'far', 'sparse', 'thin'])
def test_contact_grid_equals_streaming_search(pkg, L, kind, monkeypatch):
    """The uniform-grid search of the contact term (a warp per person-frame, shells of cells around the lowest vertex) against the
    streaming search over the whole cloud: same 32 neighbours in the same (distance, index) order, so the loss and every gradient are
    BITWISE equal.  'deep': the floor 1.5 m below the feet, out of reach of the fine grid (answered by a coarser level); 'far': no scene
    point within the shells any level visits (every person-frame falls back to the streaming kernel);
    'sparse': 40 points in all; 'thin': all points on one line (a degenerate bounding box)."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch = meta[:5]
    opt = gh.make_optimizer(pkg, g, data, meta, coefs=dict(gh.COEFS, depth=0.0, silhouette=0.0), max_scene_points=300000)
    gh.prepare(opt, g, data, meta)
    rng = np.random.default_rng(5)
    if kind == 'floor':
        M = 300000
        cloud = np.stack([rng.uniform(-4, 4, M), 1.0 + 0.02 * rng.standard_normal(M), rng.uniform(1, 9, M)], -1)
        cloud[5000:5100] = cloud[5000]
    elif kind == 'deep':
        M = 300000
        cloud = np.stack([rng.uniform(-4, 4, M), 2.5 + 0.02 * rng.standard_normal(M), rng.uniform(1, 9, M)], -1)
    elif kind == 'far':
        M = 100000
        cloud = np.stack([rng.uniform(40, 44, M), rng.uniform(0.8, 1.2, M), rng.uniform(60, 64, M)], -1)
    elif kind == 'sparse':
        cloud = np.stack([rng.uniform(-2, 2, 40), rng.uniform(0.5, 1.5, 40), rng.uniform(2, 6, 40)], -1)
    else:
        M = 20000
        cloud = np.stack([np.zeros(M), np.ones(M), rng.uniform(1, 9, M)], -1)
    cloud = cloud.astype(np.float32)
    ctx, st = opt.ctx, opt._stream()
    c = 31
    for which, key in ((L.P_POSES_T, 'poses_T'), (L.P_POSES_SMPL, 'poses_smpl'), (L.P_BETAS, 'betas'), (L.P_XSCALE, 'xscale')):
        ctx.set_param(which, g[f'c{c}_p_{key}'], st)
    ctx.set_param(L.P_ZMIN_LIN, g[f'c{c}_p_zmin_lin'], st); ctx.set_param(L.P_ZMAX_LIN, g[f'c{c}_p_zmax_lin'], st)
    out = {}
    for grid in ('1', '0'):
        monkeypatch.setenv('MH_KNN_GRID', grid)
        opt.set_scene_pcd(cloud)                                                   # (re)builds the grid, or drops it
        ctx.call('mh_fit_grads', 0, 0, st)
        out[grid] = (ctx.read_losses(st).copy(), opt._view(L.BUF_GRADS).cpu().numpy().copy())
        stats = np.zeros(4, np.int64)
        ctx.call('mh_debug_knn_stats', L.ptr(stats), st)
        if grid == '0':
            assert stats[2] == 0                                                   # no grid: the streaming kernel alone
        else:
            assert stats[2] == len(cloud) and stats[1] >= 1
            assert stats[0] == (0 if kind == 'far' else T * N), stats              # who answered
    assert np.isfinite(out['1'][0]).all() and out['1'][0][L.L_CONTACT] > 0
    assert np.array_equal(out['1'][0].view(np.uint32), out['0'][0].view(np.uint32))
    assert np.array_equal(out['1'][1].view(np.uint32), out['0'][1].view(np.uint32))
    opt.ctx.close()


def _two_shard_cycle(pkg, L, g, data, meta, cycle, refresh=False):
    """World-size-2 frame sharding emulated on ONE GPU: two contexts (rank 0 / 1), halo frames and the shared gradient block
    exchanged through the host exactly as sharding.exchange_halo / allreduce_shared do over NCCL.  ``refresh``: the filtered
    vertices come from the sharded One-Euro refresh itself (``optimizer._refresh_filters``: rank 0 scans first, hands its filter
    state to rank 1, then the filtered boundary frames are swapped) instead of being injected; returns them as a third value."""
    import torch
    N, T, W, H, batch = meta[:5]
    sh = sys.modules[pkg.__name__ + '.sharding']
    opts = []
    for r in range(2):
        o = gh.make_optimizer(pkg, g, data, meta)
        o.rank, o.world = r, 2
        gh.prepare(o, g, data, meta)
        opts.append(o)
    assert opts[0].t1 == opts[1].t0 and opts[0].t1 % batch == 0
    halos = []
    for o in opts:
        ctx, st, c = o.ctx, o._stream(), cycle
        sl = slice(o.t0, o.t1)
        ctx.set_param(L.P_POSES_T, g[f'c{c}_p_poses_T'][sl], st); ctx.set_param(L.P_POSES_SMPL, g[f'c{c}_p_poses_smpl'][sl], st)
        ctx.set_param(L.P_BETAS, g[f'c{c}_p_betas'], st); ctx.set_param(L.P_BETAS_REF, g['init_betas'], st)
        ctx.set_param(L.P_ZMIN_LIN, g[f'c{c}_p_zmin_lin'][sl], st); ctx.set_param(L.P_ZMAX_LIN, g[f'c{c}_p_zmax_lin'][sl], st)
        ctx.set_param(L.P_XSCALE, g[f'c{c}_p_xscale'], st)
        if len(g[f'c{c}_scene_pcd']):
            o.set_scene_pcd(g[f'c{c}_scene_pcd'])
        if c >= 50 and not refresh:
            gh.set_filtered(o, g['verts_filtered'])
        ctx.call('mh_halo_pack', st)
        halos.append(o._view(L.BUF_HALO_SEND).view(2, -1).clone())
    filtered = None
    if refresh:
        a, b = opts
        a.ctx.call('mh_refresh_filters', 0.01, 0.02, 0.001, 0.5, 25.0, 1, a._stream())                   # rank 0: first = 1
        torch.cuda.synchronize()
        b._view(L.BUF_CARRY_IN).copy_(a._view(L.BUF_CARRY_OUT))                                          # sharding.send_carry / pass_carry
        b.ctx.call('mh_refresh_filters', 0.01, 0.02, 0.001, 0.5, 25.0, 0, b._stream())                   # rank 1 continues the scan
        torch.cuda.synchronize()
        Fa = a._view(L.BUF_FILTERED).view(a.T_local + 2, -1)
        Fb = b._view(L.BUF_FILTERED).view(b.T_local + 2, -1)
        Fa[a.T_local + 1].copy_(Fb[1])                                                                   # filtered halo frames
        Fb[0].copy_(Fa[a.T_local])
        filtered = torch.cat([Fa[1:a.T_local + 1], Fb[1:b.T_local + 1]]).clone()
    opts[0]._view(L.BUF_HALO_RECV).view(2, -1)[1].copy_(halos[1][0])           # rank 0 <- rank 1's first frame
    opts[1]._view(L.BUF_HALO_RECV).view(2, -1)[0].copy_(halos[0][1])           # rank 1 <- rank 0's last frame
    opts[0].ctx.call('mh_fit_grads', 0, 1, opts[0]._stream())
    opts[1].ctx.call('mh_fit_grads', 1, 0, opts[1]._stream())
    torch.cuda.synchronize()
    shared = opts[0]._view(L.BUF_SHARED) + opts[1]._view(L.BUF_SHARED)
    for o in opts:
        o._view(L.BUF_SHARED).copy_(shared)
    losses = opts[0].ctx.read_losses(opts[0]._stream())
    log = sh.log_from_loss_block(losses, (T + batch - 1) // batch)
    cat = lambda which, shape_fn: np.concatenate([o.ctx.get_grad(which, shape_fn(o.T_local)) for o in opts], 0)
    grads = {'poses_T': cat(L.P_POSES_T, lambda t: (t, N, 1, 3)), 'poses_smpl': cat(L.P_POSES_SMPL, lambda t: (t, N, 72)),
             'zmin_lin': cat(L.P_ZMIN_LIN, lambda t: (t, 1, 1)), 'zmax_lin': cat(L.P_ZMAX_LIN, lambda t: (t, 1, 1)),
             'betas': opts[0].ctx.get_grad(L.P_BETAS, (1, N, 10)), 'xscale': opts[0].ctx.get_grad(L.P_XSCALE, (1, N, 1, 1))}
    for o in opts:
        o.ctx.close()
    if refresh:
        return log, grads, filtered
    return log, grads


@pytest.mark.parametrize('cycle', [31, 51])
def test_two_shards_equal_one(pkg, L, cycle):
    """Frame sharding + 1-frame halo + shared-gradient sum reproduces the unsharded reference cycle."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    log, grads = _two_shard_cycle(pkg, L, g, data, meta, cycle)
    for k, v in log.items():
        ref = float(g[f'c{cycle}_log_{k}'])
        assert abs(v - ref) <= 1e-4 * abs(ref) + 1e-9, (k, v, ref)
    for nm, gr in grads.items():
        ref = g[f'c{cycle}_g_{nm}'].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, nm


def test_sharded_filter_refresh_equals_single(pkg, L):
    """The One-Euro refresh run shard after shard with the filter state handed over (``CARRY_OUT`` -> ``CARRY_IN``, ``first = 0``)
    gives the SAME filtered vertices, bit for bit, as one context scanning the whole sequence; with the filtered boundary frames
    swapped the cycle-50 gradients then reproduce the unsharded reference cycle."""
    import torch
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch = meta[:5]
    c = 50
    one = gh.make_optimizer(pkg, g, data, meta)
    gh.prepare(one, g, data, meta)
    ctx, st = one.ctx, one._stream()
    ctx.set_param(L.P_POSES_T, g[f'c{c}_p_poses_T'], st); ctx.set_param(L.P_POSES_SMPL, g[f'c{c}_p_poses_smpl'], st)
    ctx.set_param(L.P_BETAS, g[f'c{c}_p_betas'], st); ctx.set_param(L.P_XSCALE, g[f'c{c}_p_xscale'], st)
    ctx.call('mh_refresh_filters', 0.01, 0.02, 0.001, 0.5, 25.0, 1, st)
    torch.cuda.synchronize()
    F1 = one._view(L.BUF_FILTERED).view(T + 2, -1)[1:T + 1].clone()
    one.ctx.close()
    log, grads, F2 = _two_shard_cycle(pkg, L, g, data, meta, c, refresh=True)
    assert torch.equal(F1, F2)                                                   # bit for bit
    ref_f = torch.from_numpy(g['verts_filtered'].reshape(T, N, -1)).to(F1.device)
    assert float((F1.view(T, N, -1)[..., :3 * L.V] - ref_f).abs().max()) < 2e-6  # metres, vs the reference's filtered vertices
    for k, v in log.items():
        ref = float(g[f'c{c}_log_{k}'])
        assert abs(v - ref) <= 1e-4 * abs(ref) + 1e-9, (k, v, ref)
    for nm, gr in grads.items():
        ref = g[f'c{c}_g_{nm}'].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, nm


def test_error_paths(pkg, L):
    g, data, meta = gh.load_fit('fit_c1.npz')
    N, T, W, H, batch = meta[:5]
    with pytest.raises(L.MhError):
        L.Context(T, 40, H, W)                                                 # more than 32 persons
    with pytest.raises(L.MhError):
        L.Context(3, N, H, W, B=2, t0=1, T_total=4)                            # shard start not on a batch edge
    opt = gh.make_optimizer(pkg, g, data, meta)
    gh.prepare(opt, g, data, meta, ingest=False)
    with pytest.raises(L.MhError):
        opt.ctx.call('mh_fit_grads', 0, 0, opt._stream())                      # fit before ingest
    bad = dict(data)
    bad['seg_mask'] = data['seg_mask'] * 0.5                                   # non-binary instance masks
    with pytest.raises(L.MhError):
        opt._ingest(gh.ListLoader(bad, batch))
    with pytest.raises(L.MhError):
        opt.set_scene_pcd(np.zeros((5, 3), np.float32))                        # fewer than 32 scene points
    with pytest.raises(L.MhError):
        opt.set_scene_pcd(np.zeros((W * H + 1, 3), np.float32))                # beyond M_max
    with pytest.raises(L.MhError):
        opt.ctx.get_param(L.P_BETAS, (N, 11))                                  # wrong size


def test_empty_and_degenerate_frames(pkg, L):
    """Persons without a mask / without confident joints / outside the image contribute exactly what the reference's
    validity gates say (optimizer.py:404-409, 436-438, 472)."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch = meta[:5]
    d2 = {k: v.copy() for k, v in data.items()}
    d2['seg_mask'][:, 1] = 0                                                   # person 1 has no mask at all
    d2['pose2d'][0, 0, :, 2] = 0.1                                             # person 0 has no confident joint in frame 0
    opt = gh.make_optimizer(pkg, g, d2, meta)
    log, grads = gh.teacher_forced_cycle(opt, g, d2, meta, 31)
    import torch
    from oracle import fit_ref, synth
    model = synth.load_model_tensors(gh.model_dir())
    fr = fit_ref.FitRef(model, (W, H), T, g['cam_K'], gh.COEFS)
    c = 31
    fr.set_variables(g[f'c{c}_p_poses_T'], g[f'c{c}_p_poses_smpl'], g[f'c{c}_p_betas'], d2['valid_smpl'], g[f'c{c}_p_zmin_lin'],
                     g[f'c{c}_p_zmax_lin'], g[f'c{c}_p_xscale'])
    fr.betas_ref = torch.from_numpy(g['init_betas'])
    fr.set_scene_pcd(g[f'c{c}_scene_pcd'])
    olog, _ = fr.cycle_grads(d2, [np.arange(s, min(s + batch, T)) for s in range(0, T, batch)])
    ograds = {nm: p.grad.numpy() for nm, p in zip(gh.NAMES, fr.leaves())}
    for k, v in log.items():
        assert abs(v - olog[k]) <= 1e-4 * abs(olog[k]) + 1e-9, (k, v, olog[k])
    for nm, gr in grads.items():
        ref = ograds[nm].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, nm
    # a body far outside the frustum renders nothing and must not fault
    ctx, st = opt.ctx, opt._stream()
    far = g[f'c{c}_p_poses_T'].copy()
    far[:, 0, :, 0] += 50.0
    ctx.set_param(L.P_POSES_T, far, st)
    ctx.call('mh_fit_grads', 0, 0, st)
    assert np.all(np.isfinite(ctx.read_losses(st)))


def _median_inputs(meta, seed=0):
    """Random background masks / images with never-background pixels, odd and even counts and duplicated values."""
    N, T, W, H = meta[:4]
    rng = np.random.default_rng(seed)
    back = (rng.random((T, H, W)) > 0.35).astype(np.uint8)
    back[:, 3:7, 5:9] = 0                                                     # never background
    back[1:, 10, :] = 0                                                        # a single sample
    back[:, 12, :] = 1                                                         # all frames (even count for T = 4)
    images = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    images[:, 12, :8] = images[0, 12, :8]                                      # duplicates
    return back, images


def test_scene_median_on_device(n2, pkg, L):
    """Exact radix-selection median (csrc/mh_scene.cu) vs np.ma.median (fhsog.py:180-202) on the same per-frame depths."""
    import torch
    opt, g, data, meta = n2
    sc = sys.modules[pkg.__name__ + '.scene']
    N, T, W, H = meta[:4]
    ctx, st = opt.ctx, opt._stream()
    c = 31
    ctx.set_param(L.P_ZMIN_LIN, g[f'c{c}_p_zmin_lin'], st); ctx.set_param(L.P_ZMAX_LIN, g[f'c{c}_p_zmax_lin'], st)
    back, images = _median_inputs(meta)
    ctx.call('mh_scene_set_back', 0, T, L.ptr(back), L.ptr(images), st)
    opt._have_images = True
    depths = np.empty((T, H, W), np.float32)
    ctx.call('mh_scene_depths', 0, T, L.ptr(depths))
    ref_img, ref_depth, ref_mask = sc.aggregate_scene_geometry_median(depths, images, back.astype(np.float32))
    depth, mask = opt._device_median(0)
    img = opt._device_median(1)
    assert np.array_equal(mask, ref_mask)
    assert np.array_equal(depth, ref_depth)                                    # bit-exact, incl. 0 where never background
    assert np.array_equal(img, ref_img)
    # frame-sharded: two contexts with half the frames each, histograms summed between passes as the NCCL all-reduce does
    halves = []
    for r in range(2):
        o = gh.make_optimizer(pkg, g, data, meta)
        o.rank, o.world = r, 2
        gh.prepare(o, g, data, meta)
        sl = slice(o.t0, o.t1)
        o.ctx.set_param(L.P_ZMIN_LIN, g[f'c{c}_p_zmin_lin'][sl], o._stream()); o.ctx.set_param(L.P_ZMAX_LIN, g[f'c{c}_p_zmax_lin'][sl], o._stream())
        o.ctx.call('mh_scene_set_back', 0, o.T_local, L.ptr(np.ascontiguousarray(back[sl])), L.ptr(np.ascontiguousarray(images[sl])), o._stream())
        halves.append(o)
    HW = H * W
    for which, npass, planes in ((0, 10, 1), (1, 4, 3)):
        for p in range(npass):
            for o in halves:
                o.ctx.call('mh_scene_median_pass', which, p, o._stream())
            torch.cuda.synchronize()
            if p < npass - 1:
                n = (1 if p == 0 else 16 * planes) * HW
                tot = halves[0]._view(L.BUF_MEDIAN_HIST)[:n] + halves[1]._view(L.BUF_MEDIAN_HIST)[:n]
                for o in halves:
                    o._view(L.BUF_MEDIAN_HIST)[:n].copy_(tot)
            else:
                a0, a1 = halves[0]._view(L.BUF_MEDIAN_AUX), halves[1]._view(L.BUF_MEDIAN_AUX)
                le = a0[:planes * HW] + a1[:planes * HW]
                ab = torch.minimum(a0[3 * HW:(3 + planes) * HW], a1[3 * HW:(3 + planes) * HW])
                for o in halves:
                    o._view(L.BUF_MEDIAN_AUX)[:planes * HW].copy_(le)
                    o._view(L.BUF_MEDIAN_AUX)[3 * HW:(3 + planes) * HW].copy_(ab)
            torch.cuda.synchronize()
        for o in halves:
            if which == 0:
                d2 = np.empty((H, W), np.float32); m2 = np.empty((H, W), np.uint8)
                o.ctx.call('mh_scene_median_finish', 0, L.ptr(d2), L.ptr(m2), None, o._stream())
                assert np.array_equal(d2, ref_depth) and np.array_equal(m2.astype(bool), ref_mask)
            else:
                i2 = np.empty((H, W, 3), np.uint8)
                o.ctx.call('mh_scene_median_finish', 1, None, None, L.ptr(i2), o._stream())
                assert np.array_equal(i2, ref_img)
    for o in halves:
        o.ctx.close()


def test_full_fit_runs_like_the_reference(pkg, L):
    """The whole drop-in call sequence on config C1 (init 30 iterations, fit 52 cycles with scene updates at cycle >= 30,
    filters at cycle 50): output keys / shapes of optimizer.py:619-636 and loss levels of the reference's run.  Trajectories
    are chaotic, so only levels are compared."""
    g, data, meta = gh.load_fit('fit_c1.npz')
    N, T, W, H, batch, num_iter, init_iter = meta
    opt = gh.make_optimizer(pkg, g, data, meta)
    opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=init_iter, batch_size=batch)
    log = opt.fit(gh.ListLoader(data, batch), num_iter=num_iter)
    assert len(log) == num_iter and list(log[0]) == ['loss_pose24j', 'loss_depth', 'loss_silhouette', 'reg_ref_poses', 'reg_scale',
                                                      'reg_contact', 'reg_foot_sliding', 'reg_vel', 'reg_filter_verts']
    for k in ('loss_pose24j', 'loss_depth', 'loss_silhouette'):
        assert abs(log[0][k] - g['log_' + k][0]) <= 1e-4 * abs(g['log_' + k][0])                # the first cycle is exact
        assert log[-1][k] <= 3.0 * g['log_' + k][-1] + 1e-6                                        # same level at the end
    assert log[29]['reg_contact'] == 0 and log[31]['reg_contact'] > 0 and log[49]['reg_filter_verts'] == 0 and log[50]['reg_filter_verts'] > 0
    v = opt.get_optimized_variables()
    for k in ('scale_factor', 'poses_T', 'poses_smpl', 'betas_smpl', 'valid_smpl', 'min_z', 'max_z', 'scene_depth', 'scene_img', 'scene_mask'):
        assert v[k].shape == g['final_' + k].shape, k
    assert np.abs(v['poses_T'] - g['final_poses_T']).max() < 0.15                                 # metres; oracle-vs-reference spread is 0.035
    assert v['scene_mask'].min() == 1


def _oracle_cycle_custom(g, data, meta, cycle, T_use=None, **kw):
    import torch
    from oracle import fit_ref, synth
    N, T, W, H, batch = meta[:5]
    T_use = T if T_use is None else T_use
    model = synth.load_model_tensors(gh.model_dir())
    fr = fit_ref.FitRef(model, (W, H), T_use, g['cam_K'], gh.COEFS, **kw)
    c = cycle
    fr.set_variables(g[f'c{c}_p_poses_T'][:T_use], g[f'c{c}_p_poses_smpl'][:T_use], g[f'c{c}_p_betas'], data['valid_smpl'][:T_use],
                     g[f'c{c}_p_zmin_lin'][:T_use], g[f'c{c}_p_zmax_lin'][:T_use], g[f'c{c}_p_xscale'])
    fr.betas_ref = torch.from_numpy(g['init_betas'])
    if len(g[f'c{c}_scene_pcd']):
        fr.set_scene_pcd(g[f'c{c}_scene_pcd'])
    d = {k: v[:T_use] for k, v in data.items()}
    log, _ = fr.cycle_grads(d, [np.arange(s, min(s + batch, T_use)) for s in range(0, T_use, batch)])
    return log, {nm: p.grad.numpy().copy() for nm, p in zip(gh.NAMES, fr.leaves())}


def test_distortion_and_joint_weights(pkg, L):
    """camera_projection_torch with the 5 distortion coefficients (transforms.py:78-90, the code's own y term) and non-uniform
    pose17j_weights (optimizer.py:108-130, 420) vs the oracle."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch = meta[:5]
    Kd = np.array([0.1, 0.01, 0.001, 0.002, 0.0001], np.float32)
    w17 = np.linspace(0.5, 2.0, 17).astype(np.float32)
    opt = gh.make_optimizer(pkg, g, data, meta, cam_dist_coef=Kd, pose17j_weights=w17)
    log, grads = gh.teacher_forced_cycle(opt, g, data, meta, 31)
    olog, ograds = _oracle_cycle_custom(g, data, meta, 31, cam_dist_coef=Kd, pose17j_weights=w17)
    for k, v in log.items():
        assert abs(v - olog[k]) <= 1e-4 * abs(olog[k]) + 1e-9, (k, v, olog[k])
    for nm, gr in grads.items():
        ref = ograds[nm].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, nm
    assert abs(log['loss_pose24j'] - float(g['c31_log_loss_pose24j'])) > 1e-3 * float(g['c31_log_loss_pose24j'])      # the options matter


def test_single_frame_sequence(pkg, L):
    """num_frames = 1: no temporal pairs, no foot-sliding pairs (optimizer.py:177-179, 512-518, 560)."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch, num_iter, init_iter = meta
    d1 = {k: v[:1] for k, v in data.items()}
    m1 = (N, 1, W, H, 1, num_iter, init_iter)
    g1 = {k: (g[k][:1] if (k.startswith('c31_p_') and g[k].shape[0] == T) else g[k]) for k in g.files}
    opt = gh.make_optimizer(pkg, g, d1, m1)
    log, grads = gh.teacher_forced_cycle(opt, g1, d1, m1, 31)
    olog, ograds = _oracle_cycle_custom(g, data, (N, T, W, H, 1), 31, T_use=1)
    assert log['reg_vel'] == 0 and log['reg_foot_sliding'] == 0
    for k, v in log.items():
        assert abs(v - olog[k]) <= 1e-4 * abs(olog[k]) + 1e-9, (k, v, olog[k])
    for nm, gr in grads.items():
        ref = ograds[nm].reshape(gr.shape)
        assert np.abs(gr - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, nm
    y = opt.one_euro_filter(np.zeros((1, 5), np.float32), 0.01, 0.02)
    assert tuple(y.shape) == (1, 5)


def test_fixed_scale_factor_is_not_updated(pkg, L):
    """scale_factor given to init_optimized_variables -> xscale_factor is not an optimised leaf (optimizer.py:279-282, 350-353)."""
    g, data, meta = gh.load_fit('fit_c1.npz')
    N, T, W, H, batch = meta[:5]
    opt = gh.make_optimizer(pkg, g, data, meta)
    sf = np.full(N, 1.21, np.float32)
    opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], scale_factor=sf, num_iter=2,
                                 batch_size=batch)
    assert opt.optim_scale_factor is False
    x0 = opt.ctx.get_param(L.P_XSCALE, (N,))
    assert np.allclose(x0, 2.0, atol=1e-5)                                     # 1.1^2 = 1.21
    opt._ingest(gh.ListLoader(data, batch))
    st = opt._stream()
    opt.ctx.call('mh_reset_optimizer', st)
    p0 = opt.ctx.get_param(L.P_POSES_T, (T, N, 3))
    opt.ctx.call('mh_fit_grads', 0, 0, st)
    opt.ctx.call('mh_fit_update', 0.01, st)
    assert np.array_equal(opt.ctx.get_param(L.P_XSCALE, (N,)), x0)
    assert np.abs(opt.ctx.get_param(L.P_POSES_T, (T, N, 3)) - p0).max() > 0
    assert np.allclose(opt.get_optimized_variables()['scale_factor'].reshape(-1), 1.21, atol=1e-5)


def test_coarse_binning_and_capacity_errors(pkg, L):
    """Bodies whose tile count exceeds the bin table use coarser bins (same results); exceeding the tile-list or the
    depth-winner capacity is reported as MH_E_CAPACITY, never silently dropped."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    opt = gh.make_optimizer(pkg, g, data, meta)
    ref_log, ref_grads = gh.teacher_forced_cycle(opt, g, data, meta, 31)
    opt.ctx.call('mh_debug_set_render_caps', 2, 0, 0)                          # at most 2 bins per body -> binning granularity is coarsened
    log, grads = gh.teacher_forced_cycle(opt, g, data, meta, 31)
    for k in log:
        assert abs(log[k] - ref_log[k]) <= 1e-6 * abs(ref_log[k]) + 1e-12, k
    for nm in grads:
        assert np.abs(grads[nm] - ref_grads[nm]).max() <= 1e-5 * np.abs(ref_grads[nm]).max() + 1e-9, nm
    opt.ctx.call('mh_debug_set_render_caps', 0, 0, 16)                         # 16 depth winners per body
    with pytest.raises(L.MhError, match='capacity'):
        gh.teacher_forced_cycle(opt, g, data, meta, 31)
    opt2 = gh.make_optimizer(pkg, g, data, meta)
    gh.prepare(opt2, g, data, meta)
    opt2.ctx.call('mh_debug_set_render_caps', 0, 64, 0)                        # 64 tile-list entries per body
    with pytest.raises(L.MhError, match='capacity'):
        gh.teacher_forced_cycle(opt2, g, data, meta, 31)


class _PinnedLoader(object):
    """Batches as slices of PINNED host tensors (what a DataLoader with pin_memory=True hands over)."""
    def __init__(self, inputs, batch):
        import torch
        self.t = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in inputs.items()}
        self.batch, self.T = batch, len(inputs['idxs'])

    def __iter__(self):
        for s in range(0, self.T, self.batch):
            yield {k: v[s:s + self.batch] for k, v in self.t.items()}


@pytest.mark.parametrize('mode', ['default', 'host', 'device'])
def test_mask_ingest_paths_give_the_same_planes(pkg, L, mode, monkeypatch):
    """float32 instance masks reach the device two ways -- packed to bit planes by the host cores (one process per host) or copied
    full-size and packed on the device (one process per GPU) -- from pageable or pinned memory; the planes, the losses and the
    gradients of a cycle do not depend on the way."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch = meta[:5]
    out = {}
    for which in ('reference', mode):
        if which == 'reference':
            monkeypatch.setenv('MH_INGEST_HOST_PACK', '1')                     # pageable source, host packing
            loader = gh.ListLoader(data, 2)
        else:
            if mode == 'default':
                monkeypatch.delenv('MH_INGEST_HOST_PACK', raising=False)
            else:
                monkeypatch.setenv('MH_INGEST_HOST_PACK', '1' if mode == 'host' else '0')
            loader = _PinnedLoader(data, 2)
        opt = gh.make_optimizer(pkg, g, data, meta)
        gh.prepare(opt, g, data, meta, ingest=False)
        opt._ingest(loader)
        dep = np.zeros((T, H, W), np.float32); seg = np.zeros((T, N, H, W), np.float32)
        opt.ctx.call('mh_read_planes', 0, T, L.ptr(dep), L.ptr(seg))
        log, grads = gh.teacher_forced_cycle(opt, g, data, meta, 31)           # (ingested already: the planes above are used)
        out[which] = (dep, seg, log, grads)
        opt.ctx.close()
    assert np.array_equal(out['reference'][0], out[mode][0]) and np.array_equal(out['reference'][1], out[mode][1])
    assert np.array_equal(out[mode][1], data['seg_mask'].astype(np.float32))
    assert out[mode][2] == out['reference'][2]
    for nm in out[mode][3]:
        assert np.array_equal(out[mode][3][nm], out['reference'][3][nm]), nm


def test_sparse_joints_key_selects_the_regressor(pkg, L):
    """``smpl_sparse_joints_key`` (``optimizer.py:40, 696, 750``): the 2-D terms use the 17 joints of the chosen output of the SMPL
    layer -- 'joints_alphapose' (default) or 'joints_h36m17' (H36M regressor in the layer's row order, relative to its pelvis joint 14:
    ``smpl.py:240-242, 369-373``)."""
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch = meta[:5]
    with pytest.raises(ValueError, match='smpl_sparse_joints_key'):
        gh.make_optimizer(pkg, g, data, meta, smpl_sparse_joints_key='joints_smpl24')
    out = {}
    for key in ('joints_alphapose', 'joints_h36m17'):
        opt = gh.make_optimizer(pkg, g, data, meta, smpl_sparse_joints_key=key)
        gh.prepare(opt, g, data, meta, ingest=False)
        betas = np.ascontiguousarray(np.tile(g['init_betas'].reshape(1, N, 10), (T, 1, 1)).reshape(T * N, 10), np.float32)
        poses = np.ascontiguousarray(data['poses_smpl'].reshape(T * N, 72), np.float32)
        verts, joints = opt.smpl_forward(betas, poses)
        reg = opt.model['J_regressor_alphapose' if key == 'joints_alphapose' else 'J_regressor_h36m17']
        ref = np.einsum('jv,bvk->bjk', reg.astype(np.float64), np.asarray(verts, np.float64).reshape(T * N, L.V, 3))
        if key == 'joints_h36m17':
            ref = ref - ref[:, 14:15]                                          # as the layer returns them
        assert np.abs(np.asarray(joints).reshape(T * N, 17, 3) - ref).max() <= 5e-6
        out[key] = np.asarray(joints).copy()
        opt.ctx.close()
    assert np.abs(out['joints_alphapose'] - out['joints_h36m17']).max() > 1e-3       # different regressors, different joints

"""Development aid: run the CUDA path on the golden inputs and PRINT the differences (no asserts)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import __graft_entry__ as ge
import gpu_harness as gh

pkg = ge.load_package()
L = sys.modules[pkg.__name__ + '._lib']


def section(name):
    print('\n==== ' + name, flush=True)


def smpl_kat():
    section('SMPL forward vs reference KAT')
    kat = np.load(os.path.join(gh.GOLDEN, 'kat_functions.npz'))
    g, data, meta = gh.load_fit('fit_c1.npz')
    opt = gh.make_optimizer(pkg, g, data, meta)
    gh.prepare(opt, g, data, meta, ingest=False)
    v, j = opt.smpl_forward(kat['smpl_betas'], kat['smpl_poses'])
    print('verts max abs diff', np.abs(v - kat['smpl_verts']).max(), ' j17', np.abs(j - kat['smpl_joints_alphapose']).max())
    y = opt.one_euro_filter(kat['oef2_in'], 0.001, 0.5).cpu().numpy()
    print('one-euro max diff', np.abs(y - kat['oef2_out']).max(), np.abs(opt.one_euro_filter(kat['oef_in'], 0.01, 0.02).cpu().numpy() - kat['oef_out']).max())


def render_check(name, c=0):
    section(f'render planes vs oracle raster ({name}, cycle {c})')
    import torch
    from oracle import raster, refmath as rm, synth
    g, data, meta = gh.load_fit(name)
    N, T, W, H, batch, _, _ = meta
    opt = gh.make_optimizer(pkg, g, data, meta)
    gh.prepare(opt, g, data, meta)
    ctx, st = opt.ctx, opt._stream()
    ctx.set_param(L.P_POSES_T, g[f'c{c}_p_poses_T'], st); ctx.set_param(L.P_POSES_SMPL, g[f'c{c}_p_poses_smpl'], st)
    ctx.set_param(L.P_BETAS, g[f'c{c}_p_betas'], st); ctx.set_param(L.P_XSCALE, g[f'c{c}_p_xscale'], st)
    model = synth.load_model_tensors(gh.model_dir())
    mt = {a: (torch.from_numpy(v) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v) for a, v in model.items()}
    mt['parents'] = [int(p) for p in model['parents']]
    faces = torch.from_numpy(model['faces'].astype(np.int64))
    Kndc = torch.from_numpy(rm.compute_calibration_matrix(1.0, 100.0, g['cam_K'], (W, H)))
    for t in range(min(T, 2)):
        for n in range(N):
            zb = np.zeros((H, W), np.float32); al = np.zeros((H, W), np.float32)
            ctx.call('mh_debug_render', t, n, L.ptr(zb), L.ptr(al))
            with torch.no_grad():
                out = rm.smpl_forward(mt, torch.from_numpy(g[f'c{c}_p_betas'][0, n:n + 1]), torch.from_numpy(g[f'c{c}_p_poses_smpl'][t, n:n + 1]))
                s = float(1.1 ** g[f'c{c}_p_xscale'].reshape(-1)[n])
                va = s * out['verts'][0] + torch.from_numpy(g[f'c{c}_p_poses_T'][t, n])
                z0, a0 = raster.render_person(va, faces, Kndc, H, W)
            z0, a0 = z0.numpy(), a0.numpy()
            cov = (z0 > 0)
            print(f't{t} n{n}: covered px oracle {cov.sum()} ours {(zb > 0).sum()} mismatch-cover {(cov != (zb > 0)).sum()} '
                  f'zbuf max diff {np.abs(zb - z0)[cov & (zb > 0)].max() if (cov & (zb > 0)).any() else -1:.3e} '
                  f'alpha max diff {np.abs(al - a0).max():.3e} n(alpha diff>1e-3) {(np.abs(al - a0) > 1e-3).sum()}')


def teacher(name, cycles=(0, 1, 30, 31, 50, 51)):
    g, data, meta = gh.load_fit(name)
    opt = gh.make_optimizer(pkg, g, data, meta)
    for c in cycles:
        section(f'teacher-forced {name} cycle {c}')
        try:
            log, grads = gh.teacher_forced_cycle(opt, g, data, meta, c)
        except Exception:
            traceback.print_exc()
            continue
        for k, v in log.items():
            ref = float(g[f'c{c}_log_{k}'])
            print(f'  {k:18s} ours {v: .8e} ref {ref: .8e} rel {abs(v - ref) / (abs(ref) + 1e-12):.2e}')
        for nm, gr in grads.items():
            ref = g[f'c{c}_g_{nm}'].reshape(gr.shape)
            print(f'  grad {nm:12s} max|ref| {np.abs(ref).max():.4e} max diff {np.abs(gr - ref).max():.4e} rel {np.abs(gr - ref).max() / (np.abs(ref).max() + 1e-20):.2e}')


def init_stage(name):
    section(f'init stage {name}')
    g, data, meta = gh.load_fit(name)
    N, T, W, H, batch, num_iter, init_iter = meta
    opt = gh.make_optimizer(pkg, g, data, meta)
    log = opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=init_iter, batch_size=batch)
    pT = opt.ctx.get_param(L.P_POSES_T, (T, N, 1, 3))
    l2d = np.array([float(l['loss_2d']) for l in log])
    print('poses_T max diff', np.abs(pT - g['init_poses_T']).max(), 'loss_2d rel diff', np.abs(l2d - g['init_loss_2d']).max() / g['init_loss_2d'].max())
    print('loss first/last', l2d[0], l2d[-1], 'ref', g['init_loss_2d'][0], g['init_loss_2d'][-1])


def full_fit(name):
    section(f'full fit {name}')
    g, data, meta = gh.load_fit(name)
    N, T, W, H, batch, num_iter, init_iter = meta
    opt = gh.make_optimizer(pkg, g, data, meta)
    opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=init_iter, batch_size=batch)
    t0 = time.time()
    log = opt.fit(gh.ListLoader(data, batch), num_iter=num_iter)
    print('fit wall', time.time() - t0, 's; launches', opt.ctx.launches())
    for k in log[0]:
        ours = np.array([l[k] for l in log]); ref = g['log_' + k]
        print(f'  {k:18s} c0 {ours[0]:.5e}/{ref[0]:.5e} c29 {ours[29]:.5e}/{ref[29]:.5e} c40 {ours[40]:.5e}/{ref[40]:.5e} last {ours[-1]:.5e}/{ref[-1]:.5e}')
    fv = opt.get_optimized_variables()
    for k in ('poses_T', 'poses_smpl', 'betas_smpl', 'scale_factor', 'min_z', 'max_z'):
        print(f'  final {k:12s} max diff {np.abs(fv[k] - g["final_" + k]).max():.4e}')


if __name__ == '__main__':
    which = sys.argv[1:] or ['smpl', 'render', 'teacher', 'init']
    for w in which:
        try:
            if w == 'smpl': smpl_kat()
            if w == 'render': render_check('fit_c1.npz'); render_check('fit_n2.npz', 31)
            if w == 'teacher': teacher('fit_c1.npz'); teacher('fit_n2.npz')
            if w == 'init': init_stage('fit_c1.npz'); init_stage('fit_n2.npz')
            if w == 'fit': full_fit('fit_c1.npz')
        except Exception:
            traceback.print_exc()

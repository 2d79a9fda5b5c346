"""Shared helpers of the GPU parity tests and of ``__graft_entry__.smoke()``: drive the CUDA path through the
package / C ABI on the inputs of the golden files and compare with the reference's dumps and the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MODEL_DIR = os.environ.get('MH_TEST_MODEL_DIR', '/tmp/mh_test_model')
COEFS = dict(proj2d=1.0, depth=0.05, silhouette=0.1, reg_velocity=0.05, reg_verts_filter=0.002,
             reg_poses=0.002, reg_scales=1e-4, reg_contact=0.001, reg_foot_sliding=0.01)
NAMES = ['poses_T', 'poses_smpl', 'betas', 'zmin_lin', 'zmax_lin', 'xscale']


def model_dir():
    from oracle import synth
    if not os.path.exists(os.path.join(MODEL_DIR, 'SMPL_NEUTRAL.pkl')):
        synth.write_model_dir(MODEL_DIR, seed=0)
    return MODEL_DIR


def load_fit(name):
    g = np.load(os.path.join(GOLDEN, name))
    N, T, W, H, batch, num_iter, init_iter = [int(x) for x in g['meta_NTWH_batch']]
    data = {k[3:]: g[k] for k in g.files if k.startswith('in_')}
    return g, data, (N, T, W, H, batch, num_iter, init_iter)


class ListLoader(object):
    """Re-iterable yielding the reference's batch dicts (contiguous frames, shuffle=False)."""
    def __init__(self, inputs, batch):
        import torch
        self.inputs, self.batch, self.torch = inputs, batch, torch
        self.T = len(inputs['idxs'])

    def __iter__(self):
        for s in range(0, self.T, self.batch):
            yield {k: self.torch.from_numpy(np.ascontiguousarray(v[s:s + self.batch])) for k, v in self.inputs.items()}


def make_optimizer(pkg, g, data, meta, coefs=None, **kw):
    N, T, W, H, batch, num_iter, init_iter = meta
    c = dict(COEFS if coefs is None else coefs)
    opt = pkg.SMPLDepthSequenceOptimizer(
        image_size=(W, H), num_frames=T, cam_K=g['cam_K'], device='cuda:0', smpl_model_parameters_path=model_dir(),
        proj2d_loss_coef=c['proj2d'], depth_loss_coef=c['depth'], silhouette_loss_coef=c['silhouette'],
        reg_velocity_coef=c['reg_velocity'], reg_verts_filter_coef=c['reg_verts_filter'], reg_poses_coef=c['reg_poses'],
        reg_scales_coef=c['reg_scales'], reg_contact_coef=c['reg_contact'], reg_foot_sliding_coef=c['reg_foot_sliding'], **kw)
    return opt


def prepare(opt, g, data, meta, ingest=True):
    """Context + ingest without running the init loop (init_optimized_variables with num_iter=0)."""
    N, T, W, H, batch, num_iter, init_iter = meta
    opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=0,
                                 batch_size=batch)
    if ingest:
        opt._ingest(ListLoader(data, batch))


def set_filtered(opt, verts_filtered):
    """verts_filtered (T_total, N, V, 3) -> the library's padded (T_local + 2, N, 20672) buffer (halo slots from the
    neighbouring frames when they exist)."""
    import torch
    L = sys.modules[type(opt).__module__.rsplit('.', 1)[0] + '._lib']
    T, N = opt.T_local, opt.num_people
    buf = np.zeros((T + 2, N, L.LD3V), np.float32)
    lo, hi = max(opt.t0 - 1, 0), min(opt.t1 + 1, verts_filtered.shape[0])
    src = verts_filtered[lo:hi].reshape(hi - lo, N, -1)
    s0 = 1 - (opt.t0 - lo)
    buf[s0:s0 + (hi - lo), :, :3 * L.V] = src
    view = opt._view(L.BUF_FILTERED)
    view.copy_(torch.from_numpy(buf.reshape(-1)).to(view.device))
    opt.ctx.call('mh_refresh_filters_flag', 1)


def teacher_forced_cycle(opt, g, data, meta, cycle, halo=None):
    """Parameters / scene / filter state of golden cycle ``cycle`` in -> (log dict, gradient dict) out."""
    import torch
    L = sys.modules[type(opt).__module__.rsplit('.', 1)[0] + '._lib']
    N, T, W, H, batch, num_iter, init_iter = meta
    if opt.ctx is None or not opt._ingested:
        prepare(opt, g, data, meta)
    ctx, st, c = opt.ctx, opt._stream(), cycle
    sl = slice(opt.t0, opt.t1)
    ctx.set_param(L.P_POSES_T, g[f'c{c}_p_poses_T'][sl], st)
    ctx.set_param(L.P_POSES_SMPL, g[f'c{c}_p_poses_smpl'][sl], st)
    ctx.set_param(L.P_BETAS, g[f'c{c}_p_betas'], st)
    ctx.set_param(L.P_BETAS_REF, g['init_betas'], st)
    ctx.set_param(L.P_ZMIN_LIN, g[f'c{c}_p_zmin_lin'][sl], st)
    ctx.set_param(L.P_ZMAX_LIN, g[f'c{c}_p_zmax_lin'][sl], st)
    ctx.set_param(L.P_XSCALE, g[f'c{c}_p_xscale'], st)
    pcd = g[f'c{c}_scene_pcd']
    if len(pcd):
        opt.set_scene_pcd(pcd)
    else:
        ctx.call('mh_set_scene', None, 0, st)
    if c >= 50:
        set_filtered(opt, g['verts_filtered'])
    else:
        ctx.call('mh_clear_filters')
    hp, hn = (0, 0) if halo is None else halo
    ctx.call('mh_fit_grads', hp, hn, st)
    losses = ctx.read_losses(st)
    pkg_shard = sys.modules[type(opt).__module__.rsplit('.', 1)[0] + '.sharding']
    log = pkg_shard.log_from_loss_block(losses, (T + batch - 1) // batch)
    Tl = opt.T_local
    grads = {
        'poses_T': ctx.get_grad(L.P_POSES_T, (Tl, N, 1, 3)), 'poses_smpl': ctx.get_grad(L.P_POSES_SMPL, (Tl, N, 72)),
        'betas': ctx.get_grad(L.P_BETAS, (1, N, 10)), 'zmin_lin': ctx.get_grad(L.P_ZMIN_LIN, (Tl, 1, 1)),
        'zmax_lin': ctx.get_grad(L.P_ZMAX_LIN, (Tl, 1, 1)), 'xscale': ctx.get_grad(L.P_XSCALE, (1, N, 1, 1)),
    }
    return log, grads


def oracle_cycle(g, data, meta, cycle):
    """The CPU oracle (oracle.fit_ref) on the same state: (log dict, gradient dict)."""
    import torch
    from oracle import fit_ref, synth
    N, T, W, H, batch, num_iter, init_iter = meta
    model = synth.load_model_tensors(model_dir())
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
    grads = {nm: p.grad.numpy().copy() for nm, p in zip(NAMES, fr.leaves())}
    return log, grads

"""Multi-GPU validation of the frame-sharded drop-in path over real NCCL:
   torchrun --nproc-per-node N tools/dist_check.py  ->  gpurun_out/dist_check_wN.npz
then `python tools/dist_check.py compare 1 2` checks that the sharded run reproduces the single-GPU run.
The run: init stage (30 Adam iterations), One-Euro refresh (carry hand-over + filtered halo exchange), scene cloud,
3 fit cycles with every term on.  `DIST_CHECK_DEFERRED=1`: the same through the reference's call sequence -- no `batch_size=` at
init (the unmodified Predictor does not pass one): the init stage runs replicated and the frames are sharded when fit() sees the
first batch (`optimizer._reshard`) -> gpurun_out/dist_check_deferred_wN.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
OUT = os.path.join(ROOT, 'gpurun_out')


def run():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    import gpu_harness as gh
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); lr = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    pkg = ge.load_package()
    g, data, meta = gh.load_fit('fit_n2.npz')
    N, T, W, H, batch, num_iter, init_iter = meta
    # 16 frames instead of 4 so that 2, 4 and 8 ranks get whole batches: tile the sequence in time
    rep = 4
    deferred = os.environ.get('DIST_CHECK_DEFERRED', '0') == '1'
    data = {k: np.concatenate([v] * rep, 0) for k, v in data.items()}
    data['idxs'] = np.arange(T * rep, dtype=np.int64)
    T = T * rep
    c = dict(gh.COEFS)
    opt = pkg.SMPLDepthSequenceOptimizer(
        image_size=(W, H), num_frames=T, cam_K=g['cam_K'], device=f'cuda:{lr}', smpl_model_parameters_path=gh.model_dir(), scene_update=False,
        proj2d_loss_coef=c['proj2d'], depth_loss_coef=c['depth'], silhouette_loss_coef=c['silhouette'], reg_velocity_coef=c['reg_velocity'],
        reg_verts_filter_coef=c['reg_verts_filter'], reg_poses_coef=c['reg_poses'], reg_scales_coef=c['reg_scales'],
        reg_contact_coef=c['reg_contact'], reg_foot_sliding_coef=c['reg_foot_sliding'])
    init_log = opt.init_optimized_variables(data['pose2d'], data['poses_smpl'], data['betas_smpl'], data['valid_smpl'], num_iter=30,
                                            batch_size=None if deferred else batch)
    v0 = opt.get_optimized_variables()
    loader = gh.ListLoader(data, batch)
    opt._ingest(loader)
    assert opt._dist == (world > 1)
    opt.set_scene_pcd(g['c31_scene_pcd'])
    opt._refresh_filters(0.01, 0.02, 0.001, 0.5)
    log = opt.fit(loader, num_iter=3)
    v = opt.get_optimized_variables()
    if rank == 0:
        os.makedirs(OUT, exist_ok=True)
        out = {'init_loss': np.array([float(l['loss_2d']) for l in init_log]), 'init_poses_T': v0['poses_T']}
        for k in log[0]:
            out['log_' + k] = np.array([l[k] for l in log])
        for k in ('poses_T', 'poses_smpl', 'betas_smpl', 'scale_factor', 'min_z', 'max_z'):
            out['final_' + k] = v[k]
        np.savez(os.path.join(OUT, f'dist_check_{"deferred_" if deferred else ""}w{world}.npz'), **out)
        print(f'world {world}: wrote results; ranges {opt.ranges}; last log', {k: float(log[-1][k]) for k in log[-1]})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def compare(a, b, tag=''):
    A = np.load(os.path.join(OUT, f'dist_check_w{a}.npz')); B = np.load(os.path.join(OUT, f'dist_check_{tag}w{b}.npz'))
    print(f'--- world {a} vs {tag}world {b}')
    ok = True
    for k in A.files:
        d = np.abs(A[k] - B[k]).max(); s = np.abs(A[k]).max()
        tol = 2e-4 * s + 1e-7
        flag = 'ok' if d <= tol else 'MISMATCH'
        ok &= d <= tol
        print(f'{k:22s} max|a| {s:.4e} max diff {d:.3e} {flag}')
    print('DIST CHECK', 'PASSED' if ok else 'FAILED')
    return ok


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'compare':
        sys.exit(0 if compare(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else '') else 1)
    run()

"""Development: SMPL forward with the SIMT / tcgen05 pose-corrective contraction (MH_GEMM_TC=0/1) on the same random bodies.
   python tools/tc_check.py run OUT.npz      (once per setting)      python tools/tc_check.py cmp A.npz B.npz"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

if sys.argv[1] == 'run':
    import torch
    import __graft_entry__ as ge
    import gpu_harness as gh
    pkg = ge.load_package()
    g, data, meta = gh.load_fit('fit_c1.npz')
    opt = gh.make_optimizer(pkg, g, data, meta)
    gh.prepare(opt, g, data, meta, ingest=False)
    kat = np.load(os.path.join(gh.GOLDEN, 'kat_functions.npz'))
    v, j = opt.smpl_forward(kat['smpl_betas'], kat['smpl_poses'])
    print('KAT verts max abs diff', np.abs(v - kat['smpl_verts']).max(), flush=True)
    rng = np.random.default_rng(5)
    nb = 300                                            # not a multiple of 128: exercises the row tail
    betas = rng.normal(0, 0.5, (nb, 10)).astype(np.float32)
    poses = rng.normal(0, 0.3, (nb, 72)).astype(np.float32)
    v2, j2 = opt.smpl_forward(betas, poses)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        opt.smpl_forward(betas, poses, want_verts=False)
    torch.cuda.synchronize()
    print('5 forwards of', nb, 'bodies:', round((time.time() - t0) * 1e3, 2), 'ms', flush=True)
    np.savez(sys.argv[2], v=v, v2=v2, j2=j2)
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in ('v', 'v2', 'j2'):
        print(k, 'max abs diff', np.abs(a[k] - b[k]).max(), 'max abs', np.abs(a[k]).max(), 'nan', int(np.isnan(b[k]).sum()))

"""Where does the construction of a C3-size context spend its time?  (mh_create = device allocations, mh_set_model = model upload +
layouts.)  Several contexts in a row, the first one kept alive as in bench.py."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import __graft_entry__ as ge

pkg = ge.load_package()
L = sys.modules[pkg.__name__ + '._lib']
smpl_io = sys.modules[pkg.__name__ + '.smpl_io']
torch.cuda.set_device(0)
torch.zeros(1, device='cuda:0')
model = smpl_io.load_smpl_model(bench.model_dir()) if hasattr(smpl_io, 'load_smpl_model') else None
keep = []
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx = L.Context(512, 8, 720, 1280, B=8, device=0, rank=0, world=1, t0=0, T_total=512, M_max=200000)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if model is not None:
        ctx.set_model(model)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'rep {rep}: mh_create {t1 - t0:.3f} s, set_model {t2 - t1:.3f} s, free GPU memory {torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB', flush=True)
    if rep == 0:
        keep.append(ctx)
    else:
        t3 = time.perf_counter()
        ctx.close()
        torch.cuda.synchronize()
        print(f'        close {time.perf_counter() - t3:.3f} s', flush=True)

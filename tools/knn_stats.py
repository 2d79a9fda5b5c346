"""Contact-term grid statistics on a bench workload: person-frames answered by the grid, cells, cell size, and the time of a cycle's
loss-term stage with and without the grid.  usage: python tools/knn_stats.py [workload]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import __graft_entry__ as ge

pkg = ge.load_package()
L = sys.modules[pkg.__name__ + '._lib']
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c3']
device = torch.device('cuda', 0)
torch.cuda.set_device(device)
opt, aux = bench.build_problem(pkg, w, device)
opt._refresh_filters(0.01, 0.02, 0.001, 0.5)
st = opt._stream()
for grid in ('1', '0'):
    os.environ['MH_KNN_GRID'] = grid
    opt.set_scene_pcd(aux['cloud'])
    opt.ctx.call('mh_set_timing', 1)
    for _ in range(4):
        opt.step_device_only(0.001)
    torch.cuda.synchronize()
    out = np.zeros(4, np.int64)
    opt.ctx.call('mh_debug_knn_stats', L.ptr(out), st)
    tm = opt.ctx.read_timing(4)
    opt.ctx.call('mh_set_timing', 0)
    print(f'grid={grid} resolved={out[0]} of {w["N"] * w["T"]} cells={out[1]} points={out[2]} h={out[3] / 1e6:.4f} m  stage_ms={tm[-1].round(3).tolist()}')

"""One teacher-forced fit cycle (all terms) + filter refresh + scene median on the small golden case, for compute-sanitizer."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import __graft_entry__ as ge
import gpu_harness as gh
pkg = ge.load_package()
L = sys.modules[pkg.__name__ + '._lib']
g, data, meta = gh.load_fit('fit_n2.npz')
opt = gh.make_optimizer(pkg, g, data, meta)
log, grads = gh.teacher_forced_cycle(opt, g, data, meta, 51)
opt._refresh_filters(0.01, 0.02, 0.001, 0.5)
N, T, W, H = meta[:4]
back = (np.random.default_rng(0).random((T, H, W)) > 0.3).astype(np.uint8)
opt.ctx.call('mh_scene_set_back', 0, T, L.ptr(back), None, opt._stream())
d, m = opt._device_median(0)
opt.update_scene_pointcloud(np.where(m, d, 5.0).astype(np.float32), np.ones((H, W), bool))
opt.ctx.call('mh_fit_grads', 0, 0, opt._stream())
opt.ctx.call('mh_fit_update', 0.01, opt._stream())
# round 2: the device scene update (median -> post-processing -> cloud), the fused cycle call and the u8 ingest entry
opt._median_passes(0)
opt.ctx.call('mh_scene_update_from_median', 1, 7, None, opt._stream())
opt.ctx.call('mh_fit_cycle', 0.01, opt._stream())
d8 = dict(data); d8['seg_mask'] = data['seg_mask'].astype(np.uint8)
opt._ingest(gh.ListLoader(d8, meta[4]))
opt.ctx.call('mh_fit_cycle', 0.01, opt._stream())
print('ok', {k: round(v, 5) for k, v in log.items()}, opt.ctx.read_losses(opt._stream())[:3])

"""Development: which (row, column, k) mapping does the tcgen05 contraction realise?  A = identity-like patterns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import __graft_entry__ as ge
import gpu_harness as gh

pkg = ge.load_package()
L = sys.modules[pkg.__name__ + '._lib']
g, data, meta = gh.load_fit('fit_c1.npz')
opt = gh.make_optimizer(pkg, g, data, meta)
gh.prepare(opt, g, data, meta, ingest=False)
ctx = opt.ctx
M = 300
A = np.zeros((M, 192), np.float32)
A[np.arange(M), np.arange(M) % 192] = 1.0
C0 = np.zeros((M, L.LD3V), np.float32); C1 = np.zeros_like(C0)
ctx.call('mh_debug_gemm_fwd', L.ptr(A), L.ptr(C0), M, 0)
ctx.call('mh_debug_gemm_fwd', L.ptr(A), L.ptr(C1), M, 1)
B = C0[:192].copy()                                      # rows of the basis
print('identity pattern: max abs diff', np.abs(C1 - C0).max(), 'max abs', np.abs(C0).max(), 'nan', int(np.isnan(C1).sum()))
Bn = B / (np.linalg.norm(B, axis=1, keepdims=True) + 1e-30)
for m in (0, 1, 2, 3, 4, 7, 8, 9, 31, 32, 127, 128, 200, 299):
    r = C1[m]
    if not np.isfinite(r).all():
        print('row', m, 'non-finite'); continue
    corr = Bn @ (r / (np.linalg.norm(r) + 1e-30))
    k = int(np.argmax(np.abs(corr)))
    print(f'row {m:3d}: |row| {np.linalg.norm(r):.4e} (expected {np.linalg.norm(C0[m]):.4e}) best basis row {k} corr {corr[k]:+.4f} expected {m % 192}')
# column mapping inside the first tile, row 0: which source column does output column n carry?
src = B[0, :256]
for n in list(range(0, 12)) + [32, 33, 64, 128, 255]:
    d = np.abs(src - C1[0, n])
    print('col', n, '<- source col', int(np.argmin(d)), 'err', float(d.min()), 'value', float(C1[0, n]), 'expected', float(src[n]))
rng = np.random.default_rng(0)
A = rng.normal(0, 0.2, (M, 192)).astype(np.float32)
ctx.call('mh_debug_gemm_fwd', L.ptr(A), L.ptr(C0), M, 0)
ctx.call('mh_debug_gemm_fwd', L.ptr(A), L.ptr(C1), M, 1)
ref = A.astype(np.float64) @ B.astype(np.float64)
print('random A: tc vs simt', np.abs(C1 - C0).max(), ' simt vs f64', np.abs(C0 - ref).max(), ' tc vs f64', np.abs(C1 - ref).max(), ' max |C|', np.abs(ref).max())

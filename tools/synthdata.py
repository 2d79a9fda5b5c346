"""Synthetic SMPL-shaped body model and synthetic multi-person sequences (pure numpy).

TEST / BENCH INPUT GENERATION ONLY -- neither product code nor oracle: ``bench.py``, ``tests/`` and ``oracle/synth.py``
import it.  The licensed
``SMPL_NEUTRAL.pkl`` and the MuPoTS data are not available offline, so every
test, golden vector and benchmark runs on this seeded stand-in (SURVEY.md
section 8d): a UV-ellipsoid with exactly V=6890 vertices / F=13776 faces, the
SMPL parent table, sparse skinning weights and sparse joint regressors with the
same shapes/dtypes the reference's ``SMPL`` class loads (``smpl.py:179-275``).
"""
import os
import pickle

import numpy as np

V, F_, NJ = 6890, 13776, 24
PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def _ellipsoid(rings=82, segs=84, axes=(0.25, 0.85, 0.15)):
    th = np.pi * (np.arange(1, rings + 1) / (rings + 1))            # polar angle from +Y pole
    ph = 2 * np.pi * np.arange(segs) / segs
    ring = np.stack([np.outer(np.sin(th), np.cos(ph)),
                     np.outer(np.cos(th), np.ones_like(ph)),
                     np.outer(np.sin(th), np.sin(ph))], axis=-1).reshape(-1, 3)
    verts = np.concatenate([[[0, 1, 0]], ring, [[0, -1, 0]]], axis=0) * np.array(axes)
    idx = lambda r, s: 1 + r * segs + (s % segs)
    faces = []
    for s in range(segs):                                           # top fan
        faces.append([0, idx(0, s + 1), idx(0, s)])
    for r in range(rings - 1):
        for s in range(segs):
            a, b, c, d = idx(r, s), idx(r, s + 1), idx(r + 1, s), idx(r + 1, s + 1)
            faces.append([a, b, d])
            faces.append([a, d, c])
    last = 1 + rings * segs
    for s in range(segs):                                           # bottom fan
        faces.append([last, idx(rings - 1, s), idx(rings - 1, s + 1)])
    return verts.astype(np.float64), np.array(faces, dtype=np.int64)


def _sparse_regressor(rng, verts, centres, k):
    """(J,V) rows: random convex weights on the k nearest vertices of each centre."""
    reg = np.zeros((len(centres), len(verts)), np.float64)
    for j, c in enumerate(centres):
        d = np.sum((verts - c) ** 2, axis=1)
        nn = np.argsort(d)[:k]
        w = rng.random(k) + 0.05
        reg[j, nn] = w / w.sum()
    return reg


def make_smpl_model(seed=0):
    """Returns the dict the reference's ``SMPL`` loader expects in ``SMPL_NEUTRAL.pkl``
    (keys ``v_template, f, shapedirs, posedirs, J_regressor, kintree_table, weights``)
    plus the four extra regressors in their on-disk shapes/dtypes."""
    rng = np.random.default_rng(seed)
    verts, faces = _ellipsoid()
    assert verts.shape == (V, 3) and faces.shape == (F_, 3)
    # joint centres: a plausible tree inside the body (root at origin, Y up in model space)
    centres = np.zeros((NJ, 3))
    layout = {0: (0, 0, 0), 1: (0.08, -0.08, 0), 2: (-0.08, -0.08, 0), 3: (0, 0.12, 0),
              4: (0.09, -0.38, 0), 5: (-0.09, -0.38, 0), 6: (0, 0.25, 0), 7: (0.08, -0.68, 0),
              8: (-0.08, -0.68, 0), 9: (0, 0.36, 0), 10: (0.09, -0.78, 0.05), 11: (-0.09, -0.78, 0.05),
              12: (0, 0.52, 0), 13: (0.07, 0.45, 0), 14: (-0.07, 0.45, 0), 15: (0, 0.62, 0),
              16: (0.15, 0.45, 0), 17: (-0.15, 0.45, 0), 18: (0.19, 0.22, 0), 19: (-0.19, 0.22, 0),
              20: (0.2, 0.0, 0), 21: (-0.2, 0.0, 0), 22: (0.2, -0.08, 0), 23: (-0.2, -0.08, 0)}
    for j, c in layout.items():
        centres[j] = c
    centres[1:] += rng.normal(0, 0.01, (NJ - 1, 3))
    J_regressor = _sparse_regressor(rng, verts, centres, 32)
    # remove the regression bias so that joint 0 sits at the origin for beta = 0
    J0 = J_regressor @ verts
    verts = verts - J0[0:1]
    J = J_regressor @ verts
    d2 = np.sum((verts[:, None, :] - J[None, :, :]) ** 2, axis=-1)          # (V,24)
    w = np.exp(-d2 / 0.02)
    order = np.argsort(-w, axis=1)
    keep = np.zeros_like(w, dtype=bool)
    np.put_along_axis(keep, order[:, :6], True, axis=1)                   # at most 6 bones / vertex
    keep &= w >= 0.08 * w.max(axis=1, keepdims=True)
    w = np.where(keep, w, 0.0)
    w /= w.sum(axis=1, keepdims=True)
    shapedirs = rng.normal(0, 0.005, (V, 3, 10))
    posedirs = rng.normal(0, 0.002, (V, 3, 207))
    kintree = np.stack([np.array([2 ** 32 - 1] + PARENTS[1:], dtype=np.int64), np.arange(NJ)], axis=0)
    model = {
        'v_template': verts, 'f': faces.astype(np.uint32), 'shapedirs': shapedirs, 'posedirs': posedirs,
        'J_regressor': J_regressor, 'kintree_table': kintree, 'weights': w,
    }
    # extra regressors (on-disk conventions of model_data/parameters/*.npy)
    c17 = verts[rng.choice(V, 17, replace=False)] * 0.8
    c17b = verts[rng.choice(V, 17, replace=False)] * 0.8
    extras = {
        'SMPL_AlphaPose_Regressor_RMSprop_6.npy': _sparse_regressor(rng, verts, c17, 40).T.astype(np.float32),
        'SMPL_MuPoTs_Regressor_v1.npy': _sparse_regressor(rng, verts, c17b, 64).T.astype(np.float32),
        'J_regressor_extra.npy': _sparse_regressor(rng, verts, verts[rng.choice(V, 9, replace=False)], 7),
        'J_regressor_h36m.npy': _sparse_regressor(rng, verts, verts[rng.choice(V, 17, replace=False)] * 0.8, 6),
    }
    return model, extras


def write_model_dir(path, seed=0, real_regressor_dir=None):
    """Writes ``SMPL_NEUTRAL.pkl`` + the four regressor ``.npy`` files into ``path``
    in the layout ``SMPLOptimizerBase`` reads (``optimizer.py:36-39, 65-72``).  If
    ``real_regressor_dir`` holds the reference's shipped regressors they are used
    instead of the synthetic ones."""
    os.makedirs(path, exist_ok=True)
    model, extras = make_smpl_model(seed)
    with open(os.path.join(path, 'SMPL_NEUTRAL.pkl'), 'wb') as f:
        pickle.dump(model, f)
    for name, arr in extras.items():
        src = os.path.join(real_regressor_dir, name) if real_regressor_dir else None
        if src and os.path.exists(src):
            arr = np.load(src)
        np.save(os.path.join(path, name), arr)
    return path


def load_model_tensors(path):
    """Model dir -> dict of float32 numpy arrays in the layouts of ``SMPL.__init__``
    (``smpl.py:201-275``): posedirs (207,3V); parents with -1 root; alphapose /
    mupots regressors transposed to (17,V)."""
    with open(os.path.join(path, 'SMPL_NEUTRAL.pkl'), 'rb') as f:
        d = pickle.load(f, encoding='latin1')
    def dense(a):
        return np.array(a.todense() if hasattr(a, 'todense') else a, dtype=np.float32)
    parents = np.array(d['kintree_table'][0]).astype(np.int64)
    parents[0] = -1
    out = {
        'v_template': dense(d['v_template']),
        'faces': np.array(d['f']).astype(np.int32),
        'shapedirs': dense(d['shapedirs'])[:, :, :10],
        'posedirs': dense(d['posedirs']).reshape(-1, np.array(d['posedirs']).shape[-1]).T.copy(),
        'J_regressor': dense(d['J_regressor']),
        'lbs_weights': dense(d['weights']),
        'parents': parents.astype(np.int32),
    }
    ap = os.path.join(path, 'SMPL_AlphaPose_Regressor_RMSprop_6.npy')
    out['J_regressor_alphapose'] = np.load(ap).T.astype(np.float32).copy()
    mp = os.path.join(path, 'SMPL_MuPoTs_Regressor_v1.npy')
    if os.path.exists(mp):
        out['J_regressor_mupots'] = np.load(mp).T.astype(np.float32).copy()
    return out


# ----------------------------------------------------------------------------
# sequences
# ----------------------------------------------------------------------------
def make_motion(N, T, seed=1):
    """Ground-truth SMPL parameters of a smooth N-person, T-frame sequence."""
    rng = np.random.default_rng(seed)
    th0 = rng.normal(0, 0.15, (1, N, 72))
    th0[..., 0] += np.pi                                   # global orient: Y down in camera space
    theta = th0 + 0.02 * np.cumsum(rng.normal(0, 1, (T, N, 72)), axis=0)
    theta[..., 66:] = 0.0
    beta_p = rng.normal(0, 0.5, (1, N, 10))
    z = 3.0 + np.arange(N) * 1.0 + rng.random(N) * 0.5 if N > 1 else np.array([4.0])
    xspread = np.linspace(-0.35, 0.35, N) if N > 1 else np.array([0.0])
    rng.shuffle(xspread)
    trans0 = np.stack([xspread * z * 0.9, np.full(N, 0.1), z], axis=-1)     # (N,3)
    drift = 0.02 * np.cumsum(rng.normal(0, 0.5, (T, N, 3)), axis=0) * np.array([1.0, 0.1, 1.0])
    trans = trans0[None] + drift
    return {'theta': theta.astype(np.float32), 'beta': beta_p.astype(np.float32),
            'trans': trans.astype(np.float32)}


def camera_for(W, H, fov=60.0):
    f = 0.5 * min(W, H) / np.tan(np.pi * fov / 360.0) * 1.35          # person spans ~0.5 H at z ~ 4 m
    return np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], np.float32)


def scene_cloud(M, seed=2, y_ground=1.0, z_wall=10.0):
    """M points on the ground plane y=+1 m (Y down) and the back wall, + 5 mm noise."""
    rng = np.random.default_rng(seed)
    mg = int(M * 0.8)
    g = np.stack([rng.uniform(-6, 6, mg), np.full(mg, y_ground), rng.uniform(1.5, z_wall, mg)], -1)
    w = np.stack([rng.uniform(-6, 6, M - mg), rng.uniform(-3, y_ground, M - mg), np.full(M - mg, z_wall)], -1)
    return (np.concatenate([g, w], 0) + rng.normal(0, 0.005, (M, 3))).astype(np.float32)


def scene_depth_plane(W, H, cam_K, y_ground=1.0, z_wall=10.0):
    """Per-pixel depth of (ground plane y=y_ground) U (wall z=z_wall) through the pixel centres."""
    u = (np.arange(W) + 0.5 - cam_K[0, 2]) / cam_K[0, 0]
    v = (np.arange(H) + 0.5 - cam_K[1, 2]) / cam_K[1, 1]
    vv = np.tile(v[:, None], (1, W))
    with np.errstate(divide='ignore'):
        zg = np.where(vv > 1e-6, y_ground / np.maximum(vv, 1e-6), np.inf)
    return np.minimum(zg, z_wall).astype(np.float32)


def assemble_inputs(zbufs, motion, joints2d, cam_K, W, H, seed=1):
    """From per-person z-buffers (T,N,H,W; <=0 empty) build the modalities the
    dataset hands to the optimiser (``datautils.py:531-542``): ``depths`` = min-max
    normalised disparity of scene U persons; ``seg_mask`` = nearest-person-wins
    instance masks (disjoint, float {0,1}, ``utils.py:314-333`` semantics);
    ``backmasks`` = no person; noisy ``pose2d`` [x, y, conf]; perturbed ROMP-like
    ``poses_smpl`` / ``betas_smpl``; ``valid_smpl`` = 1."""
    rng = np.random.default_rng(seed + 100)
    T, N = zbufs.shape[:2]
    zb = np.where(zbufs > 0, zbufs, np.inf)
    nearest = np.argmin(zb, axis=1)                                  # (T,H,W)
    zmin = np.min(zb, axis=1)
    covered = np.isfinite(zmin)
    seg = np.zeros((T, N, H, W), np.float32)
    for n in range(N):
        seg[:, n] = (covered & (nearest == n)).astype(np.float32)
    scene = scene_depth_plane(W, H, cam_K)[None]
    depth = np.where(covered, np.minimum(zmin, scene), scene)
    disp = 1.0 / depth
    dmin = disp.reshape(T, -1).min(1)[:, None, None]
    dmax = disp.reshape(T, -1).max(1)[:, None, None]
    depths = ((disp - dmin) / np.maximum(dmax - dmin, 1e-9)).astype(np.float32)
    backmasks = (~covered).astype(np.uint8)
    conf = np.where(rng.random((T, N, 17, 1)) < 0.05, 0.1, 0.9)
    pose2d = np.concatenate([joints2d + rng.normal(0, 0.3, joints2d.shape), conf], -1).astype(np.float32)
    poses_smpl = (motion['theta'] + rng.normal(0, 0.05, motion['theta'].shape)).astype(np.float32)
    poses_smpl[..., 66:] = 0.0
    betas_smpl = (np.tile(motion['beta'], (T, 1, 1)) + rng.normal(0, 0.05, (T, N, 10))).astype(np.float32)
    images = rng.integers(0, 255, (T, H, W, 3), dtype=np.uint8)
    return {
        'images': images, 'depths': depths, 'seg_mask': seg, 'backmasks': backmasks,
        'pose2d': pose2d, 'poses_smpl': poses_smpl, 'betas_smpl': betas_smpl,
        'valid_smpl': np.ones((T, N, 1), np.float32),
        'idxs': np.arange(T, dtype=np.int64),
    }

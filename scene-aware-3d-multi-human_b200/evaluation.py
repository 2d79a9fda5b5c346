"""Evaluation of optimised sequences against 3-D ground truth.

Mirror of the reference's ``mhmocap/evaluate.py:180-296`` (``compute_smpl_pred_error_3dproj``), ``:401-430``
(``masked_average_error`` / ``masked_average_pck``) and ``mhmocap/eval_mupots.py:18-42`` (``compute_mm_pck_results``): same
argument meaning, same returned keys and shapes.  The SMPL evaluation -- the only heavy part: T x N bodies, 6890 vertices each --
runs on the GPU through ``mh_smpl_regress`` (``libmhopt.so``); the per-frame Hungarian matching of annotated to predicted
persons and the masked averages are a few kFLOP and stay on the host, as in the reference.

Quirks of the reference that are reproduced (they change the numbers):
* rows of the per-frame outputs are indexed by the MATCHED PAIR (in order of the annotated person), not by the annotated person
  (``evaluate.py:263``): with fewer predictions than annotations the trailing rows stay zero / invalid;
* joint validity is ``vis > 0.49`` (``:283``), root validity ``vis[14] > 0`` (``:264``);
* the jitter of frame 0 is a copy of frame 1's (``:289``);
* with ``Kd`` the y distortion term is ``2 Kd[3] y^2`` (``transforms.py:43-47``), not the Brown-Conrady ``2 p2 x y``.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment

# CMU-Panoptic (19 joints) / AlphaPose (17 joints) -> the 15 MuPoTS joints used by the metrics: (weights, source joints) per
# output joint (evaluate.py:31-64)
_PANOPTIC_TO_MUPOTS15 = [((1.0,), (1,)), ((1.0,), (0,)), ((1.0,), (9,)), ((1.0,), (10,)), ((1.0,), (11,)), ((1.0,), (3,)), ((1.0,), (4,)),
                         ((1.0,), (5,)), ((1.0,), (12,)), ((1.0,), (13,)), ((1.0,), (14,)), ((1.0,), (6,)), ((1.0,), (7,)), ((1.0,), (8,)),
                         ((1.0,), (2,))]
_ALPHAPOSE_TO_MUPOTS15 = [((1.0,), (0,)), ((0.5, 0.5), (5, 6)), ((1.0,), (6,)), ((1.0,), (8,)), ((1.0,), (10,)), ((1.0,), (5,)), ((1.0,), (7,)),
                          ((1.0,), (9,)), ((1.0,), (12,)), ((1.0,), (14,)), ((1.0,), (16,)), ((1.0,), (11,)), ((1.0,), (13,)), ((1.0,), (15,)),
                          ((0.5, 0.5), (11, 12))]


def _remap(x, table):
    """(B, J_in, D) -> (B, len(table), D): weighted sums of source joints, accumulated in float32 like the reference (evaluate.py:66-90)."""
    x = np.asarray(x)
    y = np.zeros((x.shape[0], len(table), x.shape[2]), np.float32)
    for j, (w, src) in enumerate(table):
        wv = np.asarray(w, np.float32)[None, :, None]
        y[:, j] = (wv * x[:, np.asarray(src, int)]).sum(axis=1)
    return y


def project_points(pts3d, K, Kd=None):
    """Pinhole projection with the reference's optional 5-coefficient distortion (transforms.py:19-54); (P, 3) -> (P, 2)."""
    p = np.asarray(pts3d)
    uv = p[:, :2] / p[:, 2:3]
    if Kd is not None:
        x, y = uv[:, 0].copy(), uv[:, 1].copy()
        r = x * x + y * y
        radial = 1 + Kd[0] * r + Kd[1] * r * r + Kd[4] * r * r * r
        xd = x * radial + 2 * Kd[2] * x * y + Kd[3] * (r + 2 * x * x)
        yd = y * radial + 2 * Kd[3] * y * y + Kd[2] * (r + 2 * y * y)            # as the code, not the textbook (transforms.py:43-47)
        uv = np.stack([xd, yd], -1)
    return (uv[:, :2] @ K[:2, :2].T) + K[0:2, 2:3].T


def match_persons(ref2d, pred2d, thr=0.5):
    """Hungarian matching of annotated to predicted 2-D poses on the mean distance of the jointly visible joints
    (utils.py:278-311).  ref2d (K, J, 3) and pred2d (N, J, 3) carry [u, v, visibility]; the distance is taken over all three
    channels, as in the reference."""
    K, N = ref2d.shape[0], pred2d.shape[0]
    a = np.broadcast_to(ref2d[:, None], (K, N) + ref2d.shape[1:])
    b = np.broadcast_to(pred2d[None], (K, N) + pred2d.shape[1:])
    both = (a[..., 2] > thr) & (b[..., 2] > thr)
    dist = np.sqrt(np.sum(np.square(a - b), axis=-1))
    cost = np.full((K, N), 1e6, np.float32)
    for k in range(K):
        for n in range(N):
            if both[k, n].any():
                cost[k, n] = np.mean(dist[k, n][both[k, n]])
    return linear_sum_assignment(cost)


def compute_smpl_pred_error_3dproj(output_data, ref_poses3d, visibility, smpl_joints, cam_K, Kd=None):
    """Distances between the optimised SMPL bodies and the annotated 3-D poses (evaluate.py:180-296).

    ``output_data``: the dict ``get_optimized_variables()`` returns (``poses_T`` (T,N,1,3), ``poses_smpl`` (T,N,72), ``betas_smpl``
    (T,N,10) -- repeat the per-person shape over T as eval_mupots.py:122-123 does --, ``scale_factor`` (1|T,N,1,1)).
    ``ref_poses3d`` (T,K,17|19,3), ``visibility`` (T,K,J,1).  ``smpl_joints(betas (B,10), poses (B,72), which) -> (B,17,3)`` evaluates
    SMPL and regresses the ``'mupots'`` or ``'alphapose'`` joints (un-scaled, un-translated): ``SMPLJoints`` below runs it on the GPU.
    """
    pT = np.asarray(output_data['poses_T'])
    sc = np.asarray(output_data['scale_factor'])
    th = np.asarray(output_data['poses_smpl'])
    be = np.asarray(output_data['betas_smpl'])
    T, N = pT.shape[:2]
    if sc.shape[0] == 1:
        sc = np.tile(sc, (T, 1, 1, 1))
    K, J = ref_poses3d.shape[1:3]
    if J not in (17, 19):
        raise ValueError(f'only 17 (MuPoTS) or 19 (CMU Panoptic) joints are supported, {J} given')
    if J == 19:
        ref = _remap(np.reshape(ref_poses3d, (T * K, -1, 3)), _PANOPTIC_TO_MUPOTS15).reshape(T, K, -1, 3)
        vis = _remap(np.reshape(visibility, (T * K, -1, 1)), _PANOPTIC_TO_MUPOTS15).reshape(T, K, -1, 1)
        pred_local = _remap(smpl_joints(be.reshape(-1, 10), th.reshape(-1, 72), 'alphapose'), _ALPHAPOSE_TO_MUPOTS15).reshape(T, N, -1, 3)
    else:
        ref, vis = np.asarray(ref_poses3d)[:, :, :15], np.asarray(visibility)[:, :, :15]
        pred_local = np.asarray(smpl_joints(be.reshape(-1, 10), th.reshape(-1, 72), 'mupots')).reshape(T, N, 17, 3)[:, :, :15]
    ref2d = np.concatenate([project_points(ref.reshape(-1, 3), cam_K, Kd).reshape(T, K, -1, 2), vis], axis=-1)

    out = {k: np.zeros(s, np.float32) for k, s in (('abs_dist', (T, K, 14)), ('rel_dist', (T, K, 14)), ('valid_joints', (T, K, 14)),
                                                    ('abs_root_pos_err', (T, K)), ('valid_root', (T, K)))}
    m_ref, m_pred = np.zeros((T, K, 14, 3), np.float32), np.zeros((T, K, 14, 3), np.float32)
    for t in range(T):
        pred3d = sc[t] * pred_local[t] + pT[t]                                    # (N, 15, 3), evaluate.py:249
        p2 = project_points(pred3d.reshape(-1, 3), cam_K, Kd).reshape(N, -1, 2)
        pred2d = np.concatenate([p2, np.ones_like(p2[..., :1])], axis=-1)
        ri, pi = match_persons(ref2d[t], pred2d)
        for row, (g, p, v) in enumerate(zip(ref[t][ri], pred3d[pi], vis[t][ri])):  # row = matched pair, NOT the annotated person
            if v[14, 0] > 0:
                out['valid_root'][t, row] = 1
                out['abs_root_pos_err'][t, row] = np.sqrt(np.sum(np.square(g[14] - p[14])))
            m_ref[t, row], m_pred[t, row] = g[:14], p[:14]
            out['abs_dist'][t, row] = np.sqrt(np.sum(np.square(g[:14] - p[:14]), axis=-1))
            out['rel_dist'][t, row] = np.sqrt(np.sum(np.square((g[:14] - g[14:15]) - (p[:14] - p[14:15])), axis=-1))
            out['valid_joints'][t, row] = (v[:14, 0] > 0.49).astype(np.float32)
    step = lambda a: np.sqrt(np.sum(np.square(a[1:] - a[:-1]), axis=-1))         # noqa: E731
    jit = np.abs(step(m_ref) - step(m_pred))
    out['abs_jitter'] = np.concatenate([jit[0:1], jit], axis=0)
    return out


def masked_average_error(dist, vis):
    """Mean of ``dist`` over the entries with ``vis > 0.5`` (evaluate.py:401-416)."""
    if np.shape(dist) != np.shape(vis):
        raise ValueError(f'shape mismatch {np.shape(dist)} / {np.shape(vis)}')
    d = np.asarray(dist, np.float32).reshape(-1)
    v = (np.asarray(vis).reshape(-1) > 0.5).astype(np.float32)
    return np.sum(v * d) / np.clip(np.sum(v), 1, None)


def masked_average_pck(dist, vis, thr):
    """Fraction of the entries with ``vis > 0.5`` whose ``dist <= thr`` (evaluate.py:419-434)."""
    if np.shape(dist) != np.shape(vis):
        raise ValueError(f'shape mismatch {np.shape(dist)} / {np.shape(vis)}')
    d = np.asarray(dist, np.float32).reshape(-1)
    v = (np.asarray(vis).reshape(-1) > 0.5).astype(np.float32)
    return np.sum(v * (d <= thr)) / np.clip(np.sum(v), 1, None)


def compute_mm_pck_results(optvar, ref_poses3d, visibility, smpl_joints, cam_K, Kd=None):
    """The six sequence metrics of ``eval_mupots.py:18-42``: MPJPE absolute / root-relative and root position error in mm, PCK@150 mm
    (root-relative) and AP@250 mm (root) in percent, jitter in mm."""
    m = compute_smpl_pred_error_3dproj(optvar, ref_poses3d, visibility, smpl_joints, cam_K, Kd)
    return {
        'mm_abs_error': 1000 * masked_average_error(m['abs_dist'], m['valid_joints']),
        'mm_rel_error': 1000 * masked_average_error(m['rel_dist'], m['valid_joints']),
        'mm_mrpe': 1000 * masked_average_error(m['abs_root_pos_err'], m['valid_root']),
        'pck_rel': 100 * masked_average_pck(m['rel_dist'], m['valid_joints'], 0.15),
        'ap25_root': 100 * masked_average_pck(m['abs_root_pos_err'], m['valid_root'], 0.25),
        'abs_jitter': 1000 * masked_average_error(m['abs_jitter'], m['valid_joints']),
    }


class SMPLJoints:
    """``smpl_joints`` callable on the GPU: SMPL forward + sparse joint regression through ``mh_smpl_regress``.

    ``optimizer`` is a ``SMPLDepthSequenceOptimizer`` whose context holds the model; ``regressors`` maps ``'mupots'`` /
    ``'alphapose'`` to the dense regressor as stored in ``model_data/parameters/`` (``SMPL_MuPoTs_Regressor_v1.npy``,
    ``SMPL_AlphaPose_Regressor_RMSprop_6.npy``: (6890, J); (J, 6890) is accepted too)."""

    def __init__(self, optimizer, regressors):
        self.opt = optimizer
        self.reg = {k: np.asarray(v, np.float32) for k, v in regressors.items()}

    def __call__(self, betas, poses, which):
        return self.opt.smpl_regress(betas, poses, self.reg[which])

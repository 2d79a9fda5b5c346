"""Torch-CPU fp32 restatement of the reference-owned arithmetic on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference file:line (relative to ``/root/reference``) whose behaviour it follows;
it is written from the algorithm, not copied, and keeps the reference's quirks
because parity is defined against the reference's code, not the textbook.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14,
                16, 17, 18, 19, 20, 21]


# ----------------------------------------------------------------------------
# SMPL  (mhmocap/smpl.py)
# ----------------------------------------------------------------------------
def rodrigues(rvec):
    """Axis-angle (n,3) -> rotation matrices (n,3,3).

    Follows ``smpl.py:647-678``: the angle is ``||r + 1e-8||`` (1e-8 added to each
    component BEFORE the norm, ``:662``), the axis is ``r / angle`` (not exactly
    unit), and ``R = I + sin(a) K + (1 - cos(a)) K K``.
    """
    angle = torch.sqrt(torch.sum((rvec + 1e-8) ** 2, dim=1, keepdim=True))
    k = rvec / angle
    kx, ky, kz = k[:, 0], k[:, 1], k[:, 2]
    zero = torch.zeros_like(kx)
    K = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], dim=1).view(-1, 3, 3)
    s = torch.sin(angle).unsqueeze(-1)
    c = torch.cos(angle).unsqueeze(-1)
    eye = torch.eye(3, dtype=rvec.dtype).unsqueeze(0)
    return eye + s * K + (1.0 - c) * torch.bmm(K, K)


def smpl_forward(model, betas, poses):
    """betas (nb,10), poses (nb,72) -> dict(verts (nb,V,3), joints24 (nb,24,3), A, v_posed).

    Follows ``lbs`` (``smpl.py:490-576``) and ``batch_rigid_transform``
    (``smpl.py:692-746``): shape blend (``:532``), rest joints (``:535``),
    Rodrigues on the first 22 joints only with joints 22/23 forced to identity
    (``:544-546``), pose correctives (``:547-558``), the 23-step parent chain
    (``:725-731``), rest-pose removal (``:741-744``) and skinning (``:564-574``).
    ``model`` is a dict of torch tensors: v_template (V,3), shapedirs (V,3,10),
    posedirs (207,3V), J_regressor (24,V), lbs_weights (V,24), parents list.
    """
    nb = poses.shape[0]
    V = model['v_template'].shape[0]
    v_shaped = model['v_template'].unsqueeze(0) + torch.einsum('bl,mkl->bmk', betas, model['shapedirs'])
    J = torch.einsum('bik,ji->bjk', v_shaped, model['J_regressor'])
    R22 = rodrigues(poses[:, :66].reshape(-1, 3)).view(nb, 22, 3, 3)
    eye = torch.eye(3, dtype=poses.dtype)
    R = torch.cat([R22, eye.view(1, 1, 3, 3).expand(nb, 2, 3, 3)], dim=1)
    pose_feature = (R[:, 1:] - eye).reshape(nb, 207)
    v_posed = v_shaped + torch.matmul(pose_feature, model['posedirs']).view(nb, V, 3)

    parents = model['parents']
    rel = J.clone()
    rel[:, 1:] = J[:, 1:] - J[:, parents[1:]]
    G_R = [R[:, 0]]
    G_t = [rel[:, 0]]
    for j in range(1, 24):
        p = parents[j]
        G_R.append(torch.bmm(G_R[p], R[:, j]))
        G_t.append(torch.bmm(G_R[p], rel[:, j].unsqueeze(-1)).squeeze(-1) + G_t[p])
    G_R = torch.stack(G_R, dim=1)           # (nb,24,3,3)
    G_t = torch.stack(G_t, dim=1)           # (nb,24,3)  == posed joints
    A_t = G_t - torch.einsum('bjik,bjk->bji', G_R, J)
    # skinning: T = W . A ; v = T [v_posed; 1]
    W = model['lbs_weights']
    T_R = torch.einsum('vj,bjik->bvik', W, G_R)
    T_t = torch.einsum('vj,bji->bvi', W, A_t)
    verts = torch.einsum('bvik,bvk->bvi', T_R, v_posed) + T_t
    return {'verts': verts, 'joints24': G_t, 'A_R': G_R, 'A_t': A_t, 'v_posed': v_posed,
            'v_shaped': v_shaped, 'J': J, 'R': R}


def regress_joints(regressor, verts):
    """``vertices2joints`` (``smpl.py:603-620``): regressor (J,V), verts (nb,V,3)."""
    return torch.einsum('bik,ji->bjk', verts, regressor)


# ----------------------------------------------------------------------------
# Camera / activations  (mhmocap/transforms.py)
# ----------------------------------------------------------------------------
def camera_projection(pts3d, K, Kd=None):
    """``camera_projection_torch`` (``transforms.py:57-95``).

    pts3d (n,m,3), K (n,3,3).  uv = (xy/z) . K[:2,:2]^T + K[:2,2].  The optional
    5-coefficient distortion follows the CODE (``:78-90``): the y term uses
    ``2*Kd[3]*y*y`` (not the Brown-Conrady ``2*p2*x*y``).
    """
    z = pts3d[..., 2:]
    p = pts3d[..., :2] / z
    if Kd is not None:
        x, y = p[..., 0], p[..., 1]
        r = x * x + y * y
        rad = 1 + Kd[0] * r + Kd[1] * r * r + Kd[4] * r * r * r
        xx = x * rad + 2 * Kd[2] * x * y + Kd[3] * (r + 2 * x * x)
        yy = y * rad + 2 * Kd[3] * y * y + Kd[2] * (r + 2 * y * y)
        p = torch.stack([xx, yy], dim=-1)
    Kt = K.transpose(1, 2)
    return torch.bmm(p, Kt[:, :2, :2]) + Kt[:, 2:, :2]


def camera_inverse_projection(uvd, K):
    """``camera_inverse_projection_torch`` (``transforms.py:114-130``); uvd (n,m,3), K (n,3,3)."""
    Kt = K.transpose(1, 2)
    xy = uvd[..., 2:3] * ((uvd[..., :2] - Kt[:, 2:3, 0:2]) @ torch.linalg.inv(Kt[:, :2, :2]))
    return torch.cat([xy, uvd[..., 2:3]], dim=-1)


def compute_calibration_matrix(znear, zfar, cam_K, image_size):
    """NDC calibration 4x4 for PyTorch3D (``transforms.py:222-255``).

    ``image_size`` is (W, H).  Landscape uses fy for both axes, portrait fx,
    square the average (``:226-244``); x offset is stretched by the aspect ratio.
    """
    W, H = image_size
    if W > H:
        s1 = 2 * cam_K[1, 1] / H
        u = W / H
        w1 = u * (W - 2 * cam_K[0, 2]) / W
        h1 = (H - 2 * cam_K[1, 2]) / H
    elif H > W:
        s1 = 2 * cam_K[0, 0] / W
        u = H / W
        w1 = (W - 2 * cam_K[0, 2]) / W
        h1 = u * (H - 2 * cam_K[1, 2]) / H
    else:
        s1 = 2 * (cam_K[0, 0] + cam_K[1, 1]) / (W + H)
        w1 = (W - 2 * cam_K[0, 2]) / W
        h1 = (H - 2 * cam_K[1, 2]) / H
    f1 = zfar / (zfar - znear)
    f2 = -(zfar * znear) / (zfar - znear)
    return np.array([[s1, 0, w1, 0], [0, s1, h1, 0], [0, 0, f1, f2], [0, 0, 1, 0]], np.float32)


def get_focal(w, theta):
    """``transforms.py:262-264``."""
    return 0.5 * w / np.tan((np.pi * theta / 180.0) / 2.0)


def softplus(x):
    """Naive ``log(1 + exp(x))`` (``transforms.py:296-297``)."""
    return torch.log(1.0 + torch.exp(x))


# ----------------------------------------------------------------------------
# Losses / morphology  (mhmocap/losses.py, morphology.py)
# ----------------------------------------------------------------------------
def avg_depth_loss(y_pred, y_true, mask, eps=1e-3):
    """``build_avg_depth_loss_fn`` (``losses.py:19-30``): per (b,n) average
    log-disparity match, normaliser ``sum(mask) + 1``."""
    d_pred = mask * torch.log(torch.clamp(y_pred, eps))
    d_true = mask * torch.log(torch.clamp(y_true, eps))
    m = torch.sum(mask, dim=(2, 3))
    a = torch.sum(d_pred, dim=(2, 3)) / (m + 1)
    c = torch.sum(d_true, dim=(2, 3)) / (m + 1)
    return torch.sum(torch.square(a - c))


def masked_mse_loss(y1, y2, mask):
    """``build_masked_mse_loss_fn`` (``losses.py:33-40``)."""
    n = torch.sum(mask) + 1.0
    return torch.sum(torch.square(mask * (y1 - y2))) / n


def erode3(x):
    """``Erode2D(3)`` (``morphology.py:23-33``): 1 - clamp(conv3x3(x < 0.5)), zero
    padding applied to the ``x < 0.5`` map, so out-of-image neighbours never erode."""
    k = torch.ones(1, 1, 3, 3, dtype=torch.float32)
    return 1 - torch.clamp(F.conv2d(torch.lt(x, 0.5).float(), k, padding=1), 0, 1)


def erode5_twice3(x):
    """The optimiser's ``erode = Erode2D(3) o Erode2D(3)`` (``optimizer.py:306-309``); x (n,1,H,W)."""
    return erode3(erode3(x))


# ----------------------------------------------------------------------------
# One-Euro filter  (mhmocap/one_euro_filter.py + optimizer.py:664-675)
# ----------------------------------------------------------------------------
def one_euro_filter_sequence(y, min_cutoff, beta, frame_rate=25, d_cutoff=1.0):
    """numpy (T,...) -> filtered copy, with the optimiser's CUMULATIVE time quirk
    ``t_i = t_{i-1} + i / frame_rate`` (``optimizer.py:671``), ``dx0 = 0``,
    smoothing factor ``r / (r + 1)``, ``r = 2 pi cutoff t_e``
    (``one_euro_filter.py:7-9, 32-53``).  Arithmetic in the dtype of ``y`` like the
    reference (float32 arrays stay float32 under numpy 2 scalar promotion)."""
    y = np.array(y, copy=True)
    dt = y.dtype.type
    x_prev = y[0].copy()
    dx_prev = np.zeros_like(x_prev)
    t_prev = np.zeros_like(x_prev)
    t = np.zeros_like(x_prev)
    two_pi = 2 * math.pi
    for i in range(1, len(y)):
        t = t + (i / frame_rate)
        x = y[i].copy()
        t_e = t - t_prev
        r = two_pi * float(d_cutoff) * t_e
        a_d = r / (r + 1)
        dx = (x - x_prev) / t_e
        dx_hat = a_d * dx + (1 - a_d) * dx_prev
        cutoff = float(min_cutoff) + float(beta) * np.abs(dx_hat)
        r = two_pi * cutoff * t_e
        a = r / (r + 1)
        x_hat = a * x + (1 - a) * x_prev
        x_prev, dx_prev, t_prev = x_hat, dx_hat, t
        y[i] = x_hat
    return y.astype(np.float32)

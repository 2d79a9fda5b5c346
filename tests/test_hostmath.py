"""Analytic derivatives in csrc/mh_math.cuh (host build) vs torch autograd of the oracle.  CPU only."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import refmath as rm, raster
import hostmath

FP = ctypes.POINTER(ctypes.c_float)


def P(a):
    return a.ctypes.data_as(FP)


@pytest.fixture(scope='module')
def lib():
    return hostmath.load()


def test_rodrigues_fwd_bwd(lib):
    rng = np.random.default_rng(0)
    rs = np.concatenate([rng.normal(0, 1, (20, 3)), np.zeros((1, 3)), [[3, 0, 0]], rng.normal(0, 1e-4, (3, 3))]).astype(np.float32)
    for r in rs:
        R = np.zeros(9, np.float32)
        lib.hm_rodrigues(P(r), P(R))
        rt = torch.from_numpy(r[None]).requires_grad_(True)
        Rt = rm.rodrigues(rt)
        assert np.abs(R - Rt.detach().numpy().ravel()).max() < 2e-6
        G = rng.normal(0, 1, 9).astype(np.float32)
        (Rt.view(-1) * torch.from_numpy(G)).sum().backward()
        g = np.zeros(3, np.float32)
        lib.hm_rodrigues_bwd(P(r), P(G), P(g))
        ref = rt.grad.numpy().ravel()
        assert np.abs(g - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), (r, g, ref)


def test_pose_stage_fwd_bwd(lib, model):
    rng = np.random.default_rng(1)
    mt = {a: (torch.from_numpy(v) if v.dtype == np.float32 else v) for a, v in model.items()}
    mt['parents'] = [int(p) for p in model['parents']]
    theta = rng.normal(0, 0.5, (1, 72)).astype(np.float32)
    beta = rng.normal(0, 0.5, (1, 10)).astype(np.float32)
    th = torch.from_numpy(theta).requires_grad_(True)
    be = torch.from_numpy(beta)
    out = rm.smpl_forward(mt, be, th)
    J = out['J'].detach().clone().requires_grad_(True)
    # re-run the pose stage alone with J as a leaf
    m2 = dict(mt)
    A_ref = torch.cat([out['A_R'], out['A_t'].unsqueeze(-1)], dim=-1)[0].detach().numpy()      # (24,3,4)
    A = np.zeros(24 * 12, np.float32); pf = np.zeros(207, np.float32); R = np.zeros(24 * 9, np.float32)
    Jn = out['J'][0].detach().numpy().astype(np.float32).copy()
    lib.hm_pose_forward(P(theta[0]), P(Jn), P(A), P(pf), P(R))
    assert np.abs(A.reshape(24, 3, 4) - A_ref).max() < 1e-5
    pf_ref = (out['R'][0, 1:] - torch.eye(3)).reshape(-1).detach().numpy()
    assert np.abs(pf - pf_ref).max() < 1e-6
    # backward: random cotangents on A and pf; reference through autograd with J treated as a function of a leaf
    dA = rng.normal(0, 1, (24, 3, 4)).astype(np.float32)
    dpf = rng.normal(0, 1, 207).astype(np.float32)

    def pose_only(th_, J_):
        Rm = torch.cat([rm.rodrigues(th_[:, :66].reshape(-1, 3)).view(1, 22, 3, 3), torch.eye(3).view(1, 1, 3, 3).expand(1, 2, 3, 3)], 1)
        par = mt['parents']
        GR = [Rm[:, 0]]; Gt = [J_[:, 0]]
        for j in range(1, 24):
            p = par[j]
            GR.append(torch.bmm(GR[p], Rm[:, j]))
            Gt.append(torch.bmm(GR[p], (J_[:, j] - J_[:, p]).unsqueeze(-1)).squeeze(-1) + Gt[p])
        GR = torch.stack(GR, 1); Gt = torch.stack(Gt, 1)
        At = Gt - torch.einsum('bjik,bjk->bji', GR, J_)
        return torch.cat([GR, At.unsqueeze(-1)], -1), (Rm[:, 1:] - torch.eye(3)).reshape(1, 207)

    th2 = torch.from_numpy(theta).requires_grad_(True)
    J2 = torch.from_numpy(Jn[None]).requires_grad_(True)
    A_t, pf_t = pose_only(th2, J2)
    ((A_t[0] * torch.from_numpy(dA)).sum() + (pf_t[0] * torch.from_numpy(dpf)).sum()).backward()
    dth = np.zeros(72, np.float32); dJ = np.zeros(72, np.float32)
    lib.hm_pose_backward(P(theta[0]), P(Jn), P(dA.ravel().copy()), P(dpf), P(dth), P(dJ))
    ref_th = th2.grad.numpy().ravel(); ref_J = J2.grad.numpy().ravel()
    assert np.abs(dth - ref_th).max() <= 1e-4 * np.abs(ref_th).max()
    assert np.abs(dJ - ref_J).max() <= 1e-4 * np.abs(ref_J).max()
    assert np.all(dth[66:] == 0)


@pytest.mark.parametrize('use_kd', [False, True])
def test_projection_fwd_bwd(lib, use_kd):
    rng = np.random.default_rng(2)
    K = np.array([[1000, 3.0, 640], [0.5, 990, 360], [0, 0, 1]], np.float32)
    Kd = np.array([0.1, 0.01, 0.001, 0.002, 0.0001], np.float32) if use_kd else None
    for _ in range(10):
        p = (rng.normal(0, 1, 3) + np.array([0, 0, 4])).astype(np.float32)
        uv = np.zeros(2, np.float32)
        lib.hm_project(P(p), P(K), P(Kd) if use_kd else None, P(uv))
        pt = torch.from_numpy(p[None, None]).requires_grad_(True)
        ref = rm.camera_projection(pt, torch.from_numpy(K)[None], Kd)
        assert np.abs(uv - ref.detach().numpy().ravel()).max() < 2e-3
        gu, gv = rng.normal(0, 1, 2)
        (ref[0, 0, 0] * gu + ref[0, 0, 1] * gv).backward()
        gp = np.zeros(3, np.float32)
        lib.hm_project_bwd(P(p), P(K), P(Kd) if use_kd else None, ctypes.c_float(gu), ctypes.c_float(gv), P(gp))
        r = pt.grad.numpy().ravel()
        assert np.abs(gp - r).max() <= 1e-4 * np.abs(r).max()


def _oracle_frag(v, px, py):
    """Same arithmetic as oracle.raster.rasterize for a single (pixel, face) pair, differentiable."""
    x0, y0, z0, x1, y1, z1, x2, y2, z2 = [v[i] for i in range(9)]
    pxn = torch.tensor(px); pyn = torch.tensor(py)
    area = raster._edge(x2, y2, x0, y0, x1, y1)
    den = area + raster.K_EPS
    w0 = raster._edge(pxn, pyn, x1, y1, x2, y2) / den
    w1 = raster._edge(pxn, pyn, x2, y2, x0, y0) / den
    w2 = raster._edge(pxn, pyn, x0, y0, x1, y1) / den
    c0, c1, c2 = (torch.clamp(w, 0.0, 1.0) for w in (w0, w1, w2))
    bs = torch.clamp(c0 + c1 + c2, min=1e-5)
    pz = (c0 * z0 + c1 * z1 + c2 * z2) / bs
    d = torch.minimum(torch.minimum(raster._seg_dist(pxn, pyn, x0, y0, x1, y1), raster._seg_dist(pxn, pyn, x0, y0, x2, y2)),
                      raster._seg_dist(pxn, pyn, x1, y1, x2, y2))
    inside = bool((w0 > 0) & (w1 > 0) & (w2 > 0))
    return pz, d, inside


def test_face_eval_fwd_bwd(lib):
    rng = np.random.default_rng(3)
    n_in = 0
    for it in range(300):
        c = rng.uniform(-0.5, 0.5, 2)
        v = np.zeros(9, np.float32)
        for i in range(3):
            v[3 * i:3 * i + 2] = c + rng.normal(0, 0.01, 2)
            v[3 * i + 2] = 4 + rng.normal(0, 0.1)
        px, py = (c + rng.normal(0, 0.006 if it % 2 else 0.012, 2)).astype(np.float32)
        vt = torch.from_numpy(v).requires_grad_(True)
        pz, d, inside = _oracle_frag(vt, float(px), float(py))
        out = np.zeros(4, np.float32)
        lib.hm_face_eval(P(v), ctypes.c_float(px), ctypes.c_float(py), P(out))
        assert abs(out[0] - float(pz.detach())) <= 1e-5 * abs(float(pz.detach()))
        assert abs(out[1] - float(d.detach())) <= 1e-4 * abs(float(d.detach())) + 3e-9      # NDC^2; compared against blur >= 2e-5
        assert bool(out[2]) == inside
        n_in += inside
        gz, gd = rng.normal(0, 1, 2)
        (pz * gz + d * gd).backward()
        g = np.zeros(9, np.float32)
        lib.hm_face_bwd(P(v), ctypes.c_float(px), ctypes.c_float(py), ctypes.c_float(gz), ctypes.c_float(gd), P(g))
        ref = vt.grad.numpy()
        assert np.abs(g - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-6, (it, g, ref)
    assert n_in > 10

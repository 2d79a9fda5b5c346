"""Torch-CPU restatement of the PyTorch3D mesh rasteriser + soft-silhouette shader.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: the
algorithm lives in the un-vendored third-party dependency ``pytorch3d`` (unpinned
in the reference's ``environment.yml:13``).  This file restates the published
*naive* rasterisation semantics (``rasterize_meshes`` / ``geometry_utils`` /
``blending.sigmoid_alpha_blend`` upstream) that the
reference reaches through ``optimizer.py:210-232, 428-430, 447-448``:

* view transform R = diag(-1,-1,1), T = 0, then the 4x4 NDC calibration of
  ``transforms.py:222-255``; z stays view-space depth;
* pixel centre of output (row y, col x) is the NDC point of the FLIPPED index
  (W-1-x, H-1-y); short image side spans [-1, 1];
* per (pixel, face): skip if max z < 0, if the pixel is outside the face bbox
  inflated by sqrt(blur_radius), or if |area| <= 1e-8; barycentrics
  w = edge / (area + 1e-8); clipped (clamp to [0,1], renormalise by
  max(sum, 1e-5)) because blur_radius > 0; pz = sum(w_clip * z); skip if pz < 0;
  dist = min squared point-segment distance over the 3 edges; inside = all
  unclipped w > 0; skip if not inside and dist >= blur_radius;
* keep the K smallest pz per pixel (ascending); empty slots are -1;
* silhouette alpha = 1 - prod_k (1 - sigmoid(-signed_dist_k / sigma)), sigma = 1e-4.

Gradients come from torch autograd through exactly this arithmetic (discrete
choices -- validity, top-K membership, nearest edge -- carry no gradient).
"""
import numpy as np
import torch

K_EPS = 1e-8


def pixel_centers_ndc(H, W):
    """NDC coordinate of every output column / row centre (float32 arithmetic as upstream)."""
    def ndc(i, s1, s2):
        r = np.float32(2.0 * s1 / s2) if s1 > s2 else np.float32(2.0)
        i = i.astype(np.float32)
        return (-r / np.float32(2.0) + (r * i + r / np.float32(2.0)) / np.float32(s1)).astype(np.float32)
    xs = ndc(W - 1 - np.arange(W), W, H)
    ys = ndc(H - 1 - np.arange(H), H, W)
    return xs, ys


def world_to_ndc(verts, Kndc):
    """verts (...,3) world -> (x_ndc, y_ndc, z_view); R = diag(-1,-1,1), T = 0 (``optimizer.py:204-207``)."""
    xv = -verts[..., 0]
    yv = -verts[..., 1]
    zv = verts[..., 2]
    x = (Kndc[0, 0] * xv + Kndc[0, 2] * zv) / zv
    y = (Kndc[1, 1] * yv + Kndc[1, 2] * zv) / zv
    return torch.stack([x, y, zv], dim=-1)


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_dist(px, py, ax, ay, bx, by):
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    safe = torch.where(l2 <= K_EPS, torch.ones_like(l2), l2)
    t = (bax * (px - ax) + bay * (py - ay)) / safe
    t = torch.clamp(t, 0.0, 1.0)
    qx = ax + t * bax
    qy = ay + t * bay
    d = (px - qx) ** 2 + (py - qy) ** 2
    d_end = (px - bx) ** 2 + (py - by) ** 2
    return torch.where(l2 <= K_EPS, d_end, d)


def rasterize(verts_ndc, faces, H, W, blur_radius, K):
    """Rasterise ONE mesh.  verts_ndc (V,3) [x_ndc, y_ndc, z]; faces (F,3) long.

    Returns dict of (H,W,K) tensors: ``zbuf``, ``dists`` (signed), ``pix_to_face``
    (long, -1 empty) and ``bary`` (H,W,K,3) clipped barycentrics; differentiable
    w.r.t. ``verts_ndc``.
    """
    xs_np, ys_np = pixel_centers_ndc(H, W)
    fv = verts_ndc[faces]                                   # (F,3,3)
    with torch.no_grad():
        r = float(np.sqrt(np.float32(blur_radius)))
        fx, fy, fz = fv[..., 0], fv[..., 1], fv[..., 2]
        xmin = (fx.min(1).values - r).numpy(); xmax = (fx.max(1).values + r).numpy()
        ymin = (fy.min(1).values - r).numpy(); ymax = (fy.max(1).values + r).numpy()
        zmax = fz.max(1).values.numpy()
        # xs/ys are DEcreasing in the output index; find conservative index ranges
        xs_inc = xs_np[::-1]; ys_inc = ys_np[::-1]
        c_lo = W - np.searchsorted(xs_inc, xmax, side='right')
        c_hi = W - np.searchsorted(xs_inc, xmin, side='left')
        r_lo = H - np.searchsorted(ys_inc, ymax, side='right')
        r_hi = H - np.searchsorted(ys_inc, ymin, side='left')
        nx = np.clip(c_hi - c_lo, 0, None); ny = np.clip(r_hi - r_lo, 0, None)
        cnt = nx * ny
        cnt[zmax < 0] = 0
        total = int(cnt.sum())
    out_shape = (H * W, K)
    zbuf = torch.full(out_shape, -1.0, dtype=verts_ndc.dtype)
    dists = torch.full(out_shape, -1.0, dtype=verts_ndc.dtype)
    p2f = torch.full(out_shape, -1, dtype=torch.long)
    bary = torch.full(out_shape + (3,), -1.0, dtype=verts_ndc.dtype)
    if total == 0:
        return {'zbuf': zbuf.view(H, W, K), 'dists': dists.view(H, W, K),
                'pix_to_face': p2f.view(H, W, K), 'bary': bary.view(H, W, K, 3)}
    with torch.no_grad():
        fidx = np.repeat(np.arange(len(cnt)), cnt)
        start = np.cumsum(cnt) - cnt
        local = np.arange(total) - np.repeat(start, cnt)
        nxr = np.repeat(np.maximum(nx, 1), cnt)
        col = np.repeat(c_lo, cnt) + local % nxr
        row = np.repeat(r_lo, cnt) + local // nxr
        fidx_t = torch.from_numpy(fidx)
        pxn = torch.from_numpy(xs_np[col].copy())
        pyn = torch.from_numpy(ys_np[row].copy())
        pix = torch.from_numpy((row * W + col).astype(np.int64))
    v = fv[fidx_t]                                           # (P,3,3)
    x0, y0, z0 = v[:, 0, 0], v[:, 0, 1], v[:, 0, 2]
    x1, y1, z1 = v[:, 1, 0], v[:, 1, 1], v[:, 1, 2]
    x2, y2, z2 = v[:, 2, 0], v[:, 2, 1], v[:, 2, 2]
    area = _edge(x2, y2, x0, y0, x1, y1)
    den = area + K_EPS
    w0 = _edge(pxn, pyn, x1, y1, x2, y2) / den
    w1 = _edge(pxn, pyn, x2, y2, x0, y0) / den
    w2 = _edge(pxn, pyn, x0, y0, x1, y1) / den
    c0 = torch.clamp(w0, 0.0, 1.0); c1 = torch.clamp(w1, 0.0, 1.0); c2 = torch.clamp(w2, 0.0, 1.0)
    bs = torch.clamp(c0 + c1 + c2, min=1e-5)
    c0, c1, c2 = c0 / bs, c1 / bs, c2 / bs
    pz = c0 * z0 + c1 * z1 + c2 * z2
    d01 = _seg_dist(pxn, pyn, x0, y0, x1, y1)
    d02 = _seg_dist(pxn, pyn, x0, y0, x2, y2)
    d12 = _seg_dist(pxn, pyn, x1, y1, x2, y2)
    dist = torch.minimum(torch.minimum(d01, d02), d12)
    with torch.no_grad():
        inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
        r32 = np.float32(r)
        xmn = torch.minimum(torch.minimum(x0, x1), x2) - r32
        xmx = torch.maximum(torch.maximum(x0, x1), x2) + r32
        ymn = torch.minimum(torch.minimum(y0, y1), y2) - r32
        ymx = torch.maximum(torch.maximum(y0, y1), y2) + r32
        outside_bbox = (pxn > xmx) | (pxn < xmn) | (pyn > ymx) | (pyn < ymn)
        zero_area = (area <= K_EPS) & (area >= -K_EPS)
        valid = (~outside_bbox) & (~zero_area) & (pz >= 0) & (inside | (dist < np.float32(blur_radius)))
        vidx = torch.nonzero(valid).squeeze(1)
        if vidx.numel() == 0:
            return {'zbuf': zbuf.view(H, W, K), 'dists': dists.view(H, W, K),
                    'pix_to_face': p2f.view(H, W, K), 'bary': bary.view(H, W, K, 3)}
        pzv = pz[vidx]; pixv = pix[vidx]
        # rank within pixel by (pz, face index)
        o1 = torch.argsort(fidx_t[vidx], stable=True)
        o2 = torch.argsort(pzv[o1], stable=True)
        o = o1[o2]
        o3 = torch.argsort(pixv[o], stable=True)
        o = o[o3]
        pix_sorted = pixv[o]
        first = torch.searchsorted(pix_sorted, pix_sorted, right=False)
        rank = torch.arange(len(o)) - first
        keep = rank < K
        sel = vidx[o[keep]]
        slot = rank[keep]
        dst = pix_sorted[keep]
    signed = torch.where(inside, -dist, dist)
    zbuf = zbuf.index_put((dst, slot), pz[sel])
    dists = dists.index_put((dst, slot), signed[sel])
    p2f[dst, slot] = fidx_t[sel]
    bary = bary.index_put((dst, slot), torch.stack([c0[sel], c1[sel], c2[sel]], dim=-1))
    return {'zbuf': zbuf.view(H, W, K), 'dists': dists.view(H, W, K),
            'pix_to_face': p2f.view(H, W, K), 'bary': bary.view(H, W, K, 3)}


def silhouette_alpha(dists, pix_to_face, sigma=1e-4):
    """``SoftSilhouetteShader`` -> ``sigmoid_alpha_blend`` with default BlendParams (sigma 1e-4)."""
    mask = (pix_to_face >= 0).to(dists.dtype)
    prob = torch.sigmoid(-dists / sigma) * mask
    return 1.0 - torch.prod(1.0 - prob, dim=-1)


def render_person(verts_world, faces, Kndc, H, W):
    """Both rasterisations the optimiser runs for one person-frame
    (``optimizer.py:211-232``): depth (K=8, blur 1e-4) -> zbuf[...,0]; silhouette
    (K=4, blur 2e-5) -> alpha.  verts_world (V,3) camera-space metres."""
    vn = world_to_ndc(verts_world, Kndc)
    fr_d = rasterize(vn, faces, H, W, 1e-4, 8)
    fr_s = rasterize(vn, faces, H, W, 2e-5, 4)
    return fr_d['zbuf'][..., 0], silhouette_alpha(fr_s['dists'], fr_s['pix_to_face'])

"""Discrete edge cases of the rasteriser restatement (``oracle/raster.py``, PARITY UNPINNED: PyTorch3D's source is not available
offline) against HAND-COMPUTED values, so that the semantics the CUDA rasteriser is held to are at least
self-documented: degenerate faces, vertices behind the camera, exact depth ties, top-K membership, square / landscape / portrait
pixel-centre conventions, the blur-radius test and the clipped barycentric depth outside a face."""
import numpy as np
import torch

from oracle import raster


def _mesh(tris):
    v = torch.tensor(np.asarray(tris, np.float32).reshape(-1, 3))
    f = torch.arange(v.shape[0]).view(-1, 3)
    return v, f


def test_pixel_centres_follow_the_short_side_convention():
    # square: [-1, 1] on both axes, +X to the LEFT and +Y UP (index 0 is the largest coordinate)
    xs, ys = raster.pixel_centers_ndc(4, 4)
    assert np.allclose(xs, [0.75, 0.25, -0.25, -0.75]) and np.allclose(ys, [0.75, 0.25, -0.25, -0.75])
    # landscape 8 x 4: the short side (y) spans [-1, 1], the long side [-2, 2]
    xs, ys = raster.pixel_centers_ndc(4, 8)
    assert np.allclose(ys, [0.75, 0.25, -0.25, -0.75]) and np.allclose(xs, 1.75 - 0.5 * np.arange(8))
    # portrait 4 x 8
    xs, ys = raster.pixel_centers_ndc(8, 4)
    assert np.allclose(xs, [0.75, 0.25, -0.25, -0.75]) and np.allclose(ys, 1.75 - 0.5 * np.arange(8))


def test_one_face_covers_the_pixels_whose_centres_are_inside():
    # right triangle with vertices on pixel corners of a 4 x 4 image; blur 0+ (tiny) keeps only inside pixels
    v, f = _mesh([[[1.0, 1.0, 2.0], [-1.0, 1.0, 2.0], [1.0, -1.0, 2.0]]])
    out = raster.rasterize(v, f, 4, 4, 1e-12, 1)
    cov = (out['pix_to_face'][..., 0] >= 0).numpy()
    xs, ys = raster.pixel_centers_ndc(4, 4)
    # strictly inside <=> x + y > 0 (hypotenuse x + y = 0); a centre exactly ON the hypotenuse is not inside (barycentric 0 is not > 0)
    # but its edge distance 0 is below any positive blur radius, so it gets a fragment too -- with signed distance +0
    expect = np.array([[(x + y) >= 0 for x in xs] for y in ys])
    assert np.array_equal(cov, expect)
    assert np.allclose(out['zbuf'][..., 0].numpy()[cov], 2.0)
    d = out['dists'][..., 0].numpy()
    strictly = np.array([[(x + y) > 0 for x in xs] for y in ys])
    assert (d[strictly] < 0).all() and (d[cov & ~strictly] == 0).all()    # signed distance is negative inside


def test_degenerate_and_behind_camera_faces_are_skipped():
    good = [[0.9, 0.9, 3.0], [-0.9, 0.9, 3.0], [0.9, -0.9, 3.0]]
    zero_area = [[0.5, 0.5, 1.0], [0.0, 0.0, 1.0], [-0.5, -0.5, 1.0]]      # collinear: |area| <= 1e-8, nearer than `good`
    behind = [[0.9, 0.9, -1.0], [-0.9, 0.9, -2.0], [0.9, -0.9, -3.0]]      # every z < 0
    v, f = _mesh([good, zero_area, behind])
    out = raster.rasterize(v, f, 8, 8, 1e-4, 4)
    faces_seen = np.unique(out['pix_to_face'].numpy())
    assert set(faces_seen.tolist()) == {-1, 0}
    # a face with ONE vertex behind the camera is kept (max z >= 0); its fragments with interpolated depth < 0 are dropped
    part = [[0.9, 0.9, 2.0], [-0.9, 0.9, 2.0], [0.9, -0.9, -6.0]]
    v, f = _mesh([part])
    out = raster.rasterize(v, f, 8, 8, 1e-12, 1)
    z = out['zbuf'][..., 0].numpy()
    cov = out['pix_to_face'][..., 0].numpy() >= 0
    assert cov.any() and (z[cov] >= 0).all()
    xs, ys = raster.pixel_centers_ndc(8, 8)
    # barycentric weight of vertex 2 at (x, y) is w2 = (0.9 - y) / 1.8 ; depth 2 - 8 w2 < 0 <=> w2 > 0.25 <=> y < 0.45
    for iy, y in enumerate(ys):
        for ix, x in enumerate(xs):
            inside = (x < 0.9) and (y < 0.9) and (x + y >= 0)        # centres ON the hypotenuse count: distance 0 < blur
            assert cov[iy, ix] == (inside and y > 0.45), (x, y)


def test_depth_ties_go_to_the_lower_face_index_and_top_k_keeps_the_nearest():
    tri = lambda z: [[0.9, 0.9, z], [-0.9, 0.9, z], [0.9, -0.9, z]]        # noqa: E731
    # five coincident-in-xy faces; faces 1 and 3 at exactly the same depth
    v, f = _mesh([tri(5.0), tri(2.0), tri(4.0), tri(2.0), tri(3.0)])
    out = raster.rasterize(v, f, 4, 4, 1e-12, 4)
    p2f = out['pix_to_face'].numpy(); z = out['zbuf'].numpy()
    cov = p2f[..., 0] >= 0
    assert cov.sum() == 10                                                 # centres with x + y >= 0 (4 of them ON the hypotenuse)
    assert (p2f[cov] == np.array([1, 3, 4, 2])).all()                      # nearest first, tie -> lower index, face 0 (farthest) dropped
    assert np.allclose(z[cov], [2.0, 2.0, 3.0, 4.0])
    out1 = raster.rasterize(v, f, 4, 4, 1e-12, 1)
    assert (out1['pix_to_face'].numpy()[cov][:, 0] == 1).all()


def test_blur_radius_admits_outside_pixels_with_clipped_depth():
    # 8 x 8 pixels, pixel pitch 0.25; a small triangle around the origin with a depth slope; pixels whose squared distance to the
    # nearest edge is below the blur radius get a fragment whose depth is the CLIPPED barycentric interpolation
    v, f = _mesh([[[0.3, 0.3, 2.0], [-0.3, 0.3, 4.0], [0.3, -0.3, 6.0]]])
    blur = 0.04                                                             # radius 0.2
    out = raster.rasterize(v, f, 8, 8, blur, 1)
    xs, ys = raster.pixel_centers_ndc(8, 8)
    cov = out['pix_to_face'][..., 0].numpy() >= 0
    d = out['dists'][..., 0].numpy(); z = out['zbuf'][..., 0].numpy()

    def seg(p, a, b):
        a, b, p = np.array(a), np.array(b), np.array(p)
        t = np.clip(np.dot(p - a, b - a) / np.dot(b - a, b - a), 0, 1)
        return float(np.sum((p - (a + t * (b - a))) ** 2))
    V = [(0.3, 0.3), (-0.3, 0.3), (0.3, -0.3)]
    for iy, y in enumerate(ys):
        for ix, x in enumerate(xs):
            inside = (x < 0.3) and (y < 0.3) and (x + y > 0)
            dist = min(seg((x, y), V[0], V[1]), seg((x, y), V[0], V[2]), seg((x, y), V[1], V[2]))
            in_bbox = (-0.5 <= x <= 0.5) and (-0.5 <= y <= 0.5)
            assert cov[iy, ix] == (in_bbox and (inside or dist < blur)), (x, y)
            if cov[iy, ix]:
                assert abs(abs(d[iy, ix]) - dist) < 1e-6 and (d[iy, ix] < 0) == inside
    # the pixel centre (0.375, 0.375) lies beyond vertex 0: both other barycentrics clip to 0 -> depth of vertex 0 exactly
    iy, ix = list(ys).index(0.375), list(xs).index(0.375)
    assert cov[iy, ix] and abs(z[iy, ix] - 2.0) < 1e-6
    # (0.125, 0.375) is above edge 0-1: the weight of vertex 2 clips to 0, depth is interpolated along the edge: w1 = (0.3 - x) / 0.6
    ix2 = list(xs).index(0.125)
    w1 = (0.3 - 0.125) / 0.6
    w0 = 1 - w1 - (0.3 - 0.375) / 0.6                                       # unclipped w0 = 1 - w1 - w2, w2 < 0 is clipped away
    assert cov[iy, ix2] and abs(z[iy, ix2] - (w0 * 2.0 + w1 * 4.0) / (w0 + w1)) < 1e-5


def test_silhouette_alpha_of_stacked_fragments():
    d = torch.tensor([[[[-1e-4, 0.0, 2e-4, -1.0]]]])                        # (1,1,1,4) signed squared distances
    p2f = torch.tensor([[[[0, 1, 2, -1]]]])                                 # the last slot is empty
    a = raster.silhouette_alpha(d, p2f)
    p = [1 / (1 + np.exp(-1.0)), 0.5, 1 / (1 + np.exp(2.0))]
    assert abs(float(a) - (1 - (1 - p[0]) * (1 - p[1]) * (1 - p[2]))) < 1e-6

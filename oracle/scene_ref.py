"""numpy/cv2 restatement of the host-side scene-geometry update inside ``fit()``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
``fhsog.py:180-202`` (masked temporal median), ``utils.py:91-135`` (fill-in) and
``utils.py:174-209`` (depth post-processing) of the reference.
"""
import numpy as np


def aggregate_scene_median(depths, images, backmasks):
    """depths (T,H,W) f32, images (T,H,W,3) u8 or None, backmasks (T,H,W) {0,1}
    -> (bkg_img u8 | None, bkg_depth f32, mask bool).  ``np.ma.median`` over T of
    the background-only samples (``fhsog.py:189-202``); pixels never seen as
    background come out masked (depth data there is whatever numpy leaves)."""
    img = None
    if images is not None:
        m3 = np.tile(backmasks[..., np.newaxis] == 0, (1, 1, 1, 3))
        img = np.ma.median(np.ma.array(images, mask=m3), axis=0).data.astype(np.uint8)
    md = np.ma.median(np.ma.array(depths, mask=backmasks == 0), axis=0)
    return img, md.data.astype(np.float32), (md.mask == 0)


def fillin_values(x, mask, filter_size, metric='median'):
    """One sweep of the neighbourhood fill-in (``utils.py:91-135``): every
    masked-out pixel with at least one valid pixel in its (clipped) window takes
    the ``metric`` of the valid ones; validity is read from the INPUT mask but
    values from the array being updated in raster order (``nx``), as upstream."""
    fm = getattr(np, metric)
    nx = x.copy()
    nmask = mask.copy()
    R, C = nx.shape[:2]
    k = filter_size // 2
    rows, cols = np.nonzero(mask == 0)
    for r, c in zip(rows, cols):
        r0, r1, c0, c1 = max(0, r - k), min(R, r + k + 1), max(0, c - k), min(C, c + k + 1)
        m = mask[r0:r1, c0:c1] > 0
        if m.any():
            nx[r, c] = fm(nx[r0:r1, c0:c1][m, ...], axis=0)
            nmask[r, c] = 1
    return nx, nmask


def postprocess_depthmap(depth, mask=None, fillin_ksize=7, use_bilateral_filter=False):
    """``utils.py:174-209``: optional bilateral(9, 0.05, 25) on disparity, Sobel
    edge mask on disparity + depth (threshold 3x mean of the std-normalised sum),
    3x3 erosion x2, then fill-in sweeps until the mask is full."""
    import cv2
    if use_bilateral_filter:
        d = cv2.bilateralFilter(1.0 / np.clip(depth, 0.01, 100), 9, sigmaColor=0.05, sigmaSpace=25)
        depth = 1.0 / np.clip(d, 0.01, 100)
    disp = 1.0 / np.clip(depth, 0.1, 100)
    sob = lambda a: (np.abs(cv2.Sobel(a, cv2.CV_32F, 1, 0, ksize=3)) + np.abs(cv2.Sobel(a, cv2.CV_32F, 0, 1, ksize=3)))
    s_disp, s_depth = sob(disp), sob(depth)
    g = s_disp / s_disp.std() + s_depth / s_depth.std()
    edges = (g > 3 * g.mean()).astype(disp.dtype)
    dmask = cv2.erode((1 - edges), np.ones((3, 3)), iterations=2)
    if mask is not None:
        dmask = dmask * mask
    new_depth, new_mask = depth, dmask
    while new_mask.min() < 1:
        new_depth, new_mask = fillin_values(new_depth, new_mask, filter_size=fillin_ksize)
    return new_depth

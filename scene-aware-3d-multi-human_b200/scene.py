"""Host side of the scene-geometry update that ``fit()`` runs every cycle >= 30.

Reference functions replaced: ``aggegrate_scene_geometry_median`` (``mhmocap/fhsog.py:180-202``),
``postprocess_depthmap`` (``mhmocap/utils.py:174-209``) and ``fillin_values`` (``mhmocap/utils.py:91-135``).  The
reference walks every pixel in a Python loop; a hole pixel only ever reads pixels that were valid BEFORE the sweep
(``utils.py:127-131``: validity comes from the input mask and valid pixels are never rewritten), so one sweep is
order-independent and is evaluated here for all hole pixels at once.
"""
import numpy as np


def aggregate_scene_geometry_median(depths, images, backmasks):
    """depths (T,H,W) f32, images (T,H,W,3) u8 | None, backmasks (T,H,W) in {0,1}: per-pixel median over the frames
    in which the pixel is background.  Returns (image u8 | None, depth f32, mask bool)."""
    hidden = backmasks == 0
    image = None
    if images is not None:
        med = np.ma.median(np.ma.array(images, mask=np.broadcast_to(hidden[..., None], images.shape)), axis=0)
        image = med.data.astype(np.uint8)
    med = np.ma.median(np.ma.array(depths, mask=hidden), axis=0)
    return image, med.data.astype(np.float32), (np.ma.getmaskarray(med) == 0)


def fillin_values(x, mask, filter_size, metric='median'):
    """One fill-in sweep: every pixel with ``mask == 0`` that has a valid pixel inside its ``filter_size`` window
    (clipped at the border) takes the ``metric`` of the valid ones and becomes valid."""
    if tuple(x.shape[:2]) != tuple(mask.shape):
        raise ValueError(f'x is {x.shape[:2]} but the mask is {mask.shape}')
    if filter_size < 2:
        raise ValueError(f'the window must span at least 2 pixels, got {filter_size}')
    if metric not in ('median', 'mean', 'max', 'min'):
        raise ValueError(f'unknown reduction {metric!r}')
    reduce_fn = {'median': np.nanmedian, 'mean': np.nanmean, 'max': np.nanmax, 'min': np.nanmin}[metric]
    k = filter_size // 2
    H, W = mask.shape
    nx, nmask = x.copy(), mask.copy()
    valid = mask > 0
    holes = ~mask.astype(bool)
    if not holes.any():
        return nx, nmask
    # holes that see at least one valid pixel
    vpad = np.pad(valid, k, constant_values=False)
    win_valid = np.lib.stride_tricks.sliding_window_view(vpad, (2 * k + 1, 2 * k + 1))
    rows, cols = np.nonzero(holes)
    wv = win_valid[rows, cols]                                        # (P, k', k')
    reach = wv.reshape(len(rows), -1).any(axis=1)
    rows, cols, wv = rows[reach], cols[reach], wv[reach]
    if len(rows) == 0:
        return nx, nmask
    chans = x.reshape(H, W, -1)
    out = nx.reshape(H, W, -1)
    for ch in range(chans.shape[2]):
        xp = np.pad(chans[..., ch].astype(np.float64), k, constant_values=np.nan)
        win = np.lib.stride_tricks.sliding_window_view(xp, (2 * k + 1, 2 * k + 1))[rows, cols]
        vals = np.where(wv, win, np.nan).reshape(len(rows), -1)
        out[rows, cols, ch] = reduce_fn(vals, axis=1).astype(x.dtype) if x.dtype != np.uint8 else reduce_fn(vals, axis=1)
    nmask[rows, cols] = 1
    return nx, nmask


def postprocess_depthmap(depth, mask=None, fillin_ksize=7, use_bilateral_filter=False):
    """Remove flying pixels from a depth map: optional bilateral filter on the disparity, Sobel edges of disparity and
    depth (threshold 3 x the mean of their std-normalised sum), two 3x3 erosions of the keep-mask, then fill-in
    sweeps until every pixel is valid."""
    import cv2
    if use_bilateral_filter:
        disp = cv2.bilateralFilter(1.0 / np.clip(depth, 0.01, 100), 9, sigmaColor=0.05, sigmaSpace=25)
        depth = 1.0 / np.clip(disp, 0.01, 100)
    disp = 1.0 / np.clip(depth, 0.1, 100)

    def grad_mag(a):
        return np.abs(cv2.Sobel(a, cv2.CV_32F, 1, 0, ksize=3)) + np.abs(cv2.Sobel(a, cv2.CV_32F, 0, 1, ksize=3))

    g_disp, g_depth = grad_mag(disp), grad_mag(depth)
    g = g_disp / g_disp.std() + g_depth / g_depth.std()
    edges = (g > 3 * g.mean()).astype(disp.dtype)
    keep = cv2.erode(1 - edges, np.ones((3, 3)), iterations=2)
    if mask is not None:
        keep = keep * mask
    new_depth, new_mask = depth, keep
    while new_mask.min() < 1:
        new_depth, new_mask = fillin_values(new_depth, new_mask, filter_size=fillin_ksize)
    return new_depth

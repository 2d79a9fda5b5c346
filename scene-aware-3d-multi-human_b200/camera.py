"""Camera helpers of the optimiser's constructor (``mhmocap/transforms.py:222-264``)."""
import numpy as np


def get_focal(w, theta):
    """Focal length in pixels of a ``theta``-degree field of view over ``w`` pixels (``transforms.py:262-264``)."""
    return 0.5 * w / np.tan((np.pi * theta / 180.0) / 2.0)


def compute_calibration_matrix(znear, zfar, cam_K, image_size):
    """4x4 NDC calibration PyTorch3D's ``FoVPerspectiveCameras(K=...)`` is given (``transforms.py:222-255``).

    ``image_size`` = (W, H).  The short image side spans [-1, 1]: a landscape image uses ``fy`` for BOTH axes, a
    portrait one ``fx``, a square one their average; the principal-point offset of the long axis is stretched by
    the aspect ratio.
    """
    W, H = image_size
    fx, fy, cx, cy = cam_K[0, 0], cam_K[1, 1], cam_K[0, 2], cam_K[1, 2]
    if W > H:
        s = 2 * fy / H
        ox = (W / H) * (W - 2 * cx) / W
        oy = (H - 2 * cy) / H
    elif H > W:
        s = 2 * fx / W
        ox = (W - 2 * cx) / W
        oy = (H / W) * (H - 2 * cy) / H
    else:
        s = 2 * (fx + fy) / (W + H)
        ox = (W - 2 * cx) / W
        oy = (H - 2 * cy) / H
    a = zfar / (zfar - znear)
    b = -(zfar * znear) / (zfar - znear)
    return np.array([[s, 0, ox, 0], [0, s, oy, 0], [0, 0, a, b], [0, 0, 1, 0]], dtype=np.float32)

"""Synthetic inputs for the oracle-side tests: re-exports the pure-numpy generators of ``tools/synthdata.py`` and adds
``make_sequence``, which renders the ground-truth instance masks / depth with the ORACLE's SMPL forward and rasteriser
(CPU, small sizes only).  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
from synthdata import *                                     # noqa: F401,F403
from synthdata import (V, F_, NJ, PARENTS, make_smpl_model, write_model_dir, load_model_tensors, make_motion, camera_for,
                       scene_cloud, scene_depth_plane, assemble_inputs)     # noqa: F401


def make_sequence(model_dir, N, T, W, H, seed=1):
    """Full CPU generation (small sizes only): GT motion -> oracle SMPL forward ->
    oracle rasteriser (blur 0, K=1) -> ``assemble_inputs``.  Returns (inputs, cam_K, motion)."""
    import torch
    from . import refmath, raster
    mt = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v)
          for k, v in load_model_tensors(model_dir).items()}
    mt['parents'] = [int(p) for p in mt['parents']]
    faces = torch.from_numpy(mt['faces'].astype(np.int64))
    motion = make_motion(N, T, seed)
    cam_K = camera_for(W, H)
    with torch.no_grad():
        out = refmath.smpl_forward(mt, torch.from_numpy(np.tile(motion['beta'], (T, 1, 1)).reshape(-1, 10)),
                                   torch.from_numpy(motion['theta'].reshape(-1, 72)))
        verts = out['verts'].view(T, N, V, 3) + torch.from_numpy(motion['trans']).view(T, N, 1, 3)
        j17 = refmath.regress_joints(mt['J_regressor_alphapose'], verts.view(T * N, V, 3))
        Kt = torch.from_numpy(cam_K)[None].expand(T * N, 3, 3)
        joints2d = refmath.camera_projection(j17, Kt).view(T, N, 17, 2).numpy()
        Kndc = torch.from_numpy(refmath.compute_calibration_matrix(1.0, 100.0, cam_K, (W, H)))
        zb = np.zeros((T, N, H, W), np.float32)
        for t in range(T):
            for n in range(N):
                vn = raster.world_to_ndc(verts[t, n], Kndc)
                zb[t, n] = raster.rasterize(vn, faces, H, W, 0.0, 1)['zbuf'][..., 0].numpy()
    inputs = assemble_inputs(zb, motion, joints2d, cam_K, W, H, seed)
    return inputs, cam_K, motion

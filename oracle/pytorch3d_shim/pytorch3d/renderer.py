"""``pytorch3d.renderer`` stand-in (see package docstring)."""
import torch

from oracle import raster as _raster


class FoVPerspectiveCameras(object):
    def __init__(self, R=None, T=None, K=None, device=None, **kwargs):
        self.R, self.T, self.K = R, T, K

    def transform(self, verts):
        # row-vector convention: X_view = X_world R + T ; clip = [X_view, 1] K^T ; NDC xy = clip.xy / clip.w ; z = view z
        R, T, K = self.R[0], self.T[0], self.K[0]
        vv = verts @ R + T
        x = (K[0, 0] * vv[..., 0] + K[0, 1] * vv[..., 1] + K[0, 2] * vv[..., 2] + K[0, 3])
        y = (K[1, 0] * vv[..., 0] + K[1, 1] * vv[..., 1] + K[1, 2] * vv[..., 2] + K[1, 3])
        w = (K[3, 0] * vv[..., 0] + K[3, 1] * vv[..., 1] + K[3, 2] * vv[..., 2] + K[3, 3])
        return torch.stack([x / w, y / w, vv[..., 2]], dim=-1)


class RasterizationSettings(object):
    def __init__(self, image_size=256, blur_radius=0.0, faces_per_pixel=1,
                 perspective_correct=None, **kwargs):
        self.image_size = image_size
        self.blur_radius = blur_radius
        self.faces_per_pixel = faces_per_pixel
        self.perspective_correct = perspective_correct


class Fragments(object):
    def __init__(self, pix_to_face, zbuf, bary_coords, dists):
        self.pix_to_face, self.zbuf, self.bary_coords, self.dists = pix_to_face, zbuf, bary_coords, dists


class MeshRasterizer(object):
    def __init__(self, cameras=None, raster_settings=None):
        self.cameras = cameras
        self.raster_settings = raster_settings

    def __call__(self, meshes, **kwargs):
        s = self.raster_settings
        H, W = s.image_size
        assert not s.perspective_correct
        verts = meshes.verts_padded()
        faces = meshes.faces_padded()
        vn = self.cameras.transform(verts)
        outs = [_raster.rasterize(vn[i], faces[i].long(), H, W, s.blur_radius, s.faces_per_pixel)
                for i in range(vn.shape[0])]
        st = lambda k: torch.stack([o[k] for o in outs], dim=0)
        return Fragments(st('pix_to_face'), st('zbuf'), st('bary'), st('dists'))


class SoftSilhouetteShader(object):
    def __call__(self, fragments, meshes, **kwargs):
        alpha = _raster.silhouette_alpha(fragments.dists, fragments.pix_to_face)
        img = torch.ones(alpha.shape + (4,), dtype=alpha.dtype)
        return torch.cat([img[..., :3], alpha.unsqueeze(-1)], dim=-1)


class MeshRenderer(object):
    def __init__(self, rasterizer=None, shader=None):
        self.rasterizer = rasterizer
        self.shader = shader

    def __call__(self, meshes, **kwargs):
        return self.shader(self.rasterizer(meshes), meshes)

"""``pytorch3d.structures`` stand-in (see package docstring)."""


class Meshes(object):
    def __init__(self, verts, faces):
        # verts (Nm,V,3) float32, faces (Nm,F,3) int (optimizer.py:427-428)
        self._verts = verts
        self._faces = faces

    def verts_padded(self):
        return self._verts

    def faces_padded(self):
        return self._faces

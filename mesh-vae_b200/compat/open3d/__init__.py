"""`o3d.io.read_triangle_mesh(path)` -> object with `.vertices` / `.triangles` (model.py:36) without open3d."""
import types

from meshvae_b200 import formats as _formats


class _TriangleMesh:
    def __init__(self, v, f):
        self.vertices, self.triangles = v, f


def _read_triangle_mesh(path):
    v, f = _formats.load_obj(path)
    return _TriangleMesh(v, f)


io = types.SimpleNamespace(read_triangle_mesh=_read_triangle_mesh)

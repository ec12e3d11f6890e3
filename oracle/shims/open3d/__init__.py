"""TEST-ONLY shim for open3d: only `o3d.io.read_triangle_mesh` (model.py:36) is given a body."""
import types
import numpy as np


class _TriMesh:
    def __init__(self, v, f):
        self.vertices = v
        self.triangles = f


def _read_triangle_mesh(path):
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                vs.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                fs.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
    return _TriMesh(np.asarray(vs, dtype=np.float64), np.asarray(fs, dtype=np.int32))


io = types.SimpleNamespace(read_triangle_mesh=_read_triangle_mesh)
geometry = types.SimpleNamespace()
utility = types.SimpleNamespace()

"""TEST-ONLY shim for psbody.mesh.Mesh (MPI-IS/mesh, absent from this image).

Provides what mesh_operations.py:5,208,273 and model.py:37 consume: `.v`, `.f`, an OBJ
reader and `compute_aabb_tree().nearest(points, True)` -> (faces[1,n], parts[1,n], points[n,3])
with part codes 0 = triangle interior, 1..3 = edge (v[n-1], v[n%3]), 4..6 = vertex n-4
(as consumed at mesh_operations.py:226-240).  Brute-force closest point on triangle
(Ericson, Real-Time Collision Detection 5.1.5) - exact enough for <= 5k x 2.5k pairs.
"""
import numpy as np


class _Tree:
    def __init__(self, v, f):
        self.v = np.asarray(v, dtype=np.float64)
        self.f = np.asarray(f, dtype=np.int64)

    def nearest(self, pts, nearest_part=False):
        pts = np.asarray(pts, dtype=np.float64)
        a = self.v[self.f[:, 0]][None]
        b = self.v[self.f[:, 1]][None]
        c = self.v[self.f[:, 2]][None]
        n = pts.shape[0]
        out_f = np.zeros(n, dtype=np.uint32)
        out_p = np.zeros(n, dtype=np.uint32)
        out_v = np.zeros((n, 3), dtype=np.float64)
        chunk = 256
        for s in range(0, n, chunk):
            p = pts[s:s + chunk, None, :]
            ab, ac = b - a, c - a
            ap = p - a
            d1 = (ab * ap).sum(-1)
            d2 = (ac * ap).sum(-1)
            bp = p - b
            d3 = (ab * bp).sum(-1)
            d4 = (ac * bp).sum(-1)
            cp = p - c
            d5 = (ab * cp).sum(-1)
            d6 = (ac * cp).sum(-1)
            vc = d1 * d4 - d3 * d2
            vb = d5 * d2 - d1 * d6
            va = d3 * d6 - d5 * d4
            m, k = d1.shape
            part = np.full((m, k), -1, dtype=np.int64)
            close = np.zeros((m, k, 3))

            def assign(mask, code, point):
                sel = mask & (part < 0)
                part[sel] = code
                close[sel] = np.broadcast_to(point, close.shape)[sel]

            assign((d1 <= 0) & (d2 <= 0), 4, a)
            assign((d3 >= 0) & (d4 <= d3), 5, b)
            with np.errstate(divide="ignore", invalid="ignore"):
                t_ab = d1 / (d1 - d3)
                assign((vc <= 0) & (d1 >= 0) & (d3 <= 0), 1, a + t_ab[..., None] * ab)
                assign((d6 >= 0) & (d5 <= d6), 6, c)
                t_ac = d2 / (d2 - d6)
                assign((vb <= 0) & (d2 >= 0) & (d6 <= 0), 3, a + t_ac[..., None] * ac)
                t_bc = (d4 - d3) / ((d4 - d3) + (d5 - d6))
                assign((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), 2, b + t_bc[..., None] * (c - b))
                denom = 1.0 / (va + vb + vc)
                v_ = vb * denom
                w_ = vc * denom
                assign(np.ones_like(part, dtype=bool), 0, a + ab * v_[..., None] + ac * w_[..., None])
            d = ((close - p) ** 2).sum(-1)
            best = d.argmin(1)
            idx = np.arange(m)
            out_f[s:s + m] = best
            out_p[s:s + m] = part[idx, best]
            out_v[s:s + m] = close[idx, best]
        if nearest_part:
            return out_f[None, :], out_p[None, :], out_v
        return out_f[None, :], out_v


class Mesh:
    def __init__(self, v=None, f=None, filename=None):
        if filename is not None:
            vs, fs = [], []
            with open(filename) as fh:
                for line in fh:
                    if line.startswith("v "):
                        vs.append([float(t) for t in line.split()[1:4]])
                    elif line.startswith("f "):
                        fs.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
            v, f = np.asarray(vs), np.asarray(fs)
        self.v = None if v is None else np.asarray(v, dtype=np.float64)
        self.f = None if f is None else np.asarray(f).astype(np.uint32)

    def compute_aabb_tree(self):
        return _Tree(self.v, self.f)

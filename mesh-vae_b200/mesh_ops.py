"""Operator set-up for the mesh pyramid: what `mesh_operations.py:13-278` computes once per
run from the template - adjacency A_i, down-sampling D_i (QSlim edge collapses that keep a vertex subset) and
up-sampling U_i (barycentric coordinates of every fine vertex on the coarse surface) - without `psbody` / `open3d`.

Host-side numpy / scipy, run once at start-up; nothing here touches the device.  The results must be the SAME
operators the reference builds, because checkpoints trained on them are only valid for them:

  * A_i and D_i are reproduced exactly (`tests/test_cpu_mesh_ops.py` against the golden file the unchanged reference
    produced, `tests/golden/make_operators.py`).  For D that means reproducing the reference's collapse ORDER: the
    same quadrics (plane of every face from the SVD of its homogeneous vertex matrix, mesh_operations.py:56-66), the
    same cost arithmetic (`p^T (Q_r + Q_c) p` evaluated as two matrix products on [4,1] columns, :116-121), the same
    priority queue semantics (:130-171: costs are only re-evaluated when an entry is popped, a popped entry whose
    cost has grown is pushed back, endpoints of queued entries are renamed in place after every collapse - without
    re-heapifying - and both endpoints of a collapsed edge inherit the summed quadric).
  * U_i is reproduced to floating-point tolerance (exactly three entries per row, explicit zeros kept, :202-250).

What is different is the cost: the reference scans the whole queue twice and sorts the whole face array after
every collapse (26 s for the 4998-vertex template); here queue entries are objects indexed per vertex, the vertex
count is maintained from per-vertex corner counts, and the closest-point queries go through a k-d tree over the
triangle centroids instead of an AABB tree walk per point.
"""
import heapq
import math
from typing import List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
from scipy.spatial import cKDTree


class Mesh:
    """`psbody.mesh.Mesh` as far as the drivers use it: `.v` float64 [N,3], `.f` uint32 [F,3], OBJ reader,
    `compute_aabb_tree().nearest(points, True)`."""

    def __init__(self, v=None, f=None, filename=None):
        if filename is not None:
            from . import formats
            v, f = formats.load_obj(filename)
        self.v = None if v is None else np.asarray(v, dtype=np.float64)
        self.f = None if f is None else np.asarray(f).astype(np.uint32)

    def compute_aabb_tree(self):
        return ClosestPointQuery(self.v, self.f)


# ---- adjacency --------------------------------------------------------------------------------------------
def get_vert_connectivity(mesh_v, mesh_f) -> sp.csc_matrix:
    """[N,N] matrix with one count per (directed) face edge in both directions (mesh_operations.py:13-30): interior
    edges of a manifold mesh get 2.0.  Canonical CSC, so `.tocoo()` lists the entries by column, then row."""
    f = np.asarray(mesh_f).astype(np.int64)
    n = len(mesh_v)
    a = np.concatenate([f[:, 0], f[:, 1], f[:, 2]])
    b = np.concatenate([f[:, 1], f[:, 2], f[:, 0]])
    m = sp.coo_matrix((np.ones(2 * len(a)), (np.concatenate([a, b]), np.concatenate([b, a]))), shape=(n, n)).tocsc()
    m.sum_duplicates()
    m.sort_indices()
    return m


def get_vertices_per_edge(mesh_v, mesh_f) -> np.ndarray:
    """[E,2] vertex pairs, each undirected edge once with first < second (mesh_operations.py:32-43)"""
    vc = get_vert_connectivity(mesh_v, mesh_f).tocoo()
    keep = vc.row < vc.col
    return np.stack([vc.row[keep], vc.col[keep]], 1)


# ---- QSlim ------------------------------------------------------------------------------------------------
def vertex_quadrics(mesh) -> np.ndarray:
    """[N,4,4]: sum over the faces around a vertex of the outer product of the face's unit plane equation
    (mesh_operations.py:46-68).  The plane is the last right-singular vector of [v_a 1; v_b 1; v_c 1]; the
    accumulation runs in face order, as the reference's loops do (np.add.at is sequential)."""
    f = np.asarray(mesh.f).astype(np.int64)
    verts = np.concatenate([mesh.v[f], np.ones((len(f), 3, 1))], axis=2)            # [F,3,4]
    eq = np.linalg.svd(verts)[2][:, -1, :]                                         # [F,4]
    # the normal's length exactly as the reference takes it (np.linalg.norm of a [3,1] column = sqrt(x . x) through
    # BLAS; a reduction along an axis may round differently in the last bit, and the collapse order hangs on it)
    length = np.array([np.linalg.norm(e[0:3].reshape(-1, 1)) for e in eq])
    eq = eq / length[:, None]
    outer = eq[:, :, None] * eq[:, None, :]                                        # [F,4,4]
    q = np.zeros((len(mesh.v), 4, 4))
    np.add.at(q, f.reshape(-1), np.repeat(outer, 3, axis=0))
    return q


class _Entry:
    """queue element ordered like the reference's `(cost, (r, c))` tuples; endpoints are renamed in place"""
    __slots__ = ("cost", "r", "c")

    def __init__(self, cost, r, c):
        self.cost, self.r, self.c = cost, r, c

    def __lt__(self, other):
        if self.cost != other.cost:
            return self.cost < other.cost
        if self.r != other.r:
            return self.r < other.r
        return self.c < other.c


def _get_sparse_transform(faces: np.ndarray, num_original_verts: int):
    """faces over the surviving vertices re-indexed 0..M-1, and the [M, N] selection matrix (mesh_operations.py:72-85)"""
    verts_left = np.unique(faces.reshape(-1))
    remap = np.zeros(int(verts_left.max()) + 1, dtype=np.int64)
    remap[verts_left] = np.arange(len(verts_left))
    new_faces = remap[faces.reshape(-1)].reshape(-1, 3)
    mtx = sp.csc_matrix((np.ones(len(verts_left)), (np.arange(len(verts_left)), verts_left)),
                        shape=(len(verts_left), num_original_verts))
    return new_faces, mtx


def qslim_decimator_transformer(mesh, factor=None, n_verts_desired=None):
    """(new_faces [F',3], D csc [M,N]) - mesh_operations.py:87-199.  Vertices are never moved: a collapse keeps one
    endpoint, so D selects rows."""
    if factor is None and n_verts_desired is None:
        raise Exception("Need either factor or n_verts_desired.")
    if n_verts_desired is None:
        n_verts_desired = math.ceil(len(mesh.v) * factor)
    n = len(mesh.v)
    qv = vertex_quadrics(mesh)
    hom = np.concatenate([mesh.v, np.ones((n, 1))], axis=1).reshape(n, 4, 1)        # homogeneous positions as columns

    def costs(r, c):
        qsum = qv[r] + qv[c]
        p1, p2 = hom[r], hom[c]
        destroy_c = p1.T.dot(qsum).dot(p1)[0, 0]          # same two products as the reference (bit-identical costs)
        destroy_r = p2.T.dot(qsum).dot(p2)[0, 0]
        return destroy_c, destroy_r, qsum

    # undirected edges in the reference's order: entries of the symmetric adjacency by column, then row, with r <= c
    adj = get_vertices_per_edge(mesh.v, mesh.f)
    sym = sp.csc_matrix((np.ones(len(adj)), (adj[:, 0], adj[:, 1])), shape=(n, n))
    sym = (sym + sym.T).tocoo()
    queue: List[_Entry] = []
    by_vertex: List[List[_Entry]] = [[] for _ in range(n)]
    for r, c in zip(sym.row.tolist(), sym.col.tolist()):
        if r > c:
            continue
        dc, dr, _ = costs(r, c)
        e = _Entry(min(dc, dr), r, c)
        heapq.heappush(queue, e)
        by_vertex[r].append(e)
        by_vertex[c].append(e)

    faces = np.asarray(mesh.f).astype(np.int64).copy()
    corners = np.bincount(faces.reshape(-1), minlength=n)         # face corners per vertex; > 0 <=> vertex still in use
    nverts_total = n
    while nverts_total > n_verts_desired:
        e = heapq.heappop(queue)
        r, c = e.r, e.c
        if r == c:
            continue
        dc, dr, qsum = costs(r, c)
        cost = min(dc, dr)
        if cost > e.cost:                    # outdated entry: back into the queue with its current cost
            e.cost = cost
            heapq.heappush(queue, e)
            continue
        if dc < dr:
            to_destroy, to_keep = c, r
        else:
            to_destroy, to_keep = r, c
        hit = faces == to_destroy
        faces[hit] = to_keep
        corners[to_keep] += corners[to_destroy]
        corners[to_destroy] = 0
        for q in by_vertex[to_destroy]:      # rename the endpoint in every queued entry (heap order is not repaired)
            if q.r == to_destroy:
                q.r = to_keep
            if q.c == to_destroy:
                q.c = to_keep
        by_vertex[to_keep].extend(by_vertex[to_destroy])
        by_vertex[to_destroy] = []
        qv[r] = qsum
        qv[c] = qsum
        touched = hit.any(axis=1)
        if touched.any():
            t = faces[touched]
            dead = (t[:, 0] == t[:, 1]) | (t[:, 1] == t[:, 2]) | (t[:, 2] == t[:, 0])
            if dead.any():
                np.subtract.at(corners, t[dead].reshape(-1), 1)
                keep = np.ones(len(faces), dtype=bool)
                keep[np.flatnonzero(touched)[dead]] = False
                faces = faces[keep]
        nverts_total = int(np.count_nonzero(corners))
    return _get_sparse_transform(faces, n)


# ---- closest point on a triangle mesh -------------------------------------------------------------------------
def _closest_on_triangles(p, a, b, c):
    """Closest points of the triangles (a,b,c) [M,3] to the points p [M,3] (Ericson, Real-Time Collision Detection
    5.1.5).  -> (points [M,3], part [M]) with the part codes `setup_deformation_transfer` expects: 0 interior,
    1..3 edge (corner part-1, corner part%3), 4..6 corner part-4."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
    bp = p - b
    d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
    cp = p - c
    d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
    vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
    part = np.full(len(p), -1, dtype=np.int64)
    close = np.zeros_like(p)

    def assign(mask, code, point):
        sel = mask & (part < 0)
        part[sel] = code
        close[sel] = point[sel]

    with np.errstate(divide="ignore", invalid="ignore"):
        assign((d1 <= 0) & (d2 <= 0), 4, a)
        assign((d3 >= 0) & (d4 <= d3), 5, b)
        assign((vc <= 0) & (d1 >= 0) & (d3 <= 0), 1, a + (d1 / (d1 - d3))[:, None] * ab)
        assign((d6 >= 0) & (d5 <= d6), 6, c)
        assign((vb <= 0) & (d2 >= 0) & (d6 <= 0), 3, a + (d2 / (d2 - d6))[:, None] * ac)
        assign((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), 2, b + ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[:, None] * (c - b))
        denom = 1.0 / (va + vb + vc)
        assign(np.ones(len(p), dtype=bool), 0, a + ab * (vb * denom)[:, None] + ac * (vc * denom)[:, None])
    return close, part


class ClosestPointQuery:
    """`Mesh.compute_aabb_tree()`: exact closest surface point per query, through a k-d tree over the triangle
    centroids.  The closest VERTEX bounds the distance to the surface (d0); a triangle can only hold a closer point if
    its centroid lies within d0 + its bounding radius, so only those are evaluated."""

    def __init__(self, v, f):
        self.v = np.asarray(v, dtype=np.float64)
        self.f = np.asarray(f).astype(np.int64)
        tri = self.v[self.f]
        self.centroid = tri.mean(1)
        self.radius = np.linalg.norm(tri - self.centroid[:, None, :], axis=2).max(1)
        self.vtree = cKDTree(self.v)
        self.ctree = cKDTree(self.centroid)

    def nearest(self, pts, nearest_part=False):
        pts = np.asarray(pts, dtype=np.float64)
        d0, _ = self.vtree.query(pts)
        cand = self.ctree.query_ball_point(pts, d0 + self.radius.max() + 1e-12, return_sorted=True)
        counts = np.fromiter((len(c) for c in cand), dtype=np.int64, count=len(cand))
        pid = np.repeat(np.arange(len(pts)), counts)
        fid = np.fromiter((j for c in cand for j in c), dtype=np.int64, count=int(counts.sum()))
        close, part = _closest_on_triangles(pts[pid], self.v[self.f[fid, 0]], self.v[self.f[fid, 1]], self.v[self.f[fid, 2]])
        d = ((close - pts[pid]) ** 2).sum(-1)
        # per query the candidate of smallest distance; ties go to the smallest face index (candidates are sorted)
        order = np.lexsort((fid, d, pid))
        first = order[np.concatenate([[0], np.cumsum(counts)[:-1]])]
        out_f, out_p, out_v = fid[first].astype(np.uint32), part[first].astype(np.uint32), close[first]
        if nearest_part:
            return out_f[None, :], out_p[None, :], out_v
        return out_f[None, :], out_v


def setup_deformation_transfer(source, target, use_normals=False) -> sp.csc_matrix:
    """U [len(target.v), len(source.v)]: every target vertex as a combination of the three corners of its closest
    source triangle (mesh_operations.py:202-250): barycentric fit inside the triangle, a two-corner fit on an edge
    (fitted to the TARGET point, as the reference does), a single 1.0 at a corner; always three stored entries."""
    nt = target.v.shape[0]
    faces, parts, points = source.compute_aabb_tree().nearest(target.v, True)
    faces, parts = faces.ravel().astype(np.int64), parts.ravel().astype(np.int64)
    sf = np.asarray(source.f).astype(np.int64)
    tri = sf[faces]                                                   # [nt,3] corner ids
    coeffs = np.zeros((nt, 3))
    for i in np.flatnonzero(parts == 0):
        coeffs[i] = np.linalg.lstsq(source.v[tri[i]].T, points[i], rcond=None)[0]
    for i in np.flatnonzero((parts > 0) & (parts <= 3)):
        j0, j1 = parts[i] - 1, parts[i] % 3
        a = np.vstack((source.v[tri[i, j0]], source.v[tri[i, j1]])).T
        t = np.linalg.lstsq(a, target.v[i], rcond=None)[0]
        coeffs[i, j0], coeffs[i, j1] = t[0], t[1]
    at_corner = np.flatnonzero(parts > 3)
    coeffs[at_corner, parts[at_corner] - 4] = 1.0
    rows = np.repeat(np.arange(nt), 3)
    return sp.csc_matrix((coeffs.reshape(-1), (rows, tri.reshape(-1))), shape=(nt, source.v.shape[0]))


def generate_transform_matrices(mesh, factors: Sequence[float]) -> Tuple[list, list, list, list]:
    """(M, A, D, U) of mesh_operations.py:253-278: meshes, adjacencies (COO), down- and up-sampling transforms (COO)"""
    M, A, D, U = [mesh], [get_vert_connectivity(mesh.v, mesh.f).tocoo()], [], []
    for factor in factors:
        ds_f, ds_d = qslim_decimator_transformer(M[-1], factor=1.0 / factor)
        D.append(ds_d.tocoo())
        new_mesh = Mesh(v=ds_d.dot(M[-1].v), f=ds_f)
        M.append(new_mesh)
        A.append(get_vert_connectivity(new_mesh.v, new_mesh.f).tocoo())
        U.append(setup_deformation_transfer(M[-1], M[-2]).tocoo())
    return M, A, D, U

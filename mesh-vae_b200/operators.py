"""Operator hand-off: the reference's uncoalesced COO operators (model.py:24-32) and the
(edge_index, norm) pairs of ChebConv_batch.norm (nn/conv.py:541-555) -> immutable CSR +
transpose-CSR on the device, built once per operator and cached.

The build is the host routine `mvb_csr_from_coo_host` (stable counting sort: entries of a row stay
in COO order, duplicates kept), so it can be tested without a GPU; the arrays are then uploaded."""
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib


def csr_from_coo_host(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray, n_rows: int, n_cols: int,
                      transpose: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """numpy in / numpy out wrapper of the C ABI host routine (CSR of P, or of P^T)."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    nnz = int(rows.shape[0])
    n_out_rows, n_out_cols = (n_cols, n_rows) if transpose else (n_rows, n_cols)
    rowptr = np.zeros(n_out_rows + 1, dtype=np.int32)
    colidx = np.zeros(max(nnz, 1), dtype=np.int32)
    v = np.zeros(max(nnz, 1), dtype=np.float32)
    rc = _lib.lib.mvb_csr_from_coo_host(n_out_rows, n_out_cols, nnz, rows.ctypes.data, cols.ctypes.data,
                                        vals.ctypes.data, 1 if transpose else 0, rowptr.ctypes.data,
                                        colidx.ctypes.data, v.ctypes.data)
    _lib.check(rc, "mvb_csr_from_coo_host")
    return rowptr, colidx[:nnz], v[:nnz]


class MeshOperator:
    """A fixed sparse operator P [n_rows, n_cols] resident on one device as CSR(P) and CSR(P^T)."""

    def __init__(self, rows, cols, vals, n_rows: int, n_cols: int, device):
        rows = np.asarray(rows)
        cols = np.asarray(cols)
        vals = np.asarray(vals)
        self.n_rows, self.n_cols, self.nnz = int(n_rows), int(n_cols), int(rows.shape[0])
        self.device = torch.device(device)
        rp, ci, v = csr_from_coo_host(rows, cols, vals, n_rows, n_cols, transpose=False)
        rpt, cit, vt = csr_from_coo_host(rows, cols, vals, n_rows, n_cols, transpose=True)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)  # noqa: E731
        self.rowptr, self.colidx, self.vals = up(rp), up(ci), up(v)
        self.rowptr_t, self.colidx_t, self.vals_t = up(rpt), up(cit), up(vt)
        # one nonzero per row with value 1 (the D matrices, mesh_operations.py:72-85): pure row selection
        self.is_selection = bool(self.nnz == n_rows and np.all(np.diff(rp) == 1) and np.all(v == 1.0))
        # square operators: rows >= n_active are empty and no entry references a column >= n_active
        # (the coarse operator applied to a finer tensor, models/cheb_VAE.py:288) -> closed-form rows
        self.n_active = self.n_rows
        if n_rows == n_cols:
            self.n_active = int(max(rows.max(), cols.max())) + 1 if self.nnz else 0

    def csr_bytes(self) -> int:
        return (self.n_rows + 1) * 4 + self.nnz * 8


_CACHE: Dict[tuple, Tuple[MeshOperator, tuple]] = {}


def _key(idx: torch.Tensor, val: Optional[torch.Tensor], n_rows: int, n_cols: int, kind: str, device):
    return (kind, idx.data_ptr(), idx._version, None if val is None else (val.data_ptr(), val._version),
            int(idx.shape[-1]), n_rows, n_cols, str(idx.device), str(device))


def from_edges(edge_index: torch.Tensor, norm: torch.Tensor, n: int, device=None) -> MeshOperator:
    """L_hat as used by propagate with flow='source_to_target' (nn/conv.py:172): out[t] sums
    norm[e] * x[edge_index[0][e]] over edges with edge_index[1][e] == t; N comes from x (size=None,
    nn/conv.py:160-169), so a coarse operator on a finer tensor simply has empty rows."""
    device = norm.device if device is None else torch.device(device)
    key = _key(edge_index, norm, n, n, "edges", device)
    hit = _CACHE.get(key)
    if hit is None:
        ei = edge_index.detach().cpu().numpy()
        op = MeshOperator(ei[1], ei[0], norm.detach().float().cpu().numpy(), n, n, device)
        _CACHE[key] = hit = (op, (edge_index, norm))      # keep the keyed storages alive
    return hit[0]


def from_sparse(mat: torch.Tensor, device=None) -> MeshOperator:
    """A torch sparse COO matrix used through _indices()/_values() exactly as given (nn/pool.py:19)."""
    idx, val = mat._indices(), mat._values()
    device = val.device if device is None else torch.device(device)
    key = _key(idx, val, int(mat.shape[0]), int(mat.shape[1]), "coo", device)
    hit = _CACHE.get(key)
    if hit is None:
        i = idx.detach().cpu().numpy()
        op = MeshOperator(i[0], i[1], val.detach().float().cpu().numpy(), int(mat.shape[0]), int(mat.shape[1]),
                          device)
        _CACHE[key] = hit = (op, (idx, val, mat))
    return hit[0]


def clear_cache():
    _CACHE.clear()

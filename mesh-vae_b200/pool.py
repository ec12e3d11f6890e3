"""Mesh down/up-sampling backed by the sm_100a SpMM kernel.

* `SurfacePool` - drop-in for nn/pool.py:13-23 (`forward(x[B,N,F], pool_mat: sparse COO[M,N])`).
* `Pool`        - drop-in for the free function models/cheb_cls.py:22-27 (`Pool(x, trans, dim=1)`).
The COO operator is used uncoalesced, exactly as given (`_indices()/_values()`); the backward pass
uses the precomputed transpose-CSR, so there are no atomics and gradients are deterministic."""
import torch

from . import functional as Fn
from . import operators


class SurfacePool(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, x, pool_mat, dtype=None):
        op = operators.from_sparse(pool_mat, x.device)
        return Fn.from_vertex_major(Fn.pool(Fn.to_vertex_major(x), op))


def Pool(x, trans, dim=1):
    if dim != 1 or x.dim() != 3:
        raise NotImplementedError("Pool is implemented for x[B,N,F] with dim=1 (models/cheb_cls.py:99)")
    op = operators.from_sparse(trans, x.device)
    return Fn.from_vertex_major(Fn.pool(Fn.to_vertex_major(x), op))

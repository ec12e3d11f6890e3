"""TEST-ONLY leaf shim for torch-scatter==2.0.9 (absent from this image).

Lets the UNCHANGED reference files import in the build container so that golden
vectors can be generated (tests/golden/make_golden.py).  Never imported by the
product package.  Semantics: torch-scatter 2.0.9 `scatter_sum` is
`zeros(...).scatter_add_(dim, broadcast(index), src)`; call sites in the reference:
nn/conv.py:363-364, nn/conv.py:551, models/cheb_cls.py:26.
"""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(src.dim() - index.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_sum = scatter_add


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_add(src, index, dim, out, dim_size)
    raise NotImplementedError("oracle shim: only reduce='add'/'sum' is used by the reference")


def gather_csr(*a, **k):  # name only (nn/conv.py:37); SparseTensor path never taken
    raise NotImplementedError


def segment_csr(*a, **k):
    raise NotImplementedError

"""torch-scatter 2.0.9 surface the reference touches (nn/conv.py:551, models/cheb_cls.py:19,26):
`scatter_add` == zeros().scatter_add_.  Only init-time callers remain once `accelerate()` has
swapped the reference's `Pool` for the native one."""
import torch


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    if dim < 0:
        dim += src.dim()
    if index.dim() == 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape).expand_as(src)
    if out is None:
        size = list(src.size())
        size[dim] = dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce not in ("sum", "add"):
        raise NotImplementedError(reduce)
    return scatter_add(src, index, dim, out, dim_size)

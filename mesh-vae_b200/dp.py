"""Data-parallel plumbing (new - the reference is single-device, main.py:194-195): batch sharding
and the ONE flat-gradient all-reduce per step.  Device-agnostic on purpose so the N>1 path is
covered by world_size-2 `gloo` tests on CPU; on B200s the backend is NCCL over NVLink/NVSwitch.

Semantics: every rank holds a contiguous slice of the global batch; the loss is the mean over the
batch (models/cheb_VAE.py:342), so the gradient of the global-batch mean loss is the average of
the per-rank gradients when the slices are equal."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous slice of the global batch owned by `rank`; the batch must divide
    evenly (unequal slices would bias the mean of per-rank means)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def ragged_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """slice of a global batch that does NOT divide evenly (the last batch of an epoch without drop_last): as even as
    it can be; a rank may own nothing"""
    return global_batch * rank // world, global_batch * (rank + 1) // world


def ragged_weight(global_batch: int, rank: int, world: int) -> float:
    """factor for a rank's local-MEAN gradient so that (sum over ranks) / world is the gradient of the mean over the
    whole global batch: local_count * world / global_batch (0 for a rank whose slice is empty)"""
    lo, hi = ragged_bounds(global_batch, rank, world)
    return (hi - lo) * world / float(global_batch)


def live_parameters(params: Sequence[torch.nn.Parameter]) -> List[torch.nn.Parameter]:
    """parameters that received a gradient (dec_lin_1 of cheb_VAE never does, quirk 7; torch's Adam
    skips such parameters, and so does the flat exchange buffer)"""
    return [p for p in params if p.grad is not None]


def flat_layout(params: Sequence[torch.Tensor], align: int = 32) -> Tuple[List[int], int]:
    """offsets (in elements, each a multiple of `align`) of the parameters in a flat buffer and the
    buffer length; 32 fp32 = 128 bytes, so every view keeps the kernels' 16-byte vector paths"""
    offsets, off = [], 0
    for p in params:
        offsets.append(off)
        off += (p.numel() + align - 1) // align * align
    return offsets, off


def flat_views(flat: torch.Tensor, params: Sequence[torch.Tensor], offsets: Sequence[int]) -> List[torch.Tensor]:
    return [flat[o:o + p.numel()].view_as(p) for p, o in zip(params, offsets)]


def pack_grads(params: Sequence[torch.nn.Parameter], views: Sequence[torch.Tensor]) -> None:
    """copy every parameter's gradient into its (aligned) view of the flat exchange buffer - one
    multi-tensor launch"""
    torch._foreach_copy_(list(views), [p.grad for p in params])


def unpack_grads(views: Sequence[torch.Tensor], params: Sequence[torch.nn.Parameter]) -> None:
    torch._foreach_copy_([p.grad for p in params], list(views))


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """in-place SUM all-reduce of the flat gradient buffer; the 1/world scaling is applied by the
    consumer (folded into mvb_adam_step on the GPU)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat

"""Data-parallel plumbing (new - the reference is single-device, main.py:194-195): batch sharding
and the ONE flat-gradient all-reduce per step.  Device-agnostic on purpose so the N>1 path is
covered by world_size-2 `gloo` tests on CPU; on B200s the backend is NCCL over NVLink/NVSwitch.

Semantics: every rank holds a contiguous slice of the global batch; the loss is the mean over the
batch (models/cheb_VAE.py:342), so the gradient of the global-batch mean loss is the average of
the per-rank gradients when the slices are equal."""
from typing import List, Optional, Sequence, Tuple

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


class PeerBuffers:
    """This rank's flat gradient buffer and signal pad in memory that every peer on the node has mapped, for the fused
    exchange + optimizer launch (`mvb_dp_begin` / `mvb_dp_reduce_adam`, csrc/mvb_dp.cu).  Two ways to get there:
    torch symmetric memory (`torch.distributed._symmetric_memory`: one rendezvous, no extra CUDA contexts) and, where
    that is not available, CUDA IPC handles exchanged through the process group (`torch.multiprocessing.reductions`).
    `backend`: "symm", "ipc" or None = try them in that order (env MVB_DP_PEER overrides).  Collective: every rank of
    `group` must construct it at the same point."""

    def __init__(self, n: int, device: torch.device, group=None, backend: Optional[str] = None):
        import ctypes
        import os
        from ._lib import lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > lib.mvb_dp_max_world():
            raise RuntimeError(f"PeerBuffers: world size {self.world} > {lib.mvb_dp_max_world()}")
        backend = os.environ.get("MVB_DP_PEER", backend)
        pad_words = int(lib.mvb_dp_pad_bytes()) // 4
        errors = []
        self.backend = None
        for b in ([backend] if backend else ["symm", "ipc"]):
            try:
                if b == "symm":
                    self._keep = self._alloc_symm(n, pad_words, device)
                elif b == "ipc":
                    self._keep = self._alloc_ipc(n, pad_words, device)
                else:
                    raise ValueError(f"unknown peer-memory backend {b!r}")
                self.backend = b
                break
            except Exception as e:  # noqa: BLE001
                errors.append(f"{b}: {type(e).__name__}: {e}")
        # every rank must have ended up on the same backend (or all fall back to NCCL together)
        flags = [None] * self.world
        dist.all_gather_object(flags, self.backend, group=self.group)
        if any(f != flags[0] for f in flags) or self.backend is None:
            raise RuntimeError("PeerBuffers: no common peer-memory backend (" + "; ".join(errors) + f"; ranks report {flags})")
        self.flat_g, self.pad, grad_ptrs, pad_ptrs = self._keep[:4]
        self.state = torch.zeros(int(lib.mvb_dp_state_bytes()) // 4, device=device, dtype=torch.int32)
        arr = ctypes.c_void_p * self.world
        self.grad_ptrs, self.pad_ptrs = arr(*grad_ptrs), arr(*pad_ptrs)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)          # every pad is zero and mapped before the first signal is sent

    def _alloc_symm(self, n, pad_words, device):
        import torch.distributed._symmetric_memory as symm_mem
        g = symm_mem.empty(n, dtype=torch.float32, device=device)
        pad = symm_mem.empty(pad_words, dtype=torch.int32, device=device)
        g.zero_()
        pad.zero_()
        hg = symm_mem.rendezvous(g, self.group)
        hp = symm_mem.rendezvous(pad, self.group)
        return g, pad, [int(p) for p in hg.buffer_ptrs], [int(p) for p in hp.buffer_ptrs], hg, hp

    def _alloc_ipc(self, n, pad_words, device):
        from torch.multiprocessing.reductions import reduce_tensor
        g = torch.zeros(n, device=device, dtype=torch.float32)
        pad = torch.zeros(pad_words, device=device, dtype=torch.int32)
        torch.cuda.synchronize(device)
        mine = (reduce_tensor(g), reduce_tensor(pad))
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=self.group)
        peers_g, peers_p = [], []
        for r, ((fg, ag), (fp, ap)) in enumerate(allh):
            peers_g.append(g if r == self.rank else fg(*ag))
            peers_p.append(pad if r == self.rank else fp(*ap))
        return g, pad, [t.data_ptr() for t in peers_g], [t.data_ptr() for t in peers_p], peers_g, peers_p

"""Batch hand-off of the drivers: the collate semantics of the PyG `DataLoader`
the reference uses (main.py:256-259, data.py:103-111) and a rank-sharded loader for data-parallel runs.

A dataset item is the 8-tuple of `MeshData.__getitem__` (data.py:103-111):
    (Data(x [N,3] f32, y, edge_index), ori_data [N,3] f64, label int, filename str,
     ori_mesh [N,3] f32, R [3,3] f32, m [1,3] f32, s [1] f32)
and a batch is what PyG's collater makes of a list of those: a `Batch` whose `.x` is the concatenation
[B*N,3] with `.num_graphs = B` (the model reshapes it back, models/cheb_VAE.py:195-200), stacked tensors, a
LongTensor of labels and a list of names.  `edge_index` is read and ignored by the model (quirk 11): it is
carried once, not replicated B times with offsets.
"""
from typing import Iterator, List, Sequence

import numpy as np
import torch

from . import dp


class MeshBatch:
    """stand-in for torch_geometric.data.Batch: `.x`, `.y`, `.edge_index`, `.num_graphs`, `.to()`"""

    def __init__(self, x, y, edge_index, num_graphs):
        self.x, self.y, self.edge_index, self.num_graphs = x, y, edge_index, num_graphs

    def to(self, device, non_blocking: bool = False):
        mv = lambda t: t.to(device, non_blocking=non_blocking) if torch.is_tensor(t) else t      # noqa: E731
        return MeshBatch(mv(self.x), mv(self.y), mv(self.edge_index), self.num_graphs)

    def pin_memory(self):
        pm = lambda t: t.pin_memory() if torch.is_tensor(t) else t      # noqa: E731
        return MeshBatch(pm(self.x), pm(self.y), pm(self.edge_index), self.num_graphs)


def _as_tensor(v):
    return v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))


def collate(items: Sequence[Sequence]) -> tuple:
    """list of MeshData items -> (batch, x_gt [B,N,3], y [B] int64, names, gt_mesh [B,N,3], R [B,3,3], m [B,1,3], s [B,1])"""
    cols = list(zip(*items))
    data = cols[0]
    xs = [_as_tensor(d.x) for d in data]
    ys = [_as_tensor(d.y) for d in data] if getattr(data[0], "y", None) is not None else None
    batch = MeshBatch(torch.cat(xs, 0), None if ys is None else torch.cat(ys, 0), getattr(data[0], "edge_index", None),
                      len(items))
    out: List = [batch]
    for col in cols[1:]:
        first = col[0]
        if isinstance(first, str):
            out.append(list(col))
        elif torch.is_tensor(first) or isinstance(first, np.ndarray):
            out.append(torch.stack([_as_tensor(c) for c in col], 0))
        else:
            out.append(torch.as_tensor(col))
    return tuple(out)


class ShardedMeshLoader:
    """Iterates GLOBAL batches of `batch_size * world` items in one (optionally shuffled, seeded per epoch)
    order on every rank and yields this rank's contiguous slice of each (dp.shard_bounds) - the union over ranks
    of what one reference DataLoader with the global batch size would deliver.  Every rank yields the same
    number of batches; with drop_last=False the final global batch is split as evenly as it can be and ranks
    left without an item skip it TOGETHER only if it is empty for all, so collectives stay matched: a rank
    with an empty final slice re-uses the first item of that global batch (weight 1 in a batch-mean gradient -
    set drop_last=True for exact global-batch equivalence).  world=1 is the plain loader."""

    def __init__(self, dataset, batch_size: int, shuffle: bool = False, seed: int = 0, rank: int = 0, world: int = 1,
                 drop_last: bool = False, pin_memory: bool = False):
        self.dataset, self.batch_size, self.shuffle, self.seed = dataset, int(batch_size), shuffle, int(seed)
        self.rank, self.world, self.drop_last, self.pin = int(rank), int(world), drop_last, pin_memory
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def global_order(self) -> np.ndarray:
        n = len(self.dataset)
        if not self.shuffle:
            return np.arange(n)
        return np.random.Generator(np.random.PCG64([self.seed, self.epoch])).permutation(n)

    def __len__(self) -> int:
        g = self.batch_size * self.world
        n = len(self.dataset)
        return n // g if self.drop_last else (n + g - 1) // g

    def global_chunk_sizes(self) -> List[int]:
        """items of every GLOBAL batch of an epoch (identical on all ranks): the step engine takes its replayed-graph
        path only for full global batches - a decision every rank must make alike, or their collectives mismatch"""
        g, n = self.batch_size * self.world, len(self.dataset)
        return [min(g, n - i * g) for i in range(len(self))]

    def index_batches(self) -> Iterator[np.ndarray]:
        order = self.global_order()
        g = self.batch_size * self.world
        for i in range(len(self)):
            chunk = order[i * g:(i + 1) * g]
            if len(chunk) % self.world == 0:
                lo, hi = dp.shard_bounds(len(chunk), self.rank, self.world)
            else:                                   # ragged final global batch: as even as it can be
                lo, hi = dp.ragged_bounds(len(chunk), self.rank, self.world)
            mine = chunk[lo:hi]
            yield mine if len(mine) else chunk[:1]

    def __iter__(self):
        for idx in self.index_batches():
            batch = collate([self.dataset[int(i)] for i in idx])
            if self.pin and torch.cuda.is_available():
                batch = tuple(b.pin_memory() if hasattr(b, "pin_memory") else b for b in batch)
            yield batch
        self.epoch += 1

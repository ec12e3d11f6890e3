"""Minimal PyG containers (data.py:14,110; main.py:17): the model reads .x / .edge_index / .num_graphs."""
import torch


class Data:
    def __init__(self, x=None, y=None, edge_index=None, num_graphs=1):
        self.x, self.y, self.edge_index, self.num_graphs = x, y, edge_index, num_graphs

    def to(self, device, **kw):
        for k in ("x", "y", "edge_index"):
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, **kw))
        return self


class Batch(Data):
    """what the loader yields for a list of `Data`: `.x` concatenated along dim 0, `.num_graphs` set"""


from torch.utils.data import Dataset  # noqa: E402,F401  (main.py:17 imports it and never uses it)


class DataLoader(torch.utils.data.DataLoader):
    """`torch_geometric.data.DataLoader(dataset, batch_size, shuffle, num_workers)` (main.py:256-259, inference.py:70):
    torch's loader with PyG's collate semantics for `MeshData` items (meshvae_b200.loader.collate)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        from meshvae_b200.loader import collate
        kwargs.pop("follow_batch", None)
        kwargs.pop("exclude_keys", None)
        kwargs.setdefault("collate_fn", collate)
        super().__init__(dataset, batch_size, shuffle, **kwargs)

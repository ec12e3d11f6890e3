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

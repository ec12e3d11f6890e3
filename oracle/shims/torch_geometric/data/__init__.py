"""TEST-ONLY minimal containers (data.py:14,110; main.py:17; cheb_VAE.py:195-200 needs .x/.edge_index/.num_graphs)."""
import torch


class Data:
    def __init__(self, x=None, y=None, edge_index=None, num_graphs=1):
        self.x, self.y, self.edge_index, self.num_graphs = x, y, edge_index, num_graphs

    def to(self, device):
        for k in ("x", "y", "edge_index"):
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Dataset(torch.utils.data.Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None):
        self.root, self.transform, self.pre_transform = root, transform, pre_transform


DataLoader = torch.utils.data.DataLoader

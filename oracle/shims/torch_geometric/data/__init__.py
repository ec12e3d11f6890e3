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


def _collate(batch):
    """torch_geometric 2.0.4 loader.dataloader.Collater: a list of Data -> one batch (x concatenated along dim 0,
    num_graphs set); tensors stacked; numbers -> tensors; strings kept as a list; tuples collated element-wise"""
    elem = batch[0]
    if isinstance(elem, Data):
        xs = [d.x for d in batch]
        ys = [d.y for d in batch]
        return Data(x=torch.cat(xs, 0), y=torch.cat(ys, 0) if torch.is_tensor(ys[0]) else None,
                    edge_index=elem.edge_index, num_graphs=len(batch))
    if torch.is_tensor(elem):
        return torch.stack(batch, 0)
    if isinstance(elem, float):
        return torch.tensor(batch, dtype=torch.float)
    if isinstance(elem, int):
        return torch.tensor(batch)
    if isinstance(elem, str):
        return batch
    if isinstance(elem, (tuple, list)):
        return [_collate(list(s)) for s in zip(*batch)]
    raise TypeError(f"DataLoader found invalid type: {type(elem)}")


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, follow_batch=None, exclude_keys=None, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=_collate, **kwargs)

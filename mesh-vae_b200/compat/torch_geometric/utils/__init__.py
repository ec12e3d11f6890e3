"""torch_geometric.utils helpers the reference calls at model construction (models/cheb_cls.py:16,72)."""
import torch


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loop = torch.arange(n, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        edge_weight = torch.cat([edge_weight, edge_weight.new_full((n,), fill_value)])
    return torch.cat([edge_index, loop], dim=1), edge_weight


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros(n, dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones(index.size(0)))

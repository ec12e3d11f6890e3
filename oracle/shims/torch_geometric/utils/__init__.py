"""TEST-ONLY restatement of torch_geometric.utils (2.0.4) helpers used at nn/conv.py:16,41,544,
models/cheb_cls.py:16,72."""
import torch
from torch_scatter import scatter_add


def maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return num_nodes
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    if edge_attr is None:
        return edge_index, None
    return edge_index, edge_attr[mask]


def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    n = maybe_num_nodes(edge_index, num_nodes)
    loop = torch.arange(0, n, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        loop_w = edge_weight.new_full((n,), fill_value)
        edge_weight = torch.cat([edge_weight, loop_w], dim=0)
    return torch.cat([edge_index, loop], dim=1), edge_weight


def degree(index, num_nodes=None, dtype=None):
    n = maybe_num_nodes(index, num_nodes)
    out = torch.zeros((n,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones((index.size(0),)))


def get_laplacian(edge_index, edge_weight=None, normalization=None, dtype=None, num_nodes=None):
    edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    n = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, row, dim=0, dim_size=n)
    if normalization is None:
        edge_index, _ = add_self_loops(edge_index, num_nodes=n)
        edge_weight = torch.cat([-edge_weight, deg], dim=0)
    elif normalization == "sym":
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float("inf"), 0)
        edge_weight = dis[row] * edge_weight * dis[col]
        edge_index, edge_weight = add_self_loops(edge_index, -edge_weight, fill_value=1.0, num_nodes=n)
    else:
        dinv = 1.0 / deg
        dinv.masked_fill_(dinv == float("inf"), 0)
        edge_weight = dinv[row] * edge_weight
        edge_index, edge_weight = add_self_loops(edge_index, -edge_weight, fill_value=1.0, num_nodes=n)
    return edge_index, edge_weight

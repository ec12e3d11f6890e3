"""TEST-ONLY restatement (from the published algorithm) of torch_geometric==2.0.4
`torch_geometric.nn.conv.cheb_conv.ChebConv`, the third-party class used by
models/cheb_cls.py:18,76,95.  Source is NOT under /root/reference -> parity for this class is
"unpinned" (DESIGN.md); it follows the vendored older form at nn/conv.py:464-521 with
node_dim=-2, K bias-free Linear layers (glorot) and a zero bias.
"""
import math
import torch
from torch.nn import Parameter
from torch_scatter import scatter_add
from torch_geometric.utils import remove_self_loops, add_self_loops, get_laplacian


class ChebConv(torch.nn.Module):
    node_dim = -2

    def __init__(self, in_channels, out_channels, K, normalization="sym", bias=True, **kwargs):
        super().__init__()
        assert K > 0
        assert normalization in [None, "sym", "rw"]
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.normalization = normalization
        self.lins = torch.nn.ModuleList(
            [torch.nn.Linear(in_channels, out_channels, bias=False) for _ in range(K)])
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        for lin in self.lins:
            a = math.sqrt(6.0 / (lin.weight.size(0) + lin.weight.size(1)))
            lin.weight.data.uniform_(-a, a)
        if self.bias is not None:
            self.bias.data.fill_(0)

    def __norm__(self, edge_index, num_nodes, edge_weight, normalization, lambda_max, dtype=None, batch=None):
        edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
        edge_index, edge_weight = get_laplacian(edge_index, edge_weight, normalization, dtype, num_nodes)
        if batch is not None and lambda_max.numel() > 1:
            lambda_max = lambda_max[batch[edge_index[0]]]
        edge_weight = (2.0 * edge_weight) / lambda_max
        edge_weight.masked_fill_(edge_weight == float("inf"), 0)
        edge_index, edge_weight = add_self_loops(edge_index, edge_weight, fill_value=-1.0, num_nodes=num_nodes)
        return edge_index, edge_weight

    def propagate(self, edge_index, x, norm):
        # flow source_to_target: messages from edge_index[0] summed at edge_index[1], along dim -2
        x_j = x.index_select(self.node_dim, edge_index[0])
        msg = norm.view(-1, 1) * x_j
        return scatter_add(msg, edge_index[1], dim=self.node_dim, dim_size=x.size(self.node_dim))

    def forward(self, x, edge_index, edge_weight=None, batch=None, lambda_max=None):
        if self.normalization != "sym" and lambda_max is None:
            raise ValueError("You need to pass `lambda_max` to `forward() in`"
                             "case the normalization is non-symmetric.")
        if lambda_max is None:
            lambda_max = torch.tensor(2.0, dtype=x.dtype, device=x.device)
        if not isinstance(lambda_max, torch.Tensor):
            lambda_max = torch.tensor(lambda_max, dtype=x.dtype, device=x.device)
        edge_index, norm = self.__norm__(edge_index, x.size(self.node_dim), edge_weight,
                                         self.normalization, lambda_max, dtype=x.dtype, batch=batch)
        Tx_0 = x
        Tx_1 = x
        out = self.lins[0](Tx_0)
        if len(self.lins) > 1:
            Tx_1 = self.propagate(edge_index, x=x, norm=norm)
            out = out + self.lins[1](Tx_1)
        for lin in self.lins[2:]:
            Tx_2 = self.propagate(edge_index, x=Tx_1, norm=norm)
            Tx_2 = 2.0 * Tx_2 - Tx_0
            out = out + lin(Tx_2)
            Tx_0, Tx_1 = Tx_1, Tx_2
        if self.bias is not None:
            out = out + self.bias
        return out

"""Host-side mirror of `cheb_GCN` / `Pool` (models/cheb_cls.py:22-27, 55-114) on the native modules:
state-dict keys `cheb.{i}.lins.{k}.weight`, `cheb.{i}.bias`, `enc_lin`, `cls_layer`."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from . import operators
from .conv import ChebConv
from .pool import Pool  # noqa: F401  (re-exported: the reference defines Pool in this module)


class cheb_GCN(nn.Module):
    def __init__(self, num_feature, config, downsample_matrices, upsample_matrices, adjacency_matrices, num_nodes):
        super().__init__()
        self.n_layers = config["n_layers"]
        # the reference inserts into config['num_conv_filters'] IN PLACE (cheb_cls.py:60-61, quirk 9);
        # callers that rely on it (crecon.py:241 re-reads the config) keep working
        self.filters = config["num_conv_filters"]
        self.filters.insert(0, num_feature)
        self.z = config["num_classes"]
        self.K = config["polygon_order"]
        self.downsample_matrices = downsample_matrices
        self.upsample_matrices = upsample_matrices
        self.adjacency_matrices = adjacency_matrices
        self.A_edge_index = []
        for i in range(len(num_nodes)):
            ei = adjacency_matrices[i]._indices()
            self.A_edge_index.append(ei[:, ei[0] != ei[1]])
        f = self.filters
        self.cheb = nn.ModuleList([ChebConv(f[i], f[i + 1], self.K[i]) for i in range(len(f) - 2)])
        for conv in self.cheb:
            conv.fuse_relu = True
        self.enc_lin = nn.Linear(downsample_matrices[-1].shape[0] * f[-2], 128)
        self.cls_layer = nn.Linear(128, self.z)
        self.reset_parameters()

    def forward(self, data):
        x = data
        b = x.shape[0]
        x = x.reshape(b, -1, self.filters[0])
        for i in range(self.n_layers):
            conv = self.cheb[i]
            if conv.fuse_relu and x.is_cuda:
                # conv + relu + Pool(D) (cheb_cls.py:95-99) as one call: a single mesh-resident launch at the
                # coarse levels, the pool / conv / pool composition elsewhere
                op = conv.mesh_operator(self.A_edge_index[i], x.shape[1], x.device, x.dtype)
                d_op = operators.from_sparse(self.downsample_matrices[i], x.device)
                x = Fn.from_vertex_major(Fn.cheb_layer(Fn.to_vertex_major(x), conv.stacked_weight(), conv.bias, op, None, d_op,
                                                       relu=True))
                continue
            x = conv(x, self.A_edge_index[i])
            if not conv.fuse_relu:      # F.relu of cheb_cls.py:97, fused into the conv epilogue otherwise
                x = F.relu(x)
            x = Pool(x, self.downsample_matrices[i])
        if x.is_cuda:
            # x.reshape(B, -1) of cheb_cls.py:101 read straight from the vertex-major buffer; both Linears native
            h = Fn.linear(Fn.to_vertex_major(x), self.enc_lin.weight, self.enc_lin.bias, relu=True, x_vm=True)
            return Fn.linear(h, self.cls_layer.weight, self.cls_layer.bias)
        x = x.reshape(b, self.enc_lin.in_features)
        return self.cls_layer(F.relu(self.enc_lin(x)))

    def reset_parameters(self):
        nn.init.normal_(self.enc_lin.weight, 0, 0.1)
        nn.init.normal_(self.cls_layer.weight, 0, 0.1)
        for i in range(self.n_layers):
            self.cheb[i].reset_parameters()

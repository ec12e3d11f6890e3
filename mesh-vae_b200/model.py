"""Model selection and operator hand-off (`model.py:24-69` of the reference): template -> mesh pyramid -> uncoalesced
COO torch tensors (int64 indices, fp32 values, quirk 12) -> `cheb_VAE` / `cheb_GCN` on the device.  No `psbody`, no
`open3d`: the template is read by `formats.load_obj`, the operators come from `mesh_ops` (row f1)."""
import os

import numpy as np
import torch

from . import formats, mesh_ops
from .cheb_cls import cheb_GCN
from .cheb_vae import cheb_VAE


def scipy_to_torch_sparse(scp_matrix) -> torch.Tensor:
    """scipy COO -> torch sparse COO exactly as handed over by model.py:24-32: entries in the given order, NOT
    coalesced (U keeps its explicit zeros), LongTensor indices, FloatTensor values."""
    m = scp_matrix.tocoo() if not hasattr(scp_matrix, "row") else scp_matrix
    i = torch.LongTensor(np.vstack((m.row, m.col)))
    v = torch.FloatTensor(m.data)
    return torch.sparse_coo_tensor(i, v, torch.Size(m.shape), check_invariants=False)


def build_operators(template_path_or_mesh, downsampling_factors, device="cpu"):
    """-> (M, A_t, D_t, U_t, num_nodes): the meshes and the torch operators of model.py:36-46"""
    mesh = template_path_or_mesh
    if isinstance(mesh, (str, os.PathLike)):
        v, f = formats.load_obj(mesh)
        mesh = mesh_ops.Mesh(v=v, f=f)
    M, A, D, U = mesh_ops.generate_transform_matrices(mesh, downsampling_factors)
    to = lambda mats: [scipy_to_torch_sparse(m).to(device) for m in mats]      # noqa: E731
    return M, to(A), to(D), to(U), [len(m.v) for m in M]


def get_model(config, device, model_type=None, save_init=True):
    """model.py:35-69: `config['type']` (or `model_type`) selects `cheb_VAE` (3 input features) or `cheb_GCN`
    (6: the reference feeds it vertex + displacement channels); `initial_weight.pt` is written to the checkpoint
    directory unless save_init=False."""
    M, A_t, D_t, U_t, num_nodes = build_operators(config["template"], config["downsampling_factors"], device)
    num_feature = M[0].v.shape[1]
    if model_type is None:
        model_type = config["type"]
    if model_type == "cheb_VAE":
        print("Using model: cheb_VAE")
        net = cheb_VAE(num_feature, config, D_t, U_t, A_t, num_nodes, model=config["model"]).to(device)
    elif model_type == "cheb_GCN":
        print("Using model: cheb_GCN")
        net = cheb_GCN(num_feature * 2, config, D_t, U_t, A_t, num_nodes).to(device)
    else:
        raise ValueError(f"unknown model type {model_type!r} (cheb_VAE | cheb_GCN)")
    for name, p in net.named_parameters():
        print(name, ":", p.size())
    if save_init:
        os.makedirs(config["checkpoint_dir"], exist_ok=True)
        formats.save_initial_weight(net, config["checkpoint_dir"])
    return net

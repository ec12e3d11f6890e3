"""Delivery mode 1 (import-path drop-in): the UNCHANGED reference model files import and construct on
top of `compat/`.  Needs the reference tree (/root/reference in the build container, or its staged unmodified copy
oracle/_ref/reference); forward + backward of these same files on a B200: tests/test_gpu_dropin.py."""
import os
import subprocess
import sys
import pytest

from tests.helpers import ROOT
from oracle.ref_loader import reference_root

REF = reference_root() or "/nonexistent"       # /root/reference (build container) or the staged oracle/_ref/reference

SCRIPT = r"""
import sys, copy
sys.path.insert(0, {root!r})
import meshvae_b200 as mvb
compat = mvb.install_compat()
sys.path.insert(1, {ref!r})
sys.path.insert(2, {shims!r})          # only open3d / psbody leaves (utils.py:4) come from the test shims
import torch
from models.cheb_VAE import cheb_VAE            # the reference's own file, unchanged
from models.cheb_cls import cheb_GCN
import nn.conv, nn.pool, logpdf
assert nn.conv.ChebConv_batch is mvb.ChebConv_batch and nn.pool.SurfacePool is mvb.SurfacePool
from oracle.mesh_vae_oracle import load_operators, DEFAULT_CONFIG
A, D, U, nn_ = load_operators({npz!r})
net = cheb_VAE(3, copy.deepcopy(DEFAULT_CONFIG), D, U, A, nn_, model="optimal_sigma_VAE")
assert isinstance(net.cheb[0], mvb.ChebConv_batch) and isinstance(net.pool, mvb.SurfacePool)
assert sum(p.numel() for p in net.parameters()) == 712642
mvb.accelerate(net)
assert net.cheb[0].fuse_relu and not net.cheb_dec[-1].fuse_relu
gcn = cheb_GCN(6, copy.deepcopy(DEFAULT_CONFIG), D, U, A, nn_)
assert isinstance(gcn.cheb[0], mvb.ChebConv)
print("COMPAT_OK")
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_unchanged_reference_models_construct_on_native_modules():
    code = SCRIPT.format(root=ROOT, ref=REF, shims=os.path.join(ROOT, "oracle", "shims"),
                         npz=os.path.join(ROOT, "tests", "golden", "operators_template5k.npz"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "COMPAT_OK" in out.stdout, out.stdout + out.stderr


SCRIPT_F1 = r"""
import sys, os, tempfile
sys.path.insert(0, {root!r})
import meshvae_b200 as mvb
mvb.install_compat()
sys.path.insert(1, {ref!r})             # NO test shims: psbody / open3d / mesh_operations resolve inside compat/
import numpy as np, torch
import psbody.mesh, open3d, mesh_operations
assert psbody.mesh.Mesh is mvb.mesh_ops.Mesh and mesh_operations.generate_transform_matrices is mvb.mesh_ops.generate_transform_matrices
from model import get_model                      # the reference's own model.py (model.py:35-69), unchanged
from config_parser import read_config            # and its own parser
cfg = read_config(os.path.join({ref!r}, "files", "default.cfg"))
cfg["template"] = os.path.join({ref!r}, "template", "template5k.obj")
cfg["checkpoint_dir"] = tempfile.mkdtemp()
net = get_model(cfg, "cpu")
assert type(net).__module__ == "models.cheb_VAE" and sum(p.numel() for p in net.parameters()) == 712642
assert os.path.exists(os.path.join(cfg["checkpoint_dir"], "initial_weight.pt"))
g = np.load({npz!r})
for i in range(4):                               # the operators the reference model holds are the golden ones
    d, u = net.downsample_matrices[i], net.upsample_matrices[i]
    assert np.array_equal(d._indices()[1].numpy(), g[f"D{{i}}_col"]) and np.array_equal(u._indices()[1].numpy(), g[f"U{{i}}_col"])
    assert np.abs(u._values().numpy() - g[f"U{{i}}_val"]).max() <= 1e-6
assert mvb.formats.read_config(os.path.join({ref!r}, "files", "default.cfg")) == read_config(os.path.join({ref!r}, "files", "default.cfg"))
print("COMPAT_F1_OK")
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_get_model_runs_without_psbody_or_open3d():
    code = SCRIPT_F1.format(root=ROOT, ref=REF, npz=os.path.join(ROOT, "tests", "golden", "operators_template5k.npz"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "COMPAT_F1_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


SCRIPT_DRIVERS = r"""
import sys
sys.path.insert(0, {root!r})
import meshvae_b200 as mvb
mvb.install_compat()
sys.path.insert(1, {ref!r})
sys.path.insert(2, {plot!r})            # matplotlib only (absent from the image; the drivers never call it here)
import torch
import main, inference, crecon           # the reference's drivers, unchanged: every import resolves on compat/
from torch_geometric.data import DataLoader, Data
assert main.get_model.__module__ == "model" and main.Mesh is mvb.mesh_ops.Mesh
class DS(torch.utils.data.Dataset):
    def __len__(self): return 5
    def __getitem__(self, i):
        x = torch.full((7, 3), float(i))
        return Data(x=x, y=x, edge_index=torch.zeros(2, 3, dtype=torch.long)), x.double(), i % 2, "f%d.obj" % i, x, torch.eye(3), torch.zeros(1, 3), torch.ones(1)
batches = list(DataLoader(DS(), batch_size=2, shuffle=False, num_workers=0))
assert len(batches) == 3
x, x_gt, y, f, gt, R, m, s = batches[0]
assert x.num_graphs == 2 and x.x.shape == (14, 3) and x_gt.shape == (2, 7, 3) and y.tolist() == [0, 1] and f == ["f0.obj", "f1.obj"]
assert R.shape == (2, 3, 3) and m.shape == (2, 1, 3) and s.shape == (2, 1) and batches[2][0].num_graphs == 1
print("COMPAT_DRIVERS_OK")
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_drivers_import_on_compat():
    code = SCRIPT_DRIVERS.format(root=ROOT, ref=REF, plot=os.path.join(ROOT, "oracle", "shims_plot"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "COMPAT_DRIVERS_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]

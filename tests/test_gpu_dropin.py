"""Delivery mode 1 on a B200 (north star: "models/cheb_VAE.py and models/cheb_cls.py run unchanged"):
the reference's OWN model files - imported from the staged, unmodified copy oracle/_ref/reference (oracle/make_ref.py;
/root/reference in the build container) with `compat/` first on sys.path - run forward + backward on the GPU through
the native ChebConv_batch / SurfacePool / ChebConv / logpdf, and are checked against the golden vectors the same files
produced on the CPU through the leaf shims (tests/golden/make_golden.py).  Both as constructed and after
`accelerate()` (fused ReLU, fused loss epilogue).  fp32 tolerance 1e-4 (max-norm per tensor) + the per-element gate
of tests/helpers.elem_err on the outputs.  Runs in a subprocess: the reference's top-level module names (`nn`,
`models`, `utils`, `main` ...) must not leak into the test session."""
import json
import os
import subprocess
import sys

import pytest

from tests.helpers import ROOT
from oracle.ref_loader import reference_root

pytestmark = pytest.mark.gpu
REF = reference_root()

SCRIPT = r"""
import sys, copy, json, os
sys.path.insert(0, {root!r})
import numpy as np, torch
import meshvae_b200 as mvb
mvb.install_compat()
sys.path.insert(1, {ref!r})
sys.path.insert(2, {shims!r})          # only the open3d / psbody leaves of utils.py:4 come from the test shims
from models.cheb_VAE import cheb_VAE            # the reference's own files, unchanged
from models.cheb_cls import cheb_GCN
import nn.conv, nn.pool
assert nn.conv.ChebConv_batch is mvb.ChebConv_batch and nn.pool.SurfacePool is mvb.SurfacePool
assert os.path.realpath(sys.modules["models.cheb_VAE"].__file__).startswith(os.path.realpath({ref!r}))
from torch_geometric.data import Data
from tests.helpers import seeded_state_dict, seeded_batch, rel_err, elem_err, GOLDEN, OPERATORS_NPZ
from oracle.mesh_vae_oracle import load_operators, DEFAULT_CONFIG
dev = torch.device("cuda:0")
A, D, U, nn_ = load_operators(OPERATORS_NPZ)
Ad, Dd, Ud = [a.to(dev) for a in A], [d.to(dev) for d in D], [u.to(dev) for u in U]
T = torch.from_numpy
res = {{}}
gv = np.load(os.path.join(GOLDEN, "golden_vae.npz"))
cfg = copy.deepcopy(DEFAULT_CONFIG); cfg["dropout"] = 0.0
net = cheb_VAE(3, cfg, Dd, Ud, Ad, nn_, model=cfg["model"])
assert type(net).__module__ == "models.cheb_VAE"
net.load_state_dict(seeded_state_dict(net, 7))
net = net.to(dev)
x, y, _ = seeded_batch(2, nn_[0], 11)
y_hot = torch.nn.functional.one_hot(y, 2).to(dev)
c0 = mvb._lib.lib.mvb_launch_count()
for accel in (False, True):
    if accel:
        mvb.accelerate(net)
    for m_type, x_gt in (("train", x.double()), ("test", x.clone())):
        net.zero_grad()
        net.train() if m_type == "train" else net.eval()
        data = Data(x=x.reshape(-1, 3).clone().to(dev), edge_index=None, num_graphs=2)
        torch.manual_seed(1234)        # the reference draws eps on the global CPU generator (cheb_VAE.py:316)
        loss, correct, recon, (kld, rec, z_), y_hat = net(data, x_gt.to(dev), y_hot, m_type=m_type)
        tag = ("accel_" if accel else "plain_") + m_type
        errs = {{"loss": rel_err(loss, T(gv[m_type + "_loss"])), "recon": rel_err(recon, T(gv[m_type + "_recon"])),
                "kld": rel_err(kld, T(gv[m_type + "_kld"])), "rec": rel_err(rec, T(gv[m_type + "_rec"])),
                "z": rel_err(z_, T(gv[m_type + "_z"])), "yhat": rel_err(y_hat, T(gv[m_type + "_yhat"])),
                "recon_elem": elem_err(recon, T(gv[m_type + "_recon"])) * 1e-4}}
        assert int(correct) == int(gv[m_type + "_correct"])
        assert str(loss.dtype) == ("torch.float64" if m_type == "train" else "torch.float32")
        if m_type == "train":
            loss.backward()
            for name, p in net.named_parameters():
                if "grad_none__" + name in gv.files:
                    assert p.grad is None, name
                elif "grad__" + name in gv.files:
                    errs["grad " + name] = rel_err(p.grad, T(gv["grad__" + name]))
                else:
                    errs["grad " + name] = rel_err(p.grad.flatten()[::97], T(gv["gradsample__" + name]))
        res[tag] = errs
    net.eval()
    so = net.sample((1 - y_hot).float(), T(gv["test_z"]).to(dev))
    res[("accel_" if accel else "plain_") + "sample"] = {{"sample_oppo": rel_err(so, T(gv["sample_oppo"]))}}
res["vae_launches"] = int(mvb._lib.lib.mvb_launch_count() - c0)
# ---- cheb_GCN (models/cheb_cls.py, PyG ChebConv signature) ----
gg = np.load(os.path.join(GOLDEN, "golden_gcn.npz"))
gcn = cheb_GCN(6, copy.deepcopy(DEFAULT_CONFIG), Dd, Ud, Ad, nn_)
assert type(gcn).__module__ == "models.cheb_cls" and isinstance(gcn.cheb[0], mvb.ChebConv)
gcn.load_state_dict(seeded_state_dict(gcn, 9))
gcn = gcn.to(dev)
xg, yg, _ = seeded_batch(2, nn_[0], 13, feats=6)
for accel in (False, True):
    if accel:
        mvb.accelerate(gcn)
    gcn.zero_grad()
    logits = gcn(xg.to(dev))
    l = torch.nn.functional.cross_entropy(logits, yg.to(dev))
    l.backward()
    errs = {{"logits": rel_err(logits, T(gg["logits"])), "loss": rel_err(l, T(gg["loss"]))}}
    for name, p in gcn.named_parameters():
        if "grad__" + name in gg.files:
            errs["grad " + name] = rel_err(p.grad, T(gg["grad__" + name]))
    res[("accel_" if accel else "plain_") + "gcn"] = errs
print("DROPIN_JSON " + json.dumps(res))
"""


@pytest.mark.skipif(REF is None, reason="no reference tree (neither /root/reference nor oracle/_ref/reference; run oracle/make_ref.py)")
def test_unchanged_reference_models_run_on_the_gpu_through_compat():
    code = SCRIPT.format(root=ROOT, ref=REF, shims=os.path.join(ROOT, "oracle", "shims"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("DROPIN_JSON ")]
    assert line, out.stdout[-3000:] + out.stderr[-6000:]
    res = json.loads(line[0][len("DROPIN_JSON "):])
    assert res.pop("vae_launches") > 200, "the native kernels did not run"
    worst = {tag: max(errs.items(), key=lambda kv: kv[1]) for tag, errs in res.items()}
    log = os.path.join(ROOT, "gpurun_out", "dropin_reference_gpu.json")
    try:
        os.makedirs(os.path.dirname(log), exist_ok=True)
        with open(log, "w") as f:
            json.dump({"reference_root": REF, "errors": res}, f, indent=1)
    except OSError:
        pass
    for tag, (name, err) in worst.items():
        assert err < 1e-4, f"{tag}: {name} rel err {err:.2e}"

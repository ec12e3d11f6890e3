"""Generate tests/golden/golden_*.npz from the UNCHANGED reference, imported from /root/reference
through the TEST-ONLY leaf shims in oracle/shims.

Build container only:   python tests/golden/make_golden.py
What is run (reference file:line):
  * nn/conv.py:532-581  ChebConv_batch.norm / forward (+ autograd backward)   -> golden_ops.npz
  * nn/pool.py:13-23    SurfacePool.forward            (+ backward)            -> golden_ops.npz
  * models/cheb_cls.py:22-27  Pool                                              -> golden_ops.npz
  * logpdf.py:7-8,22-28 KLD / gaussian_nll / softclip                           -> golden_ops.npz
  * models/cheb_VAE.py:104-351 full forward (train + test m_type), loss, backward -> golden_vae.npz
  * models/cheb_cls.py:55-114 cheb_GCN forward/backward (through the restated PyG ChebConv shim,
    i.e. that part is NOT an independent pin)                                   -> golden_gcn.npz
Inputs and parameters come from tests/helpers.py (seeded, construction-order independent), so the
tests regenerate them instead of storing them.
"""
import copy
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, ROOT)

from tests.helpers import seeded_state_dict, seeded_batch, OPERATORS_NPZ  # noqa: E402
from oracle.mesh_vae_oracle import load_operators, DEFAULT_CONFIG  # noqa: E402  (fixture loader only)

from nn.conv import ChebConv_batch  # noqa: E402   reference
from nn.pool import SurfacePool  # noqa: E402      reference
import logpdf  # noqa: E402                        reference
from models.cheb_VAE import cheb_VAE  # noqa: E402 reference
from models.cheb_cls import cheb_GCN, Pool  # noqa: E402 reference
from torch_geometric.data import Data  # noqa: E402 shim

torch.set_num_threads(1)


def npify(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def golden_ops(A, D, U, num_nodes):
    out = {}
    # ---- ChebConv_batch on levels 2 (313) and 3 (79), plus the "coarse operator on a bigger tensor" quirk
    cases = [("l2_16_16", 2, 313, 3, 16, 16, True), ("l3_16_32", 3, 79, 2, 16, 32, True),
             ("l3_32_16", 3, 79, 4, 32, 16, True), ("l2_3_16", 2, 313, 4, 3, 16, True),
             ("quirk_l4_on_313_16_3", 4, 313, 2, 16, 3, False), ("l3_6_16_K3", 3, 79, 2, 6, 16, True)]
    for ci, (name, lvl, n, b, fin, fout, has_bias) in enumerate(cases):
        K = 3 if name.endswith("K3") else 6
        conv = ChebConv_batch(fin, fout, K)
        if not has_bias:
            conv.bias = None
        conv.load_state_dict(seeded_state_dict(conv, 100 + ci))
        ei, norm = ChebConv_batch.norm(A[lvl]._indices(), num_nodes[lvl])
        g = torch.Generator().manual_seed(200 + ci)
        x = torch.randn(b, n, fin, generator=g, requires_grad=True)
        dy = torch.randn(b, n, fout, generator=g)
        y = conv(x, ei, norm)
        y.backward(dy)
        out[f"conv_{name}_y"] = y
        out[f"conv_{name}_dx"] = x.grad
        out[f"conv_{name}_dw"] = conv.weight.grad
        if has_bias:
            out[f"conv_{name}_db"] = conv.bias.grad
        out[f"norm_l{lvl}"] = norm
        out[f"norm_l{lvl}_ei"] = ei
    # ---- SurfacePool / Pool
    pool = SurfacePool()
    for pi, (name, mat, f) in enumerate([("D2", D[2], 16), ("D3", D[3], 32), ("U3", U[3], 32), ("U2", U[2], 32)]):
        g = torch.Generator().manual_seed(300 + pi)
        x = torch.randn(3, mat.shape[1], f, generator=g, requires_grad=True)
        dy = torch.randn(3, mat.shape[0], f, generator=g)
        y = pool(x, mat)
        y.backward(dy)
        out[f"pool_{name}_y"] = y
        out[f"pool_{name}_dx"] = x.grad
        out[f"pool1_{name}_y"] = Pool(x.detach(), mat)
    # ---- logpdf
    g = torch.Generator().manual_seed(400)
    mu = torch.randn(5, 16, generator=g)
    logvar = torch.randn(5, 16, generator=g)
    out["kld"] = logpdf.KLD(mu, logvar)
    out["softclip_1_m6"] = logpdf.softclip(torch.Tensor([1]), -6)
    xs = torch.randn(5, 79, 3, generator=g).double()
    rec = torch.randn(5, 79, 3, generator=g)
    out["nll_f64"] = logpdf.gaussian_nll(rec, out["softclip_1_m6"], xs)
    out["nll_f32"] = logpdf.gaussian_nll(rec, out["softclip_1_m6"], xs.float())
    return npify(out)


def golden_vae(A, D, U, num_nodes):
    out = {}
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["dropout"] = 0.0          # masks cannot match across implementations; parity runs have dropout off
    net = cheb_VAE(3, cfg, D, U, A, num_nodes, model=cfg["model"])
    net.load_state_dict(seeded_state_dict(net, 7))
    B = 2
    x, y, eps = seeded_batch(B, num_nodes[0], 11)
    y_hot = torch.nn.functional.one_hot(y, 2)
    for m_type, x_gt in (("train", x.double()), ("test", x.clone())):
        net.zero_grad()
        net.train() if m_type == "train" else net.eval()
        data = Data(x=x.reshape(-1, 3).clone(), edge_index=None, num_graphs=B)
        # reference draws eps from the global CPU generator (cheb_VAE.py:316); reproduce `eps` exactly
        torch.manual_seed(1234)
        eps_ref = torch.normal(mean=0, std=1, size=(B, 16))
        torch.manual_seed(1234)
        loss, correct, recon, (kld, rec, z_), y_hat = net(data, x_gt, y_hot, m_type=m_type)
        t = m_type
        out[f"{t}_loss"] = loss
        out[f"{t}_correct"] = correct
        out[f"{t}_recon"] = recon
        out[f"{t}_kld"] = kld
        out[f"{t}_rec"] = rec
        out[f"{t}_z"] = z_
        out[f"{t}_yhat"] = y_hat
        out[f"{t}_eps"] = eps_ref
        if m_type == "train":
            loss.backward()
            for name, p in net.named_parameters():
                if p.grad is None:
                    out[f"grad_none__{name}"] = np.zeros(0)
                    continue
                gq = p.grad
                if gq.numel() > 20000:      # big Linear weights: checksums + a strided sample
                    out[f"gradsum__{name}"] = torch.stack([gq.double().sum(), gq.double().abs().sum()])
                    out[f"gradsample__{name}"] = gq.flatten()[::97].clone()
                else:
                    out[f"grad__{name}"] = gq
    # sample() with swapped sex (main.py:152, inference.py:114)
    net.eval()
    out["sample_oppo"] = net.sample((1 - y_hot).float(), torch.from_numpy(out["test_z"] if not torch.is_tensor(out["test_z"]) else out["test_z"].detach().numpy()))
    return npify(out)


def golden_gcn(A, D, U, num_nodes):
    out = {}
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    net = cheb_GCN(6, cfg, D, U, A, num_nodes)     # mutates cfg['num_conv_filters'] (quirk 9)
    net.load_state_dict(seeded_state_dict(net, 9))
    B = 2
    x, y, _ = seeded_batch(B, num_nodes[0], 13, feats=6)
    logits = net(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    out["logits"] = logits
    out["loss"] = loss
    for name, p in net.named_parameters():
        if p.numel() <= 20000:
            out[f"grad__{name}"] = p.grad
        else:
            out[f"gradsample__{name}"] = p.grad.flatten()[::97].clone()
    return npify(out)


def main():
    A, D, U, num_nodes = load_operators(OPERATORS_NPZ)
    np.savez_compressed(os.path.join(HERE, "golden_ops.npz"), **golden_ops(A, D, U, num_nodes))
    np.savez_compressed(os.path.join(HERE, "golden_vae.npz"), **golden_vae(A, D, U, num_nodes))
    np.savez_compressed(os.path.join(HERE, "golden_gcn.npz"), **golden_gcn(A, D, U, num_nodes))
    for f in ("golden_ops.npz", "golden_vae.npz", "golden_gcn.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

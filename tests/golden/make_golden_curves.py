"""North-star loss-curve gate (BASELINE.json: "reconstruction error and loss curves within 1% over 100 steps on synthetic
hip-bone-shaped meshes"): the curves of the UNCHANGED reference - models/cheb_VAE.py through the leaf shims, torch Adam
(lr 1e-3, weight_decay 5e-4, main.py:251), batches of 16 (files/default.cfg:26) of tests/synthetic.HipLikeDataset
prepared with the reference's own utils.procrustes (utils.py:58-156) and z-score (data.py:166-184), reconstruction error
as main.py:88-93 computes it - for
  * exact : dropout 0, noise from torch.manual_seed(777) on the global CPU generator (cheb_VAE.py:316) - reproducible
            step by step by any implementation that draws the same noise;
  * drop_a / drop_b : dropout 0.2 (files/default.cfg:32) under two seeds - their difference is the reference's own
            seed-to-seed spread, the yardstick of the statistical comparison.
Build container only:   python tests/golden/make_golden_curves.py      -> tests/golden/golden_curves.npz
"""
import copy
import os
import sys
import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_loader import use_reference_on_shims  # noqa: E402
assert use_reference_on_shims(), "no reference tree"
from tests.helpers import OPERATORS_NPZ, seeded_state_dict  # noqa: E402
from tests.synthetic import HipLikeDataset, _euclid  # noqa: E402
from oracle.mesh_vae_oracle import load_operators, DEFAULT_CONFIG  # noqa: E402  (fixture loader only)
from models.cheb_VAE import cheb_VAE  # noqa: E402        the reference's model, unchanged
import utils as ref_utils  # noqa: E402                   the reference's utils.py (procrustes), unchanged
from torch_geometric.data import Data  # noqa: E402

STEPS, BATCH, N_MESH = 100, 16, 160


def curve(ds, ops, dropout, seed):
    A, D, U, nn_ = ops
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["dropout"] = dropout
    net = cheb_VAE(3, cfg, D, U, A, nn_, model=cfg["model"])
    net.load_state_dict(seeded_state_dict(net, 7))
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=5e-4)
    mean, std = torch.FloatTensor(ds.mean), torch.FloatTensor(ds.std)
    torch.manual_seed(seed)
    L, E = [], []
    for t in range(STEPS):
        items = [ds[(t * BATCH + j) % len(ds)] for j in range(BATCH)]
        data = Data(x=torch.cat([it[0].x for it in items], 0), edge_index=None, num_graphs=BATCH)
        x_gt = torch.stack([it[1] for it in items])
        hot = F.one_hot(torch.tensor([it[2] for it in items]), num_classes=2)
        opt.zero_grad()
        loss, correct, out, z, _ = net(data, x_gt, hot, m_type="train")
        loss.backward()
        opt.step()
        gt, R, m, s = (torch.stack([it[k] for it in items]) for k in (4, 5, 6, 7))
        rm = torch.bmm((out.detach() * std + mean) * s.unsqueeze(1), R) + m            # main.py:88-91
        L.append(float(loss.detach()))
        E.append(float(_euclid(rm.numpy(), gt.numpy()).mean()))                          # main.py:92-93
    return np.array(L), np.array(E)


def main():
    torch.set_num_threads(8)
    ops = load_operators(OPERATORS_NPZ)
    ds = HipLikeDataset(n=N_MESH, seed=666, procrustes_fn=ref_utils.procrustes)
    out = {}
    for tag, dropout, seed in (("exact", 0.0, 777), ("drop_a", 0.2, 1), ("drop_b", 0.2, 2)):
        out[f"{tag}_loss"], out[f"{tag}_err"] = curve(ds, ops, dropout, seed)
        print(tag, out[f"{tag}_loss"][[0, 1, 50, 99]], out[f"{tag}_err"][[0, 1, 50, 99]])
    np.savez_compressed(os.path.join(HERE, "golden_curves.npz"), **out)


if __name__ == "__main__":
    main()

"""Whole-model gradients with the tensor-core mesh layers on vs off (mvb_tune mesh_tc) at small and ragged batches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import meshvae_b200 as mvb  # noqa: E402
import bench  # noqa: E402
from tests.helpers import seeded_state_dict, seeded_batch, rel_err  # noqa: E402

dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
net.load_state_dict(seeded_state_dict(net, 7))
net.dropout.p = 0.0
net.train()
for B in (2, 3, 4, 7, 16, 64):
    x, y, eps = seeded_batch(B, nn_[0], 31)
    y_hot = torch.nn.functional.one_hot(y, 2).to(dev)
    grads = {}
    for mode in ("mesh_tc=0,0", "mesh_tc=1,0", "mesh_tc=1,1", "mesh_tc=1,2"):
        mvb._lib.tune(mode)
        net.zero_grad()
        loss, _, recon, _, _ = net(x.to(dev), x.double().to(dev), y_hot, m_type="train", eps=eps.to(dev))
        loss.backward()
        grads[mode] = ({n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}, float(loss), recon.detach().clone())
    base = grads["mesh_tc=0,0"]
    for mode in list(grads)[1:]:
        g, l, r = grads[mode]
        worst = max(((rel_err(g[n], base[0][n]), n) for n in g), key=lambda t: t[0])
        print(f"B={B} {mode}: loss {l:.6f} vs {base[1]:.6f}  recon err {rel_err(r, base[2]):.2e}  worst grad {worst[1]} {worst[0]:.2e}")
mvb._lib.tune("mesh_tc=1,0")

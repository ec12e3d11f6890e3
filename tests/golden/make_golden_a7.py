"""Pin for row A7 (PyG ChebConv as used by models/cheb_cls.py:18,76,95).  torch-geometric 2.0.4's source is not under
the reference tree, but the reference VENDORS the same operator: `ChebConv` at nn/conv.py:390-521 (its own `norm`
:464-483 = remove_self_loops -> get_laplacian('sym') -> 2/lambda_max -> add_self_loops(fill_value=-1), its own
recurrence :485-513; single weight[K,Fin,Fout] form, node_dim = 0, one graph per call).  This script runs THAT class,
unchanged, through the leaf shims on single meshes and stores inputs-by-seed / outputs / gradients; the tests check
oracle.pyg_cheb_conv (CPU) and meshvae_b200.conv.ChebConv (B200) against it with lins[k].weight = W_k^T.

Build container only:   python tests/golden/make_golden_a7.py      -> tests/golden/golden_a7.npz
"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.ref_loader import use_reference_on_shims  # noqa: E402
assert use_reference_on_shims(), "no reference tree"
from tests.helpers import OPERATORS_NPZ  # noqa: E402
from oracle.mesh_vae_oracle import load_operators  # noqa: E402  (fixture loader only)
from nn.conv import ChebConv  # noqa: E402   the reference's vendored class, unchanged

torch.set_num_threads(1)
from a7_inputs import CASES, case_inputs  # noqa: E402  (tests/golden/a7_inputs.py)


def main():
    A, D, U, nn_ = load_operators(OPERATORS_NPZ)
    out = {}
    for ci, (name, lvl, fin, fout, K, bias) in enumerate(CASES):
        n = nn_[lvl]
        x, w, b, dy = case_inputs(ci, n, fin, fout, K)
        conv = ChebConv(fin, fout, K, bias=bias)
        with torch.no_grad():
            conv.weight.copy_(w)
            if bias:
                conv.bias.copy_(b)
        ei = A[lvl]._indices()
        ys, dxs = [], []
        for m in range(x.shape[0]):                 # node_dim = 0: one graph [N, F] per call (nn/conv.py:106, :515)
            xm = x[m].clone().requires_grad_()
            y = conv(xm, ei)
            y.backward(dy[m])
            ys.append(y.detach())
            dxs.append(xm.grad)
        out[f"{name}_y"] = torch.stack(ys).numpy()
        out[f"{name}_dx"] = torch.stack(dxs).numpy()
        out[f"{name}_dw"] = conv.weight.grad.numpy()          # accumulated over the meshes
        if bias:
            out[f"{name}_db"] = conv.bias.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_a7.npz"), **out)
    print("wrote golden_a7.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

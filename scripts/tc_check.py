"""Quick A/B of the tcgen05 (3xTF32) contraction kernels against the FFMA kernels on the same inputs."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshvae_b200 as mvb
from tests.helpers import OPERATORS_NPZ, rel_err
import bench

dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
L = mvb._lib.lib
cases = [(1, 4, 16, 16, 6, True), (0, 4, 16, 16, 6, True), (3, 8, 16, 32, 6, True), (3, 8, 32, 32, 6, True),
         (2, 4, 32, 16, 6, True), (4, 2, 16, 3, 6, False), (0, 64, 16, 16, 6, True), (1, 3, 16, 16, 3, False)]
only = os.environ.get("ONLY")
for ci, (lvl, b, fin, fout, K, has_bias) in enumerate(cases):
    if only is not None and int(only) != ci:
        continue
    n = nn_[0] if lvl == 4 else nn_[lvl]
    ei, norm = mvb.ChebConv_batch.norm(A[lvl]._indices(), nn_[lvl])
    conv = mvb.ChebConv_batch(fin, fout, K, bias=has_bias).to(dev)
    conv.fuse_relu = True
    g = torch.Generator().manual_seed(ci)
    x = torch.randn(b, n, fin, generator=g).to(dev)
    dy = torch.randn(b, n, fout, generator=g).to(dev)
    res = {}
    for tc in (0, 1):
        L.mvb_set_tensor_cores(tc)
        conv.zero_grad()
        xg = x.clone().requires_grad_(True)
        y = conv(xg, ei, norm)
        y.backward(dy)
        torch.cuda.synchronize()
        res[tc] = (y.detach().clone(), xg.grad.clone(), conv.weight.grad.clone(),
                   conv.bias.grad.clone() if has_bias else None)
    errs = [rel_err(a, b_) for a, b_ in zip(res[1], res[0]) if a is not None]
    if os.environ.get("ORACLE") and ci == 6:
        from oracle import mesh_vae_oracle as O
        xo = x.cpu().clone().requires_grad_(True)
        wo = conv.weight.detach().cpu().clone().requires_grad_(True)
        bo = conv.bias.detach().cpu().clone().requires_grad_(True)
        yo = torch.relu(O.cheb_conv_batch(xo, ei.cpu(), norm.cpu(), wo, bo))
        yo.backward(dy.cpu())
        for tc in (0, 1):
            print("   vs oracle tc", tc, "y %.2e dx %.2e dw %.2e" % (rel_err(res[tc][0], yo), rel_err(res[tc][1], xo.grad), rel_err(res[tc][2], wo.grad)))
        d = (res[1][1] - res[0][1]).abs().permute(1, 0, 2).reshape(n, -1).amax(1)      # per-vertex max diff
        bad = (d > 1e-3 * res[0][1].abs().max()).nonzero().flatten()
        print("   bad vertices:", bad.numel(), bad[:20].tolist(), bad[-5:].tolist())
        db_ = (res[1][1] - res[0][1]).abs().amax(dim=(1, 2))
        print("   per-mesh max diff", [round(float(v), 4) for v in db_[:16]])
    print(f"case {ci} lvl{lvl} B{b} {fin}->{fout} K{K}: rel err tc vs ffma  y {errs[0]:.2e} dx {errs[1]:.2e} dw {errs[2]:.2e}"
          + (f" db {errs[3]:.2e}" if has_bias else ""), flush=True)
L.mvb_set_tensor_cores(1)

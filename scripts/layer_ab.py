"""A/B of the fused mesh-resident layers of the model (levels 1-3) at the benchmark batch: tensor-core mesh kernels
(mvb_tune mesh_tc=1, automatic / forced cluster size) vs the FFMA mesh kernels / the step-by-step composition.
CUDA-event time per forward and per backward call, median of `--iters` calls (operands L2-resident, as inside a step)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import meshvae_b200 as mvb  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
Fn, ops = mvb.functional, mvb.operators
B = a.batch
# (name, level, conv, up, down)
layers = []
for i in range(1, net.n_layers):
    layers.append((f"enc{i}", i, net.cheb[i], None, net.downsample_matrices[i]))
for i in range(net.n_layers - 1):
    lvl = net.n_layers - i - 1
    layers.append((f"dec{i}", lvl, net.cheb_dec[i], net.upsample_matrices[lvl], None))


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = {}
for mode in ["mesh_tc=1,0", "mesh_tc=1,1", "mesh_tc=1,2", "mesh_tc=0,0", "stepwise"]:
    if mode != "stepwise":
        mvb._lib.tune(mode)
    for name, lvl, conv, up, down in layers:
        l_op = ops.from_edges(net.A_edge_index[lvl], net.A_norm[lvl], net.A_num_nodes[lvl], dev)
        u_op = None if up is None else ops.from_sparse(up, dev)
        d_op = None if down is None else ops.from_sparse(down, dev)
        n = l_op.n_rows
        n_in = u_op.n_cols if u_op is not None else n
        k, fin, fout = conv.weight.shape
        x = torch.randn(n_in, B, fin, device=dev).requires_grad_()
        w = conv.weight.detach().clone().requires_grad_()
        bias = conv.bias.detach().clone().requires_grad_()
        if mode == "stepwise":
            def fwd():
                h = Fn.pool(x, u_op) if u_op is not None else x
                h = Fn.cheb_conv(h, w, bias, l_op, True)
                return Fn.pool(h, d_op) if d_op is not None else h
        else:
            if not Fn.cheb_layer_supported(n, B, fin, fout, k, l_op, u_op, d_op):
                continue

            def fwd():
                return Fn.cheb_layer(x, w, bias, l_op, u_op, d_op, relu=True)
        y = fwd()
        gy = torch.randn_like(y)

        def bwd():
            torch.autograd.grad(y, [x, w, bias], gy, retain_graph=True)
        with torch.no_grad():
            tf = timed(lambda: fwd(), a.iters)          # no autograd bookkeeping in the forward timing
        tb = timed(bwd, a.iters)
        out.setdefault(name, {})[mode] = {"fwd_us": round(tf, 1), "bwd_us": round(tb, 1)}
mvb._lib.tune("mesh_tc=1,0")
for name, modes in out.items():
    print(name, json.dumps(modes))
tot = {}
for name, modes in out.items():
    for m, v in modes.items():
        t = tot.setdefault(m, [0.0, 0.0, 0])
        t[0] += v["fwd_us"]; t[1] += v["bwd_us"]; t[2] += 1
print("totals (fwd_us, bwd_us, layers):", json.dumps(tot))

"""A/B of the level-0 decoder layer (pool(U0) -> conv 16->16 -> ReLU, models/cheb_VAE.py:284-285) at the benchmark batch:
row-streaming fused layer (mvb_cheb_stream_*, mvb_tune stream_tc=1) vs the step-by-step composition.  CUDA-event time per
forward / backward call, median of `--iters` calls, each after an L2 flush (cold) and back to back (warm)."""
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
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--modes", default="stream_tc=1,16;stream_tc=1,8;stream_tc=0", help="mvb_tune specs to compare, separated by ;")
a = ap.parse_args()
dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
Fn, ops = mvb.functional, mvb.operators
B = a.batch
conv, up = net.cheb_dec[net.n_layers - 1], net.upsample_matrices[0]
l_op = ops.from_edges(net.A_edge_index[0], net.A_norm[0], net.A_num_nodes[0], dev)
u_op = ops.from_sparse(up, dev)
k, fin, fout = conv.weight.shape
x = torch.randn(u_op.n_cols, B, fin, device=dev).requires_grad_()
w = conv.weight.detach().clone().requires_grad_()
bias = conv.bias.detach().clone().requires_grad_()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters, cold):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


out, res = {}, {}
modes = [m for m in a.modes.split(";") if m]
for mode in modes:
    mvb._lib.tune(mode)
    y = Fn.cheb_layer(x, w, bias, l_op, u_op, None, relu=True)
    gy = torch.randn(y.shape, device=dev, generator=torch.Generator(dev).manual_seed(1))
    g = torch.autograd.grad(y, [x, w, bias], gy, retain_graph=True)
    res[mode] = [y.detach()] + [t.detach() for t in g]

    def bwd():
        torch.autograd.grad(y, [x, w, bias], gy, retain_graph=True)
    with torch.no_grad():
        f = lambda: Fn.cheb_layer(x, w, bias, l_op, u_op, None, relu=True)      # noqa: E731
        out[mode] = {"fwd_warm_us": timed(f, a.iters, False), "fwd_cold_us": timed(f, a.iters, True)}
    out[mode].update({"bwd_warm_us": timed(bwd, a.iters, False), "bwd_cold_us": timed(bwd, a.iters, True)})
mvb._lib.tune("stream_tc=0,16")
names = ["y", "dx", "dw", "db"]
if len(modes) > 1:
    out["max_rel_diff"] = {n: float((p - q).abs().max() / q.abs().max()) for n, p, q in zip(names, res[modes[0]], res[modes[-1]])}
print(json.dumps(out))

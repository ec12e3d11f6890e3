"""Where does the time of the tensor-core mesh forward kernel go?  Back-to-back launches of one layer between two CUDA
events (the queue stays full, so the time per launch is the kernel's duration), for the probe settings of
mvb_tune mesh_dbg (1: no MMAs, 2: no recurrence steps, 4: no epilogue - results are then wrong, timing only)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import meshvae_b200 as mvb  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--once", action="store_true", help="one launch per layer and setting (under ncu)")
a = ap.parse_args()
dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
Fn, ops, L = mvb.functional, mvb.operators, mvb._lib
B = a.batch
cases = [("enc1", 1, net.cheb[1], None, net.downsample_matrices[1]), ("dec2", 1, net.cheb_dec[2], net.upsample_matrices[1], None),
         ("enc2", 2, net.cheb[2], None, net.downsample_matrices[2]), ("dec0", 3, net.cheb_dec[0], net.upsample_matrices[3], None)]
for name, lvl, conv, up, down in cases:
    l_op = ops.from_edges(net.A_edge_index[lvl], net.A_norm[lvl], net.A_num_nodes[lvl], dev)
    u_op = None if up is None else ops.from_sparse(up, dev)
    d_op = None if down is None else ops.from_sparse(down, dev)
    n_in = u_op.n_cols if u_op is not None else l_op.n_rows
    k, fin, fout = conv.weight.shape
    x = torch.randn(n_in, B, fin, device=dev)
    w, bias = conv.weight.detach(), conv.bias.detach()
    row = []
    for cmode in ("mesh_tc=1,2", "mesh_tc=1,1"):
        L.tune(cmode)
        if not Fn.cheb_layer_supported(l_op.n_rows, B, fin, fout, k, l_op, u_op, d_op):
            continue
        for dbg in (0, 1, 2, 3, 7):
            L.tune(f"mesh_dbg={dbg}")
            with torch.no_grad():
                if a.once:
                    Fn.cheb_layer(x, w, bias, l_op, u_op, d_op, relu=True)
                    continue
                for _ in range(3):
                    Fn.cheb_layer(x, w, bias, l_op, u_op, d_op, relu=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    Fn.cheb_layer(x, w, bias, l_op, u_op, d_op, relu=True)
                e1.record()
                torch.cuda.synchronize()
            row.append(f"{cmode[-1]}cta dbg{dbg}: {e0.elapsed_time(e1) * 1e3 / a.reps:6.1f} us")
        L.tune("mesh_dbg=0")
    print(name, " | ".join(row))
L.tune("mesh_tc=1,0;mesh_dbg=0")
torch.cuda.synchronize()
# phase time stamps of CTA 0 (%globaltimer, ns): 0 start, 1 alloc/init, 2 operator staged, 3 B staged, 4 T_0 written,
# 5 first barrier passed, 4+2k step k computed, 5+2k its barrier passed, 20 loop done, 21 last MMAs done, 22 epilogue, 23 end
import ctypes  # noqa: E402
L.lib.mvb_debug_mesh_prof.argtypes = [ctypes.c_void_p]
L.lib.mvb_debug_mesh_prof.restype = None
prof = torch.zeros(24, dtype=torch.int64, device=dev)
for cmode, dbgv in (("mesh_tc=1,2", 0), ("mesh_tc=1,1", 0), ("mesh_tc=1,2", 1), ("mesh_tc=1,2", 2), ("mesh_tc=1,2", 3)):
    L.tune(cmode + f";mesh_dbg={dbgv}")
    for name, lvl, conv, up, down in cases:
        l_op = ops.from_edges(net.A_edge_index[lvl], net.A_norm[lvl], net.A_num_nodes[lvl], dev)
        u_op = None if up is None else ops.from_sparse(up, dev)
        d_op = None if down is None else ops.from_sparse(down, dev)
        n_in = u_op.n_cols if u_op is not None else l_op.n_rows
        k, fin, fout = conv.weight.shape
        if not Fn.cheb_layer_supported(l_op.n_rows, B, fin, fout, k, l_op, u_op, d_op):
            continue
        x = torch.randn(n_in, B, fin, device=dev)
        with torch.no_grad():
            for it in range(3):
                prof.zero_()
                L.lib.mvb_debug_mesh_prof(prof.data_ptr())
                Fn.cheb_layer(x, conv.weight.detach(), conv.bias.detach(), l_op, u_op, d_op, relu=True)
                torch.cuda.synchronize()
                L.lib.mvb_debug_mesh_prof(None)
        t = prof.cpu().tolist()
        idx = [i for i in range(24) if t[i] > 0]
        print(name, cmode, f"dbg{dbgv}", "total %.1f us |" % ((t[23] - t[0]) / 1e3), " ".join(f"{i}:+{(t[i] - t[idx[n - 1]]) / 1e3:.1f}" for n, i in enumerate(idx) if n > 0))
L.tune("mesh_tc=1,0;mesh_dbg=0")

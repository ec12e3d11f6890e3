"""A/B the plane-group size of the 6-plane tcgen05 contraction: mvb_cheb_fwd on levels 0 / 1 (the 5 SpMM
steps are identical across variants; the difference is the contraction)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshvae_b200 as mvb
import bench
dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
L = mvb._lib
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
L.tune(f"fused_recurrence={0}")      # step-by-step recurrence: identical across variants, isolates the contraction
for lvl, B, F in [(0, 64, 16), (1, 64, 16), (1, 256, 16)]:
    n = nn_[lvl]
    ei, norm = mvb.ChebConv_batch.norm(A[lvl]._indices(), n)
    op = mvb.operators.from_edges(ei, norm, n, dev)
    x = torch.randn(n, B, F, device=dev)
    w = torch.randn(6, F, 16, device=dev) * 0.1
    bias = torch.randn(16, device=dev)
    ref = None
    for pg, cap in ((6, 4), (2, 4), (2, 3), (2, 2), (2, 1), (3, 2), (3, 1)):
        L.tune(f"tc_tuning={pg},{cap}")
        basis = torch.empty(5, n, B, F, device=dev); y = torch.empty(n, B, 16, device=dev)
        ms = {}
        for cold in (True, False):
            t = []
            for i in range(13):
                if cold: flush.fill_(float(i))
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                L.check(L.lib.mvb_cheb_fwd(n, B, F, 16, 6, n, op.nnz, L.ptr(op.rowptr), L.ptr(op.colidx), L.ptr(op.vals), L.ptr(x), L.ptr(w), L.ptr(bias), 1, L.ptr(basis), L.ptr(y), L.stream_ptr()))
                e.record(); e.synchronize()
                if i >= 3: t.append(s.elapsed_time(e))
            ms[cold] = sum(t) / len(t) * 1e3
        if ref is None: ref = y.clone()
        print(f"lvl{lvl} B{B} F{F} plane group {pg} cap {cap}: cheb_fwd cold {ms[True]:6.1f} us  warm {ms[False]:6.1f} us  max|dy| vs pg6 {float((y-ref).abs().max()):.2e}")
L.tune(f"tc_tuning={2},{4}"); L.tune(f"fused_recurrence={1}")

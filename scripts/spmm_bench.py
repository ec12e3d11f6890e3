"""A/B the SpMM variants on the level-0/1 operators: correctness (bit-identical) and cold-L2 timing."""
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
for lvl, B, F in [(0, 64, 16), (0, 64, 3), (0, 256, 16), (1, 64, 16), (2, 64, 16), (0, 16, 16)]:
    n = nn_[lvl]
    ei, norm = mvb.ChebConv_batch.norm(A[lvl]._indices(), n)
    op = mvb.operators.from_edges(ei, norm, n, dev)
    x = torch.randn(n, B, F, device=dev); z = torch.randn(n, B, F, device=dev)
    outs = {}
    for band in (0, 1):
        L.lib.mvb_set_spmm_band(band)
        y = torch.empty_like(x)
        ms = []
        for i in range(13):
            flush.fill_(float(i))
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            L.check(L.lib.mvb_spmm(n, n, L.ptr(op.rowptr), L.ptr(op.colidx), L.ptr(op.vals), L.ptr(x), L.ptr(y), L.ptr(z), None, 2.0, -1.0, B * F, L.stream_ptr()))
            e.record(); e.synchronize()
            if i >= 3: ms.append(s.elapsed_time(e))
        outs[band] = y.clone()
        u = n * B * F * 4
        alg = 3 * u + op.csr_bytes()
        t = sum(ms) / len(ms)
        print(f"lvl{lvl} B{B} F{F} band={band}: {t*1e3:7.1f} us  {alg/t/1e6:7.0f} GB/s ({alg/t/1e6/6548.2:.2f} of HBM peak)")
    print("   identical:", torch.equal(outs[0], outs[1]))

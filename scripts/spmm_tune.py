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
def run(lvl, B, F, tx, chunk, warm=False):
    n = nn_[lvl]
    ei, norm = mvb.ChebConv_batch.norm(A[lvl]._indices(), n)
    op = mvb.operators.from_edges(ei, norm, n, dev)
    x = torch.randn(n, B, F, device=dev); z = torch.randn(n, B, F, device=dev); y = torch.empty_like(x)
    L.lib.mvb_set_spmm_shape(tx, chunk)
    ms = []
    for i in range(13):
        if not warm: flush.fill_(float(i))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        L.check(L.lib.mvb_spmm(n, n, L.ptr(op.rowptr), L.ptr(op.colidx), L.ptr(op.vals), L.ptr(x), L.ptr(y), L.ptr(z), None, 2.0, -1.0, B * F, L.stream_ptr()))
        e.record(); e.synchronize()
        if i >= 3: ms.append(s.elapsed_time(e))
    u = n * B * F * 4
    return sum(ms) / len(ms) * 1e3, (3 * u + op.csr_bytes()) / (sum(ms) / len(ms)) / 1e6
for (lvl, B, F) in [(0, 64, 16), (0, 256, 16), (1, 64, 16)]:
    for warm in (False, True):
        print(f"--- lvl{lvl} B{B} F{F} {'warm' if warm else 'cold'}")
        for tx in (4, 8, 16, 32, 256):
            row = []
            for chunk in (0, 32, 64, 128, 256, 512, 1024):
                if chunk and chunk < 256 // tx: row.append("   -  "); continue
                t, gbs = run(lvl, B, F, tx, chunk, warm)
                row.append(f"{t:6.1f}")
            print(f"tx={tx:3d} chunk(auto,32,64,128,256,512,1024): " + " ".join(row))

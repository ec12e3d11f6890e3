import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshvae_b200 as mvb
L = mvb._lib
lib = L.lib
dev = torch.device("cuda:0")
N, B, Fin, Fout, K = 4998, int(os.environ.get("B", "64")), 16, 16, 6
g = torch.Generator().manual_seed(0)
x = torch.randn(N, B, Fin, generator=g).to(dev)
basis = torch.randn(K - 1, N, B, Fin, generator=g).to(dev)
dy = torch.randn(N, B, Fout, generator=g).to(dev)
w = (torch.randn(K, Fin, Fout, generator=g) * 0.1).to(dev)
rowptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
colidx = torch.zeros(1, dtype=torch.int32, device=dev)
vals = torch.zeros(1, device=dev)
ref_dx = dy.double() @ (w[0] - w[2] + w[4]).double().t()
T = torch.cat([x.unsqueeze(0), basis], 0).double()                       # [K,N,B,Fin]
ref_dw = torch.einsum("knbi,nbo->kio", T, dy.double())
ref_db = dy.double().sum((0, 1))
for tc in (0, 1):
    lib.mvb_set_tensor_cores(tc)
    for rep in range(2):
        dx = torch.zeros_like(x); dw = torch.zeros_like(w); db = torch.zeros(Fout, device=dev)
        nb = lib.mvb_cheb_bwd_workspace_bytes(N, B, Fin, Fout, K, N, 1)
        ws = torch.zeros(nb, device=dev, dtype=torch.uint8)
        L.check(lib.mvb_cheb_bwd(N, B, Fin, Fout, K, N, 0, L.ptr(rowptr), L.ptr(colidx), L.ptr(vals), L.ptr(x), L.ptr(basis),
                                 L.ptr(w), None, L.ptr(dy), L.ptr(dx), L.ptr(dw), L.ptr(db), L.ptr(ws), nb, L.stream_ptr()))
        torch.cuda.synchronize()
        e = (dx.double() - ref_dx).abs()
        rows_bad = (e.amax(2) > 1e-3).nonzero()
        print(f"tc {tc} rep {rep}: dx maxerr {float(e.max()):.3e} bad rows {rows_bad.shape[0]} first {rows_bad[:6].tolist()} "
              f"dw relerr {float((dw.double()-ref_dw).abs().max()/ref_dw.abs().max()):.2e} db relerr {float((db.double()-ref_db).abs().max()/ref_db.abs().max()):.2e}")
        if rows_bad.shape[0]:
            flat = (rows_bad[:, 0] * B + rows_bad[:, 1]).sort().values
            print("   flat bad rows min/max", int(flat[0]), int(flat[-1]), "tiles(64):", sorted(set((flat // 64).tolist()))[:30])
            r0 = rows_bad[0]
            print("   sample got", dx[r0[0], r0[1], :6].tolist(), "ref", ref_dx[r0[0], r0[1], :6].tolist())

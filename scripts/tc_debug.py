import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshvae_b200 as mvb
L = mvb._lib
lib = L.lib
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=220, precision=1, sci_mode=False)
N, B, Fin, Fout, K = 256, 1, 16, 16, 1
x = (torch.arange(Fin, device=dev).float() + 1).repeat(N, B, 1).contiguous()          # x[r][f] = f+1
x = x + 100 * (torch.arange(N, device=dev).float() % 8 == 3).view(N, 1, 1)             # row marker
dy = (torch.arange(Fout, device=dev).float() * 0.5 + 1).repeat(N, B, 1).contiguous()  # dy[r][o] = 1 + o/2
w = torch.zeros(K, Fin, Fout, device=dev)
for tc in (0, 1):
    lib.mvb_set_tensor_cores(tc)
    dw = torch.full((K, Fin, Fout), -7.0, device=dev)
    db = torch.full((Fout,), -7.0, device=dev)
    nb = lib.mvb_cheb_bwd_workspace_bytes(N, B, Fin, Fout, K, N, 0)
    ws = torch.zeros(nb, device=dev, dtype=torch.uint8)
    L.check(lib.mvb_cheb_bwd(N, B, Fin, Fout, K, N, 0, None, None, None, L.ptr(x), None, L.ptr(w), None, L.ptr(dy), None,
                             L.ptr(dw), L.ptr(db), L.ptr(ws), nb, L.stream_ptr()))
    torch.cuda.synchronize()
    print("tc", tc, "db", db.cpu())
    print(dw[0].cpu())
    part = ws[:4 * 20 * 16 * 2].view(torch.float32)
    print("partial block 0 rows 0..19:\n", part[:20 * 16].view(20, 16).cpu())

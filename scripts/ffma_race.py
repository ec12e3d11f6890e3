import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshvae_b200 as mvb
import bench
dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
L = mvb._lib.lib
L.mvb_set_tensor_cores(int(os.environ.get("TC", "0")))
B = int(os.environ.get("B", "64"))
ei, norm = mvb.ChebConv_batch.norm(A[0]._indices(), nn_[0])
conv = mvb.ChebConv_batch(16, 16, 6).to(dev)
conv.fuse_relu = True
g = torch.Generator().manual_seed(6)
x = torch.randn(B, nn_[0], 16, generator=g).to(dev).requires_grad_(True)
dy = torch.randn(B, nn_[0], 16, generator=g).to(dev)
y = conv(x, ei, norm)
y.backward(dy)
torch.cuda.synchronize()
print("done", float(x.grad.abs().sum()))

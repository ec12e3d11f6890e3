"""Run exactly ONE graph-replayed cheb_VAE training step inside a cudaProfilerStart/Stop window.
Use under:  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
(after the same command has exited 0 without ncu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
mvb, net, A, nn_ = bench.build_model(dev)
from meshvae_b200.engine import TrainEngine  # noqa: E402
eng = TrainEngine(net, a.batch, use_graph=not a.no_graph)
eng.capture(warmup=3)
g = torch.Generator().manual_seed(1)
x = torch.randn(a.batch, nn_[0], 3, generator=g).pin_memory()
eng.step(x, x.double().pin_memory(), torch.randint(0, 2, (a.batch,), generator=g))
for _ in range(3):
    eng.device_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.device_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step; loss", float(eng.loss), "launches/step", eng.launches_per_step)

"""Row f3 measurement: one epoch of the training loop at 64 meshes per batch, three ways, same captured step:
  (a) reference-style bookkeeping (main.py:83-94): three .cpu() reads, the [B,N,3] reconstruction copied to the host,
      de-normalisation + Procrustes bmm + vertex distances on the CPU, every batch;
  (b) loop.train_epoch: mvb_recon_error + mvb_epoch_meter_add on the device, one read-back per epoch;
  (c) the bare steps (engine.step with the loss read-back only) as the floor.
Synthetic hip-shaped meshes; host tensors pinned once (the data loader is not what is measured).
      python scripts/loop_overhead.py [batches]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import meshvae_b200 as mvb  # noqa: E402
from meshvae_b200 import engine, loop  # noqa: E402
import bench  # noqa: E402

B = 64
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
eng = engine.TrainEngine(net, B)
eng.capture()
n = nn_[0]
g = torch.Generator().manual_seed(0)
mean, std = torch.randn(n, 3, generator=g), torch.rand(n, 3, generator=g) + 0.5
batches = []
for i in range(4):
    x = torch.randn(B, n, 3, generator=g)
    batches.append((x.pin_memory(), x.double().pin_memory(), torch.randint(0, 2, (B,), generator=g), None,
                    torch.randn(B, n, 3, generator=g).pin_memory(), torch.randn(B, 3, 3, generator=g).pin_memory(),
                    torch.randn(B, 1, 3, generator=g).pin_memory(), (torch.rand(B, 1, generator=g) + 0.5).pin_memory()))


def euclid(a, b):
    return np.sqrt(((a - b) ** 2).sum(-1))


def reference_style():
    tot = [0.0, 0.0, 0.0, 0.0]
    for i in range(nb):
        x, x_gt, y, _, gt, R, m, s = batches[i % 4]
        eng.step(x, x_gt, y, sync=False)
        tot[0] += eng.loss.cpu().numpy() * B
        tot[1] += eng.kld.mean().cpu().numpy() * B
        tot[2] += eng.rec.mean().cpu().numpy() * B
        rm = eng.recon.cpu() * std + mean
        rm = torch.bmm(rm * s.unsqueeze(1), R) + m
        tot[3] += euclid(rm.numpy(), gt.numpy()).mean() * B
    return tot


def device_style():
    class L:
        def __iter__(self):
            return (batches[i % 4] for i in range(nb))
    return loop.train_epoch(eng, L(), norm=(mean, std))


def bare():
    for i in range(nb):
        x, x_gt, y, *_ = batches[i % 4]
        eng.step(x, x_gt, y)


out = {}
for name, fn in (("bare_steps", bare), ("reference_style_bookkeeping", reference_style), ("device_bookkeeping", device_style)):
    fn()                                    # warm-up epoch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name] = {"ms_per_batch": dt / nb * 1e3, "meshes_per_s": nb * B / dt}
print(json.dumps({"workload": f"epoch of {nb} batches x {B} meshes, captured step, pinned host batches", **out}))

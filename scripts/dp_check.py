"""2-rank data-parallel check (torchrun --nproc-per-node 2): the sharded engine (bucketed, overlapped all-reduce)
against a single-process engine that sees the GLOBAL batch, same weights, same noise, dropout off."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from tests.helpers import OPERATORS_NPZ, seeded_state_dict, seeded_batch
from oracle import mesh_vae_oracle as O          # operators / config only
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
import meshvae_b200 as mvb
from meshvae_b200.engine import TrainEngine
dev = torch.device("cuda")
A, D, U, nn_ = O.load_operators(OPERATORS_NPZ)
cfg = copy.deepcopy(O.DEFAULT_CONFIG); cfg["dropout"] = 0.0
def build():
    net = mvb.cheb_VAE(3, copy.deepcopy(cfg), [d.to(dev) for d in D], [u.to(dev) for u in U], [a.to(dev) for a in A], nn_)
    net.load_state_dict(seeded_state_dict(net, 11))
    return net.to(dev)
G = 8
per = G // world
sharded = TrainEngine(build(), per, distributed=True)
sharded.capture(warmup=1)
if rank == 0:
    print("gradient exchange:", f"peer memory ({sharded.peer.backend}), fused with Adam (mvb_dp_reduce_adam)" if sharded.peer is not None
          else "NCCL all-reduce", flush=True)
single = TrainEngine(build(), G, distributed=False)
single.world = 1
single.capture(warmup=1)
worst = 0.0
for s in range(5):
    x, y, eps = seeded_batch(G, nn_[0], 900 + s)
    lo, hi = rank * per, (rank + 1) * per
    l_loc = sharded.step(x[lo:hi], x[lo:hi].double(), y[lo:hi], eps_host=eps[lo:hi])
    t = torch.tensor([l_loc], device=dev, dtype=torch.float64)
    dist.all_reduce(t)
    l_dp = float(t) / world
    l_ref = single.step(x, x.double(), y, eps_host=eps)
    worst = max(worst, abs(l_dp - l_ref) / abs(l_ref))
    if rank == 0:
        print(f"step {s}: data-parallel mean loss {l_dp:.6f}  global-batch loss {l_ref:.6f}")
pd = torch.cat([p.detach().reshape(-1) for p in sharded.opt.params])
ps = torch.cat([p.detach().reshape(-1) for p in single.opt.params])
rel = float((pd - ps).norm() / ps.norm())
# replicas: every rank must hold bit-identical parameters (ordered sums on every rank)
reps = [torch.empty_like(sharded.opt.flat_p) for _ in range(world)]
dist.all_gather(reps, sharded.opt.flat_p)
same = all(torch.equal(reps[0], r) for r in reps[1:])
if sharded.peer is not None:
    assert int(sharded.peer.state[1]) == 0, "a wait for a peer's signal timed out"
if rank == 0:
    print(f"replicas bit-identical across ranks: {same}")
    assert same
    print(f"worst relative loss deviation {worst:.2e}; parameters after 5 steps differ by {rel:.2e} (L2)")
    assert worst < 1e-4 and rel < 1e-3
    print("DP_CHECK_OK", flush=True)
# graphs that captured NCCL work must go before the communicator does
sharded.release(); single.release()
del sharded, single
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()

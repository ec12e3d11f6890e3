"""Pin for row f3: the reference's OWN epoch loops, main.train (main.py:54-96) and main.evaluate (main.py:98-179),
imported unchanged through the leaf shims and driven with the reference's own cheb_VAE on tests/synthetic.SyntheticHips
(10 meshes, batch 4: the last batch is ragged; dropout 0, parameters from tests/helpers.seeded_state_dict(net, 7),
Adam(lr 1e-3, weight_decay 5e-4) as main.py:251; reparameterisation noise from torch.manual_seed(4321) on the global
CPU generator, cheb_VAE.py:316).  Stores their return tuples; tests/test_gpu_loop.py compares meshvae_b200.loop.train /
evaluate with them, tests/test_oracle_golden.py the CPU restatement used elsewhere in the tests.

Build container only:   python tests/golden/make_golden_loops.py      -> tests/golden/golden_loops.npz
"""
import copy
import os
import sys
import tempfile
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_loader import use_reference_on_shims  # noqa: E402
assert use_reference_on_shims(), "no reference tree"
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims_plot"))          # matplotlib stand-in (main.py:26)
from tests.helpers import OPERATORS_NPZ, seeded_state_dict  # noqa: E402
from tests.synthetic import SyntheticHips  # noqa: E402
from oracle.mesh_vae_oracle import load_operators, DEFAULT_CONFIG  # noqa: E402  (fixture loader only)
import main as ref_main  # noqa: E402                     the reference's driver, unchanged
from models.cheb_VAE import cheb_VAE  # noqa: E402        the reference's model, unchanged
from torch_geometric.data import Data, DataLoader  # noqa: E402  (leaf shim with PyG's collate semantics)

torch.set_num_threads(1)
N_MESH, BATCH, SEED = 10, 4, 4321


def main():
    A, D, U, nn_ = load_operators(OPERATORS_NPZ)
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["dropout"] = 0.0
    net = cheb_VAE(3, cfg, D, U, A, nn_, model=cfg["model"])
    net.load_state_dict(seeded_state_dict(net, 7))
    ds = SyntheticHips(n=N_MESH, seed=3, data_cls=Data)
    ckpt = tempfile.mkdtemp()
    np.savez(os.path.join(ckpt, "norm.npz"), mean=ds.mean, std=ds.std)          # data.py:166-173
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=5e-4)       # main.py:251
    out = {}

    def loader():
        # a private generator for the loader's own draws (torch's DataLoader takes its base seed from the GLOBAL
        # generator otherwise, which would interleave with the model's reparameterisation noise, cheb_VAE.py:316)
        return DataLoader(ds, batch_size=BATCH, shuffle=False, generator=torch.Generator().manual_seed(0))

    torch.manual_seed(SEED)
    ev0 = ref_main.evaluate(0, net, loader(), "cpu", checkpoint_dir=ckpt)
    for e in range(2):
        tr = ref_main.train(net, loader(), opt, "cpu", ckpt)
        out[f"train{e}"] = np.array([float(v) for v in tr], dtype=np.float64)   # loss, kld, rec_loss, error, accuracy
    ev1 = ref_main.evaluate(0, net, loader(), "cpu", checkpoint_dir=ckpt)
    for tag, ev in (("eval0", ev0), ("eval1", ev1)):
        out[f"{tag}_scalars"] = np.array([float(ev[0]), float(ev[1]), float(ev[2]), float(ev[3]), float(ev[5])], dtype=np.float64)
        out[f"{tag}_errors"] = np.asarray(ev[4], dtype=np.float32)             # [N_MESH, 4998] per-vertex errors
    np.savez_compressed(os.path.join(HERE, "golden_loops.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v if v.size <= 5 else (float(v.mean()), float(v.max())))


if __name__ == "__main__":
    main()

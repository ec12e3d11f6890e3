"""Row f4 (on-disk formats) and the host side of row f3 (collate / sharded loader): CPU-only checks against the
golden outputs of the reference's own helpers (tests/golden/make_formats_golden.py) and against torch.optim.Adam."""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN

import meshvae_b200  # noqa: F401  (loads the C-ABI library; no compute here)
from meshvae_b200 import formats, loader
from meshvae_b200.engine import FlatAdam


def test_read_config_matches_reference_golden():
    cfg = formats.read_config(os.path.join(GOLDEN, "default_like.cfg"))
    want = json.load(open(os.path.join(GOLDEN, "formats_config.json")))
    assert set(cfg) == set(want)
    for k, v in want.items():
        assert cfg[k] == v and type(cfg[k]) is type(v), k
    assert formats.read_config("/nonexistent/file.cfg") is None           # config_parser.py:50-52 prints and returns None


def test_read_config_is_tolerant_and_round_trips(tmp_path):
    text = open(os.path.join(GOLDEN, "default_like.cfg")).read()
    # single-space section name, a missing optional key, a key filed under the wrong section (hand-edited crecon.cfg style)
    text = text.replace("[ChebModel  Parameters]", "[ChebModel Parameters]").replace("workers_thread       = 6\n", "")
    text = text.replace("dropout               = 0.2\n", "").replace("[Input Output]\n", "[Input Output]\ndropout = 0.35\n")
    assert "workers_thread" not in text and text.count("dropout") == 1
    p = tmp_path / "edited.cfg"
    p.write_text(text)
    cfg = formats.read_config(str(p))
    assert cfg["n_layers"] == 4 and cfg["workers_thread"] == 0 and cfg["dropout"] == 0.35
    with pytest.raises(KeyError):
        formats.read_config(str(p), strict=True)
    q = tmp_path / "rt.cfg"
    formats.write_config(str(q), cfg)
    assert formats.read_config(str(q)) == cfg


def test_save_obj_is_byte_identical_to_reference_and_loads_back(tmp_path):
    d = np.load(os.path.join(GOLDEN, "formats_save_obj_input.npz"))
    out = tmp_path / "m.obj"
    formats.save_obj(str(out), d["v"], d["f"])
    assert out.read_bytes() == open(os.path.join(GOLDEN, "formats_save_obj.obj"), "rb").read()
    v, f = formats.load_obj(str(out))
    assert f.dtype == np.int32 and np.array_equal(f, d["f"])
    assert np.allclose(v, d["v"], atol=5e-7)                                # '%f' keeps 6 decimals


def test_load_obj_variants(tmp_path):
    p = tmp_path / "v.obj"
    p.write_text("# comment\nmtllib x.mtl\nv 0 0 0\nv 1 0 0 0.5 0.5 0.5\nvn 0 0 1\nv 1 1 0\nv 0 1 0\nvt 0 0\n"
                 "f 1/1/1 2/1/1 3/1/1\nf 1//1 3//1 4//1\nf -4 -3 -2 -1\n")
    v, f = formats.load_obj(str(p))
    assert v.shape == (4, 3) and v[1].tolist() == [1.0, 0.0, 0.0]
    assert f.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2], [0, 2, 3]]


def test_template_obj_round_trip(tmp_path):
    ops = np.load(os.path.join(GOLDEN, "operators_template5k.npz"))
    out = tmp_path / "t.obj"
    formats.save_obj(str(out), ops["template_v"], ops["template_f"])
    v, f = formats.load_obj(str(out))
    assert v.shape == (4998, 3) and np.array_equal(f, ops["template_f"])
    assert np.abs(v - ops["template_v"]).max() < 1e-6


def test_norm_history_and_reports(tmp_path):
    mean, std = np.random.rand(11, 3), np.random.rand(11, 3) + 1
    assert formats.save_norm(str(tmp_path), mean, std).endswith("norm.npz")
    m2, s2 = formats.load_norm(str(tmp_path))
    assert m2.dtype == torch.float32 and torch.allclose(m2, torch.FloatTensor(mean)) and torch.allclose(s2, torch.FloatTensor(std))
    tr = (np.float64(3.0), np.float64(0.5), np.float64(2.5), np.float64(1.25), np.float64(0.75))
    va = (np.float64(4.0), np.float64(0.6), np.float64(3.4), np.float64(0.5), np.ones((4, 11), np.float32) * 2, 0.25)
    h = [formats.history_entry(1, 10.0, 2.0, tr, va)]
    path = formats.save_history(str(tmp_path), 3, h)
    back = json.load(open(path))
    assert os.path.basename(path) == "history3.json"
    assert set(back[0]) == {"epoch", "begin", "duration", "training", "validation"}                      # main.py:282-304
    assert set(back[0]["training"]) == {"loss", "kld", "reconstruction_loss", "accuracy", "error"}
    assert set(back[0]["validation"]) == {"loss", "kld", "reconstruction_loss", "accuracy", "error", "sex_change_success_rate"}
    assert back[0]["validation"]["error"] == 2.0 and back[0]["training"]["accuracy"] == 0.75
    formats.save_inference_reports(str(tmp_path), ["/d/a_f_1.obj", "/d/b_m_2.obj"], [0, 1], [1.23456, 2.5], [3.0, 4.0])
    inf = json.load(open(tmp_path / "inference.json"))
    assert inf["a_f_1.obj"] == {"sex": 0, "reconstruction_error": {"mean": float(str(np.float32(1.23456))), "max": 3.0}}
    assert json.load(open(tmp_path / "pred.json")) == {"/d/a_f_1.obj": "0", "/d/b_m_2.obj": "1"}          # inference.py:80
    assert json.load(open(tmp_path / "error_list.json"))["/d/a_f_1.obj"] == "1.2346"                     # inference.py:122


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(5, 3)
        self.dead = torch.nn.Linear(2, 2)          # never used: no gradient, no Adam state (quirk 7)
        self.b = torch.nn.Linear(3, 1)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def test_checkpoint_dict_and_adam_state_interchange(tmp_path):
    """checkpoint written by the reference's loop (torch Adam) -> FlatAdam buffers -> back: identical state"""
    torch.manual_seed(0)
    net = _Tiny()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=5e-4)
    for _ in range(3):
        opt.zero_grad()
        net(torch.randn(4, 5)).sum().backward()
        opt.step()
    path = formats.save_model(net, opt, 2, 1.5, 2.5, str(tmp_path))
    assert os.path.basename(path) == "checkpoint_2.pt"
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"state_dict", "optimizer", "epoch_num", "train_loss", "val_loss"}                 # main.py:32-39
    # a second model + FlatAdam restored from that checkpoint
    net2 = _Tiny()
    live = [p for n, p in net2.named_parameters() if not n.startswith("dead")]
    flat = FlatAdam(live, lr=9.0, weight_decay=0.0)
    ck2 = formats.load_model(net2, path, optimizer=flat)
    assert ck2["epoch_num"] == 2 and flat.lr == 1e-3 and flat.weight_decay == 5e-4 and int(flat.step_count) == 3
    for (n1, p1), (n2, p2) in zip(net.named_parameters(), net2.named_parameters()):
        assert torch.equal(p1, p2), n1
    assert all(p.data_ptr() >= flat.flat_p.data_ptr() for p in live)            # still views of the flat buffer
    sd = formats.adam_state_dict(flat, net2)
    ref = opt.state_dict()
    assert sorted(sd["state"]) == sorted(ref["state"]) == [0, 1, 4, 5]
    for i in ref["state"]:
        assert float(sd["state"][i]["step"]) == float(ref["state"][i]["step"])
        assert torch.equal(sd["state"][i]["exp_avg"], ref["state"][i]["exp_avg"])
        assert torch.equal(sd["state"][i]["exp_avg_sq"], ref["state"][i]["exp_avg_sq"])
    # and torch's Adam accepts the exported dict
    opt3 = torch.optim.Adam(net2.parameters(), lr=1.0)
    opt3.load_state_dict(sd)
    assert opt3.param_groups[0]["lr"] == 1e-3 and opt3.param_groups[0]["weight_decay"] == 5e-4
    formats.save_model(net2, flat, 7, 0.0, 0.0, str(tmp_path))
    assert sorted(torch.load(tmp_path / "checkpoint_7.pt", weights_only=False)["optimizer"]["state"]) == [0, 1, 4, 5]


# ---- loader -----------------------------------------------------------------------------------------------
class _Item:
    def __init__(self, x, edge_index):
        self.x, self.y, self.edge_index = x, x, edge_index


class _Set:
    def __init__(self, n, nv=6):
        self.n, self.nv, self.ei = n, nv, torch.zeros(2, 4, dtype=torch.long)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        x = torch.full((self.nv, 3), float(i))
        return (_Item(x, self.ei), x.double(), i % 2, f"/data/m_{'f' if i % 2 == 0 else 'm'}_{i}.obj", x + 1, torch.eye(3),
                torch.zeros(1, 3), torch.ones(1))


def test_collate_matches_pyg_batch_semantics():
    ds = _Set(5)
    batch, x_gt, y, names, gt, R, m, s = loader.collate([ds[i] for i in (3, 1, 4)])
    assert batch.num_graphs == 3 and batch.x.shape == (18, 3) and batch.x[6, 0] == 1.0           # concatenated along dim 0
    assert x_gt.shape == (3, 6, 3) and x_gt.dtype == torch.float64
    assert y.dtype == torch.int64 and y.tolist() == [1, 1, 0]
    assert names == ["/data/m_m_3.obj", "/data/m_m_1.obj", "/data/m_f_4.obj"]
    assert gt.shape == (3, 6, 3) and R.shape == (3, 3, 3) and m.shape == (3, 1, 3) and s.shape == (3, 1)
    assert batch.to("cpu").x.shape == (18, 3)


@pytest.mark.parametrize("n,bs,world,drop", [(23, 4, 2, False), (24, 4, 2, False), (23, 4, 2, True), (10, 3, 4, False), (7, 8, 1, False)])
def test_sharded_loader_partitions_the_global_batches(n, bs, world, drop):
    ds = _Set(n)
    loaders = [loader.ShardedMeshLoader(ds, bs, shuffle=True, seed=11, rank=r, world=world, drop_last=drop) for r in range(world)]
    assert len({len(ld) for ld in loaders}) == 1
    per_rank = [list(ld.index_batches()) for ld in loaders]
    assert len({len(b) for b in per_rank}) == 1                      # same number of batches on every rank
    order = loaders[0].global_order()
    g = bs * world
    seen = []
    for i in range(len(loaders[0])):
        chunk = order[i * g:(i + 1) * g]
        got = np.concatenate([per_rank[r][i] for r in range(world)])
        if len(chunk) >= world:
            assert np.array_equal(got, chunk)                        # rank slices concatenate to the global batch, in order
        else:
            assert set(got.tolist()) == set(chunk.tolist())           # ranks without an item re-use the first one
        seen += chunk.tolist()
    assert sorted(seen) == (sorted(order[:n // g * g].tolist()) if drop else list(range(n)))
    # epochs reshuffle reproducibly
    a = loaders[0].global_order().copy()
    loaders[0].set_epoch(1)
    b = loaders[0].global_order()
    loaders[0].set_epoch(0)
    assert not np.array_equal(a, b) and np.array_equal(a, loaders[0].global_order())
    batches = list(loaders[0])
    assert len(batches) == len(loaders[0]) and batches[0][0].num_graphs == len(per_rank[0][0])

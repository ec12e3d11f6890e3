"""Synthetic MeshData-shaped datasets shared by the golden generators and the loop tests (no product imports)."""
import numpy as np
import torch
import torch.nn.functional as F

from tests.helpers import OPERATORS_NPZ


class SyntheticHips(torch.utils.data.Dataset):
    """Items of data.py:103-111 - (Data(x [N,3] f32, y, edge_index), ori_data [N,3] f64, label, filename, ori_mesh [N,3],
    R [3,3], m [1,3], s [1]) - for template + noise meshes under random similarity transforms.  `data_cls` is the
    container of the first field (the loaders only read .x / .y / .edge_index)."""

    def __init__(self, n=10, seed=3, data_cls=None):
        d = np.load(OPERATORS_NPZ)
        tv = d["template_v"]
        rng = np.random.default_rng(seed)
        self.aligned = [tv + rng.normal(size=tv.shape) * 0.8 for _ in range(n)]              # "mtx2" of the Procrustes fit
        self.mean, self.std = np.mean(self.aligned, 0), np.std(self.aligned, 0) + 1e-3
        self.R, self.s, self.m, self.ori = [], [], [], []
        for a in self.aligned:
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            s, m = rng.uniform(0.5, 2.0), rng.normal(size=(1, 3)) * 10
            self.R.append(torch.FloatTensor(q)); self.s.append(torch.FloatTensor([s])); self.m.append(torch.FloatTensor(m))
            self.ori.append(torch.Tensor((a * s) @ q + m + rng.normal(size=a.shape) * 0.05))   # the original scan
        self.labels = [int(v) for v in rng.integers(0, 2, n)]
        self.data_cls = data_cls or _Data

    def __len__(self):
        return len(self.aligned)

    def __getitem__(self, i):
        ori = (torch.tensor(self.aligned[i]) - torch.tensor(self.mean)) / torch.tensor(self.std)   # float64, data.py:106
        return (self.data_cls(x=ori.float(), y=ori.float(), edge_index=torch.zeros(2, 1, dtype=torch.long)), ori, self.labels[i],
                f"/scans/hip_{'fm'[self.labels[i]]}_{i}.obj", self.ori[i], self.R[i], self.m[i], self.s[i])


class _Data:
    def __init__(self, x=None, y=None, edge_index=None):
        self.x, self.y, self.edge_index = x, y, edge_index


# ---- CPU restatement of the reference's epoch loops on the oracle model (pinned against the reference's own
# main.train / main.evaluate by tests/golden/golden_loops.npz, tests/test_oracle_golden.py) ----
def _euclid(a, b):
    return np.sqrt(((a - b) ** 2).sum(-1))          # main.py:51-52


def ref_train(model, loader, optimizer, mean, std):
    """main.py:54-96 restated on the oracle model (host bookkeeping per batch, as the reference does it)"""
    model.train()
    tot = dict(n=0, loss=0.0, kld=0.0, rec=0.0, err=0.0, correct=0)
    for batch, x_gt, y, _, gt_mesh, R, m, s in loader:
        b = batch.num_graphs
        x = batch.x.reshape(b, -1, 3)
        hot = F.one_hot(y, num_classes=2)
        optimizer.zero_grad()
        loss, correct, out, z, _ = model(x, x_gt, hot, m_type="train")
        loss.backward()
        optimizer.step()
        tot["n"] += b
        tot["loss"] += loss.detach().numpy() * b
        tot["kld"] += z[0].mean().detach().numpy() * b
        tot["rec"] += z[1].mean().detach().numpy() * b
        tot["correct"] += int(correct)
        rm = torch.bmm((out.detach() * std + mean) * s.unsqueeze(1), R) + m
        tot["err"] += _euclid(rm.numpy(), gt_mesh.numpy()).mean() * b
    n = tot["n"]
    return tot["loss"] / n, tot["kld"] / n, tot["rec"] / n, tot["err"] / n, tot["correct"] / n


def ref_evaluate(model, loader, mean, std):
    """main.py:98-180 restated (vis=False)"""
    model.eval()
    tot = dict(n=0, loss=0.0, kld=0.0, rec=0.0, correct=0, acc=0)
    errors, metas = [], []
    with torch.no_grad():
        for batch, x_gt, y, _, gt_mesh, R, m, s in loader:
            b = batch.num_graphs
            x = batch.x.reshape(b, -1, 3)
            hot = F.one_hot(y, num_classes=2)
            loss, correct, out, z, _ = model(x, x_gt, hot, m_type="test")
            tot["n"] += b
            tot["loss"] += loss.numpy() * b
            tot["kld"] += z[0].mean().numpy() * b
            tot["rec"] += z[1].mean().numpy() * b
            tot["correct"] += int(correct)
            rm = torch.bmm((out * std + mean) * s.unsqueeze(1), R) + m
            errors.append(_euclid(rm.numpy(), gt_mesh.numpy()))
            oppo = 1 - hot
            oppo_x = model.sample(oppo, z[2])
            pred = torch.argmax(model.classifier(model.encoder(oppo_x)), 1)
            tot["acc"] += int((pred == torch.argmax(oppo, 1)).sum())
            metas.append((torch.bmm((oppo_x * std + mean) * s.unsqueeze(1), R) + m).numpy())
    n = tot["n"]
    return (tot["loss"] / n, tot["kld"] / n, tot["rec"] / n, tot["correct"] / n, np.concatenate(errors, 0), tot["acc"] / n,
            np.concatenate(metas, 0))



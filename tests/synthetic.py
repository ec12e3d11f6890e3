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




# ---------------------------------------------------------------------------------------------
# "hip-bone-shaped" synthetic scans (SURVEY.md 8(d) loss-curve inputs) and the reference's data preparation
# ---------------------------------------------------------------------------------------------
def procrustes(data1, data2):
    """Restatement of the reference's utils.procrustes (utils.py:58-156; scipy's procrustes that also returns the
    similarity transform): both sets centred and scaled to unit Frobenius norm, data2 rotated / scaled onto data1.
    -> (mtx1, mtx2, disparity, [R, norm2 / s, mean2]).  Pinned against the reference's function by
    tests/test_cpu_formats.py::test_procrustes_restatement_matches_reference."""
    from scipy.linalg import orthogonal_procrustes
    mtx1 = np.array(data1, dtype=np.double, copy=True)
    mtx2 = np.array(data2, dtype=np.double, copy=True)
    mean2 = np.mean(mtx2, 0)
    mtx1 -= np.mean(mtx1, 0)
    mtx2 -= np.mean(mtx2, 0)
    norm1, norm2 = np.linalg.norm(mtx1), np.linalg.norm(mtx2)
    mtx1 /= norm1
    mtx2 /= norm2
    R, s = orthogonal_procrustes(mtx1, mtx2)
    mtx2 = np.dot(mtx2, R.T) * s
    return mtx1, mtx2, float(np.sum(np.square(mtx1 - mtx2))), [R, norm2 / s, mean2]


def hip_like_scans(n, seed=666):
    """n scans [N,3] of the template under 8 smooth deformation modes (low-order polynomials of the centred template
    coordinates, amplitudes ~ N(0, 3 mm)), 0.5 mm vertex noise, a random rotation <= 10 degrees, scale U(0.9, 1.1) and a
    translation; labels = sign of the first amplitude (so that the classifier term is learnable)."""
    tv = np.load(OPERATORS_NPZ)["template_v"].astype(np.float64)
    c = tv - tv.mean(0)
    u = c / np.abs(c).max()
    x, y, z = u[:, 0:1], u[:, 1:2], u[:, 2:3]
    e = np.eye(3)
    modes = [x * e[0], y * e[1], z * e[2], (x * y) * e[2], (y * z) * e[0], (z * x) * e[1], (x * x - y * y) * e[0], (y * y - z * z) * e[2]]
    rng = np.random.default_rng(seed)
    scans, labels = [], []
    for _ in range(n):
        a = rng.normal(size=8) * 3.0
        shape = tv + sum(ai * m for ai, m in zip(a, modes)) + rng.normal(size=tv.shape) * 0.5
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        ang = np.deg2rad(rng.uniform(0, 10))
        Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        Rm = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
        scans.append(rng.uniform(0.9, 1.1) * shape @ Rm.T + rng.normal(size=(1, 3)) * 20)
        labels.append(int(a[0] > 0))
    return tv, scans, labels


class HipLikeDataset(torch.utils.data.Dataset):
    """MeshData-shaped items (data.py:103-111) made the way data.py:120-184 makes them: Procrustes fit of every scan
    against the template, per-vertex z-score with the mean / std of the fitted set, (R, s, m) kept for the
    back-transform of main.py:88-91."""

    def __init__(self, n=256, seed=666, data_cls=None, procrustes_fn=procrustes):
        tv, scans, labels = hip_like_scans(n, seed)
        fitted, self.R, self.s, self.m, self.ori = [], [], [], [], []
        for sc in scans:
            _, mtx2, _, res = procrustes_fn(tv, sc)
            fitted.append(mtx2.copy())
            self.ori.append(torch.Tensor(sc))
            self.R.append(torch.FloatTensor(res[0]))
            self.s.append(torch.FloatTensor([res[1]]))
            self.m.append(torch.FloatTensor(np.array([res[2]])))
        self.aligned = fitted
        self.mean, self.std = np.mean(fitted, axis=0), np.std(fitted, axis=0)          # data.py:166-170
        self.labels = labels
        self.data_cls = data_cls or _Data

    def __len__(self):
        return len(self.aligned)

    def __getitem__(self, i):
        ori = (torch.tensor(self.aligned[i]) - torch.tensor(self.mean)) / torch.tensor(self.std)      # data.py:106
        return (self.data_cls(x=ori.float(), y=ori.float(), edge_index=torch.zeros(2, 1, dtype=torch.long)), ori, self.labels[i],
                f"/scans/hip_{'fm'[self.labels[i]]}_{i}.obj", self.ori[i], self.R[i], self.m[i], self.s[i])

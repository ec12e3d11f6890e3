"""CPU ORACLE for the Mesh-VAE hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch-on-CPU fp32 arithmetic (the same ATen gather / multiply /
scatter_add / matmul sequence the reference executes), the algorithms of the reference's
hot path.  It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product package
(`mesh-vae_b200/`) never imports anything under `oracle/` and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF: `tests/golden/make_golden.py` imports the
unchanged `/root/reference/{nn/conv.py,nn/pool.py,logpdf.py,models/cheb_VAE.py,models/cheb_cls.py}`
through the leaf shims in `oracle/shims/` and stores seeded input/output vectors in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against them.
One exception: `pyg_cheb_conv` restates torch-geometric==2.0.4 `ChebConv`
(requirements.txt:33), whose source is not under /root/reference -> "parity unpinned" for that
class (it is additionally cross-checked against `cheb_conv_batch`, which IS pinned).

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LOG_2PI_HALF = 0.5 * float(np.log(2 * np.pi))


# --------------------------------------------------------------------------------------
# A1  MessagePassing.propagate  (nn/conv.py:242-331 ; __collect__ :171-229 ; aggregate :346-364)
# --------------------------------------------------------------------------------------
def propagate(x: torch.Tensor, gather_idx: torch.Tensor, scatter_idx: torch.Tensor,
              weight: torch.Tensor, dim_size: int) -> torch.Tensor:
    """out[t] = sum_{e: scatter_idx[e]==t} weight[e] * x[gather_idx[e]]  along dim 0 (node_dim=0).

    Mirrors the reference's three ATen passes: index_select (nn/conv.py:199-200), the
    `norm.view(-1,1,1) * x_j` message (nn/conv.py:579-581 / nn/pool.py:22-23) and
    torch_scatter.scatter(..., reduce='add') == zeros().scatter_add_ (nn/conv.py:363-364).
    """
    msg = x.index_select(0, gather_idx)
    msg = weight.view(-1, *([1] * (x.dim() - 1))) * msg
    out = torch.zeros((dim_size,) + tuple(x.shape[1:]), dtype=msg.dtype)
    idx = scatter_idx.view(-1, *([1] * (x.dim() - 1))).expand_as(msg)
    return out.scatter_add_(0, idx, msg)


# --------------------------------------------------------------------------------------
# A2  ChebConv_batch.norm  (nn/conv.py:541-555)
# --------------------------------------------------------------------------------------
def cheb_norm(edge_index: torch.Tensor, num_nodes: int,
              edge_weight: Optional[torch.Tensor] = None,
              dtype: Optional[torch.dtype] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Self loops removed, unit weights, norm_e = -deg^-1/2[row] * w_e * deg^-1/2[col], inf -> 0."""
    keep = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, keep]
    if edge_weight is not None:
        edge_weight = edge_weight[keep]
    else:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype).scatter_add_(0, row, edge_weight)
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    return edge_index, -dis[row] * edge_weight * dis[col]


# --------------------------------------------------------------------------------------
# A3  ChebConv_batch.forward  (nn/conv.py:557-577)
# --------------------------------------------------------------------------------------
def cheb_conv_batch(x: torch.Tensor, edge_index: torch.Tensor, norm: torch.Tensor,
                    weight: torch.Tensor, bias: Optional[torch.Tensor],
                    return_basis: bool = False):
    """x [B,N,Fin] -> [B,N,Fout];  out = (((X W0 + T1 W1) + T2 W2) + ...) + b with
    T1 = L X, Tk = 2 L T(k-1) - T(k-2) run vertex-major [N,B,F] (nn/conv.py:560);
    flow = source_to_target: gather edge_index[0], scatter at edge_index[1] (nn/conv.py:172).
    N is taken from x (size=None, nn/conv.py:160-169) - this is what lets the 20-node operator
    act on a 4998-node tensor (quirk 1)."""
    K = weight.size(0)
    out = torch.matmul(x, weight[0])
    xt = x.transpose(0, 1)
    n = xt.size(0)
    t_prev2 = xt
    basis = [xt]
    if K > 1:
        t_prev1 = propagate(xt, edge_index[0], edge_index[1], norm, n)
        basis.append(t_prev1)
        out = out + torch.matmul(t_prev1.transpose(0, 1), weight[1])
    for k in range(2, K):
        t_new = 2 * propagate(t_prev1, edge_index[0], edge_index[1], norm, n) - t_prev2
        basis.append(t_new)
        out = out + torch.matmul(t_new.transpose(0, 1), weight[k])
        t_prev2, t_prev1 = t_prev1, t_new
    if bias is not None:
        out = out + bias
    if return_basis:
        return out, basis
    return out


# --------------------------------------------------------------------------------------
# A5  SurfacePool.forward  (nn/pool.py:13-23)   and   A6  Pool  (models/cheb_cls.py:22-27)
# --------------------------------------------------------------------------------------
def surface_pool(x: torch.Tensor, indices: torch.Tensor, values: torch.Tensor,
                 shape: Sequence[int]) -> torch.Tensor:
    """out[b] = P x[b]; flow = target_to_source: gather indices[1] (cols), scatter at indices[0]
    (rows), dim_size = shape[0]; COO used uncoalesced as given (nn/pool.py:19)."""
    xt = x.transpose(0, 1)
    out = propagate(xt, indices[1], indices[0], values, int(shape[0]))
    return out.transpose(0, 1)


def pool_dim1(x: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, n_rows: int) -> torch.Tensor:
    """models/cheb_cls.py:22-27: index_select(x, 1, col) * value ; scatter_add over dim 1."""
    row, col = indices[0], indices[1]
    msg = torch.index_select(x, 1, col) * values.unsqueeze(-1)
    out = torch.zeros(x.size(0), n_rows, x.size(2), dtype=msg.dtype)
    idx = row.view(1, -1, 1).expand_as(msg)
    return out.scatter_add_(1, idx, msg)


# --------------------------------------------------------------------------------------
# A7  torch-geometric 2.0.4 ChebConv as called at models/cheb_cls.py:95  (PARITY UNPINNED)
# --------------------------------------------------------------------------------------
def pyg_cheb_norm(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32):
    """lambda_max = 2, normalization 'sym': L_hat = (I - D^-1/2 A D^-1/2) - I, built as PyG does:
    off-diagonal -dis[r] dis[c], then +1 loops (get_laplacian) and -1 loops (add_self_loops,
    fill_value=-1) appended - explicit entries that cancel (cf. vendored nn/conv.py:464-487)."""
    keep = edge_index[0] != edge_index[1]
    ei = edge_index[:, keep]
    w = torch.ones(ei.size(1), dtype=dtype)
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, ei[0], w)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    w = -(dis[ei[0]] * w * dis[ei[1]])
    loops = torch.arange(num_nodes, dtype=torch.long).unsqueeze(0).repeat(2, 1)
    ei = torch.cat([ei, loops], dim=1)
    w = torch.cat([w, w.new_full((num_nodes,), 1.0)])
    w = (2.0 * w) / torch.tensor(2.0, dtype=dtype)
    w.masked_fill_(w == float("inf"), 0)
    ei = torch.cat([ei, loops], dim=1)
    w = torch.cat([w, w.new_full((num_nodes,), -1.0)])
    return ei, w


def pyg_cheb_conv(x: torch.Tensor, edge_index: torch.Tensor, lin_weights: Sequence[torch.Tensor],
                  bias: Optional[torch.Tensor]) -> torch.Tensor:
    """x [B,N,Fin] (node_dim=-2); lins[k].weight is [Fout,Fin]; out = sum_k lins[k](T_k) + bias."""
    n = x.size(-2)
    ei, w = pyg_cheb_norm(edge_index, n, x.dtype)

    def prop(t):
        msg = w.view(-1, 1) * t.index_select(-2, ei[0])
        out = torch.zeros_like(t)
        return out.scatter_add_(-2, ei[1].view(1, -1, 1).expand_as(msg), msg)

    t0 = x
    out = F.linear(t0, lin_weights[0])
    if len(lin_weights) > 1:
        t1 = prop(x)
        out = out + F.linear(t1, lin_weights[1])
    for wk in lin_weights[2:]:
        t2 = 2.0 * prop(t1) - t0
        out = out + F.linear(t2, wk)
        t0, t1 = t1, t2
    if bias is not None:
        out = out + bias
    return out


# --------------------------------------------------------------------------------------
# A10  logpdf.py:7-8 (KLD), :22-23 (gaussian_nll), :24-28 (softclip)
# --------------------------------------------------------------------------------------
def kld(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), -1)


def softclip(t: torch.Tensor, minimum: float) -> torch.Tensor:
    return minimum + F.softplus(t - minimum)


def gaussian_nll(mu: torch.Tensor, log_sigma: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return 0.5 * torch.pow((x - mu) / log_sigma.exp(), 2) + log_sigma + LOG_2PI_HALF


# --------------------------------------------------------------------------------------
# A9 / A11 / A12  cheb_VAE  (models/cheb_VAE.py:104-351)
# --------------------------------------------------------------------------------------
def _sparse_parts(m):
    """Accept a torch sparse COO tensor (as model.py:24-32 builds) or an (indices, values, shape) triple."""
    if isinstance(m, torch.Tensor):
        return m._indices(), m._values(), tuple(m.shape)
    return m


class OracleChebVAE(torch.nn.Module):
    """Restatement of cheb_VAE with the reference's parameter names (checkpoint contract):
    cheb.{i}.weight/.bias, cheb_dec.{i}.weight/.bias (cheb_dec[-1].bias is None,
    models/cheb_VAE.py:135), enc_lin, dec_lin, dec_lin_1 (dead), dec_lin_2, z_mean, z_log_var,
    classifier_layer."""

    class _Conv(torch.nn.Module):
        def __init__(self, fin, fout, k, bias=True):
            super().__init__()
            self.in_channels, self.out_channels = fin, fout
            self.weight = torch.nn.Parameter(torch.empty(k, fin, fout))
            self.bias = torch.nn.Parameter(torch.empty(fout)) if bias else None
            torch.nn.init.normal_(self.weight, 0, 0.1)          # nn/conv.py:536-538
            if self.bias is not None:
                torch.nn.init.normal_(self.bias, 0, 0.1)

        def forward(self, x, edge_index, norm):
            return cheb_conv_batch(x, edge_index, norm, self.weight, self.bias)

    def __init__(self, num_features: int, config: Dict, D, U, A, num_nodes: Sequence[int]):
        super().__init__()
        self.n_layers = config["n_layers"]
        self.filters = [num_features] + list(config["num_conv_filters"])   # cheb_VAE.py:109-111
        self.K = config["polygon_order"]
        self.D = [_sparse_parts(m) for m in D]
        self.U = [_sparse_parts(m) for m in U]
        self.A_edge_index, self.A_norm = zip(*[cheb_norm(_sparse_parts(A[i])[0], num_nodes[i])
                                               for i in range(len(num_nodes))])  # cheb_VAE.py:118-119
        f = self.filters
        self.cheb = torch.nn.ModuleList([self._Conv(f[i], f[i + 1], self.K[i]) for i in range(len(f) - 2)])
        self.cheb_dec = torch.nn.ModuleList([self._Conv(f[-i - 1], f[-i - 2], self.K[i])
                                             for i in range(len(f) - 1)])
        self.cheb_dec[-1].bias = None                                        # cheb_VAE.py:135
        self.num_class = config["num_classes"]
        self.z = config["num_style"]
        self.num_hidden = config["num_hidden"]
        nl = self.D[-1][2][0] * f[-1]
        self.classifier_layer = torch.nn.Linear(self.num_hidden, self.num_class)
        self.z_mean = torch.nn.Linear(self.num_hidden + self.num_class, self.z)
        self.z_log_var = torch.nn.Linear(self.num_hidden + self.num_class, self.z)
        self.enc_lin = torch.nn.Linear(nl, self.num_hidden)
        self.dec_lin = torch.nn.Linear(self.z + self.num_class, self.num_hidden)
        self.dec_lin_1 = torch.nn.Linear(self.z + self.num_class, self.num_hidden)   # dead (quirk 7)
        self.dec_lin_2 = torch.nn.Linear(self.num_hidden, nl)
        self.dropout = torch.nn.Dropout(p=config["dropout"])
        torch.nn.init.normal_(self.enc_lin.weight, 0, 0.1)                   # cheb_VAE.py:349-351
        torch.nn.init.normal_(self.dec_lin.weight, 0, 0.1)

    # cheb_VAE.py:261-273
    def encoder(self, x):
        for i in range(self.n_layers):
            x = F.relu(self.cheb[i](x, self.A_edge_index[i], self.A_norm[i]))
            x = surface_pool(x, *self.D[i])
        x = x.reshape(x.shape[0], self.enc_lin.in_features)
        return self.dropout(F.relu(self.enc_lin(x)))

    # cheb_VAE.py:253-258  (dropout on the already dropped-out x, quirk 8)
    def classifier(self, x):
        return F.softmax(self.classifier_layer(self.dropout(x)), dim=1)

    # cheb_VAE.py:275-292  (final conv on the COARSEST operator, quirk 1)
    def decoder(self, z):
        x = self.dropout(F.relu(self.dec_lin(z)))
        x = self.dropout(F.relu(self.dec_lin_2(x)))
        x = x.reshape(x.shape[0], -1, self.filters[-1])
        for i in range(self.n_layers):
            x = surface_pool(x, *self.U[-i - 1])
            lvl = self.n_layers - i - 1
            x = F.relu(self.cheb_dec[i](x, self.A_edge_index[lvl], self.A_norm[lvl]))
        return self.cheb_dec[-1](x, self.A_edge_index[-1], self.A_norm[-1])

    def sample(self, y, z):                                                  # cheb_VAE.py:294-305
        return self.decoder(torch.cat([y, z], -1)).reshape(z.shape[0], -1, self.filters[0])

    def reparameterize(self, mu, logvar, eps=None):                          # cheb_VAE.py:309-319
        std = torch.exp(logvar * 0.5)
        if eps is None:
            eps = torch.normal(mean=0, std=1, size=(mu.shape[0], logvar.shape[1]))
        return eps * std + mu

    def loss_function(self, x, recon_x, mu_z, logvar_z, y, y_hat):           # cheb_VAE.py:321-346
        k = kld(mu_z, logvar_z)
        log_sigma = softclip(torch.Tensor([1]), -6)
        rec = gaussian_nll(recon_x, log_sigma, x).sum(-1).sum(-1)
        correct = torch.sum(torch.argmax(y_hat, dim=1) == torch.argmax(y, dim=1))
        logqy = (y_hat * y).sum(-1).log()
        return (k + rec - 2 * logqy).mean(), correct, k, rec

    def forward(self, x, x_gt, y, m_type="test", eps=None):                  # cheb_VAE.py:190-251
        """x [B,N,3] (the reference reshapes data.x [B*N,3] with data.num_graphs, :195-200)."""
        b = x.shape[0]
        h = self.encoder(x.reshape(b, -1, self.filters[0]))
        y_hat = self.classifier(h)
        h = torch.cat([y, h], -1)
        mu, logvar = self.z_mean(h), self.z_log_var(h)
        z_ = self.reparameterize(mu, logvar, eps) if m_type == "train" else mu
        recon = self.decoder(torch.cat([y, z_], -1)).reshape(b, -1, self.filters[0])
        loss, correct, k, rec = self.loss_function(x_gt, recon, mu, logvar, y, y_hat)
        return loss, correct, recon, [k, rec, z_], y_hat


# --------------------------------------------------------------------------------------
# A8-consumer: cheb_GCN  (models/cheb_cls.py:55-114)
# --------------------------------------------------------------------------------------
class OracleChebGCN(torch.nn.Module):
    class _PygConv(torch.nn.Module):
        def __init__(self, fin, fout, k):
            super().__init__()
            self.lins = torch.nn.ModuleList([torch.nn.Linear(fin, fout, bias=False) for _ in range(k)])
            self.bias = torch.nn.Parameter(torch.zeros(fout))
            for lin in self.lins:
                a = math.sqrt(6.0 / (fin + fout))
                lin.weight.data.uniform_(-a, a)

        def forward(self, x, edge_index):
            return pyg_cheb_conv(x, edge_index, [l.weight for l in self.lins], self.bias)

    def __init__(self, num_feature: int, config: Dict, D, U, A, num_nodes: Sequence[int]):
        super().__init__()
        self.n_layers = config["n_layers"]
        self.filters = [num_feature] + list(config["num_conv_filters"])
        self.K = config["polygon_order"]
        self.D = [_sparse_parts(m) for m in D]
        self.A_edge_index = []
        for i in range(len(num_nodes)):
            ei = _sparse_parts(A[i])[0]
            self.A_edge_index.append(ei[:, ei[0] != ei[1]])                  # cheb_cls.py:71-73
        f = self.filters
        self.cheb = torch.nn.ModuleList([self._PygConv(f[i], f[i + 1], self.K[i]) for i in range(len(f) - 2)])
        self.enc_lin = torch.nn.Linear(self.D[-1][2][0] * f[-2], 128)        # cheb_cls.py:81 (filters[-2])
        self.cls_layer = torch.nn.Linear(128, config["num_classes"])
        torch.nn.init.normal_(self.enc_lin.weight, 0, 0.1)                   # cheb_cls.py:108-110
        torch.nn.init.normal_(self.cls_layer.weight, 0, 0.1)

    def forward(self, x):                                                    # cheb_cls.py:86-105
        b = x.shape[0]
        x = x.reshape(b, -1, self.filters[0])
        for i in range(self.n_layers):
            x = F.relu(self.cheb[i](x, self.A_edge_index[i]))
            x = pool_dim1(x, self.D[i][0], self.D[i][1], int(self.D[i][2][0]))
        x = x.reshape(b, self.enc_lin.in_features)
        return self.cls_layer(F.relu(self.enc_lin(x)))


# --------------------------------------------------------------------------------------
# f3  per-batch reconstruction error  (main.py:51-52, 88-93; inference.py:100-127)
# --------------------------------------------------------------------------------------
def recon_error(out: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, s: torch.Tensor, R: torch.Tensor,
                m: torch.Tensor, gt_mesh) -> Tuple[np.ndarray, np.ndarray]:
    """out [B,N,3] fp32; mean/std FloatTensor [N,3] (main.py:57-58); s [B,1], R [B,3,3], m [B,1,3] float64 as the
    DataLoader collates them; gt_mesh [B,N,3] float64.  Returns (diff.mean(-1), diff.max(-1)) per mesh."""
    recon_mesh = out.cpu() * std + mean                       # main.py:88 (fp32)
    s = s.unsqueeze(1)                                        # main.py:89
    recon_mesh = torch.bmm(recon_mesh * s, R) + m             # main.py:90 (promotes to fp64)
    recon_mesh = recon_mesh.detach().cpu().numpy()
    gt = np.asarray(gt_mesh)
    diff = np.sqrt(((recon_mesh - gt) ** 2).sum(-1))          # main.py:51-52 euclidean_distances
    return diff.mean(-1), diff.max(-1)                        # inference.py:126-127


# --------------------------------------------------------------------------------------
# fixtures
# --------------------------------------------------------------------------------------
DEFAULT_CONFIG = {            # files/default.cfg:8-9,17-21,26-33
    "n_layers": 4, "num_hidden": 512, "downsampling_factors": [4, 4, 4, 4],
    "polygon_order": [6, 6, 6, 6, 6], "num_conv_filters": [16, 16, 16, 32, 32],
    "num_classes": 2, "num_style": 16, "dropout": 0.2, "batch_size": 16,
    "learning_rate": 0.001, "weight_decay": 0.0005, "model": "optimal_sigma_VAE",
}


def load_operators(npz_path: str):
    """tests/golden/operators_template5k.npz -> lists of torch sparse COO tensors exactly as
    model.py:24-32,44-46 hands them over (uncoalesced, int64 / f32)."""
    d = np.load(npz_path)
    num_nodes = [int(v) for v in d["num_nodes"]]

    def mk(name, i):
        idx = torch.from_numpy(np.vstack((d[f"{name}{i}_row"], d[f"{name}{i}_col"]))).long()
        val = torch.from_numpy(d[f"{name}{i}_val"]).float()
        return torch.sparse_coo_tensor(idx, val, tuple(int(s) for s in d[f"{name}{i}_shape"]),
                                       check_invariants=False)

    A = [mk("A", i) for i in range(len(num_nodes))]
    D = [mk("D", i) for i in range(len(num_nodes) - 1)]
    U = [mk("U", i) for i in range(len(num_nodes) - 1)]
    return A, D, U, num_nodes

"""Training-loop parity on the GPU (-m gpu): TrainEngine (CUDA graphs, gradient sinks into the flat
buffer, fused Adam) against the CPU oracle model trained with torch.optim.Adam - the body of
main.py:67-85 - on the same weights, batches and reparameterisation noise.
north_star gate: loss curves within 1 % over 100 steps; asserted here at 1e-3 relative per step
(dropout off so that both sides are deterministic; measured ~1e-5)."""
import copy

import pytest
import torch

from tests.helpers import OPERATORS_NPZ, seeded_state_dict, seeded_batch, rel_err
from oracle import mesh_vae_oracle as O

pytestmark = pytest.mark.gpu


def _models(mvb, dropout):
    A, D, U, nn_ = O.load_operators(OPERATORS_NPZ)
    cfg = copy.deepcopy(O.DEFAULT_CONFIG)
    cfg["dropout"] = dropout
    ref = O.OracleChebVAE(3, cfg, D, U, A, nn_)
    ref.load_state_dict(seeded_state_dict(ref, 11))
    dev = torch.device("cuda:0")
    net = mvb.cheb_VAE(3, copy.deepcopy(cfg), [d.to(dev) for d in D], [u.to(dev) for u in U], [a.to(dev) for a in A], nn_)
    net.load_state_dict(seeded_state_dict(net, 11))
    return ref, net.to(dev), nn_


def test_loss_curve_matches_oracle_100_steps():
    import meshvae_b200 as mvb
    from meshvae_b200.engine import TrainEngine
    B, steps = 4, 100
    ref, net, nn_ = _models(mvb, 0.0)
    ref.train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=5e-4)          # main.py:251
    eng = TrainEngine(net, B, lr=1e-3, weight_decay=5e-4, x_gt_dtype=torch.float64, use_graph=True)
    eng.capture(warmup=2)
    batches = [seeded_batch(B, nn_[0], 100 + i) for i in range(4)]                # 4 batches, cycled (epochs)
    worst = 0.0
    for s in range(steps):
        x, y, eps = batches[s % len(batches)]
        opt.zero_grad()
        lo, *_ = ref(x, x.double(), torch.nn.functional.one_hot(y, 2), m_type="train", eps=eps)
        lo.backward()
        lg = eng.step(x.pin_memory(), x.double().pin_memory(), y.pin_memory(), eps_host=eps.pin_memory())
        if s == 0:
            # the flat exchange buffer holds exactly the oracle's gradients (sinks + the copied 3-channel convs)
            views = dict(zip([id(p) for p in eng.opt.params], eng.opt.grad_views))
            for (name, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
                if q.grad is None:
                    assert id(p) not in views, name
                else:
                    assert rel_err(views[id(p)], q.grad) < 1e-4, name
        opt.step()
        rel = abs(lg - float(lo.detach())) / abs(float(lo.detach()))
        worst = max(worst, rel)
        assert rel < 1e-3, f"step {s}: engine loss {lg} vs oracle {float(lo.detach())}"
    # The trained parameters: Adam normalises the step, so elements whose gradient is at rounding-noise
    # level drift by up to lr per step on either side.  Two runs of the CPU oracle itself whose gradients
    # differ by 1e-5 relative (the parity tolerance) end 100 steps 1-8 % apart per tensor (L2) with a loss
    # deviation of 6e-5 - measured with the reference arithmetic; the bound below is that envelope.
    po = dict(ref.named_parameters())
    for name, p in net.named_parameters():
        if name.startswith("dec_lin_1"):
            continue                                    # dead parameter (quirk 7): untouched on both sides
        a, r = p.detach().cpu().double(), po[name].detach().double()
        d = float((a - r).norm() / r.norm())
        assert d < 0.2, (name, d)
    kld, rec, correct = eng.stats()
    assert kld >= 0 and rec > 0 and 0 <= correct <= B
    print(f"worst relative loss deviation over {steps} steps: {worst:.2e}")


def test_graph_and_eager_engines_are_bit_identical():
    import meshvae_b200 as mvb
    from meshvae_b200.engine import TrainEngine
    B = 3
    losses = []
    for use_graph in (True, False):
        _, net, nn_ = _models(mvb, 0.0)
        eng = TrainEngine(net, B, use_graph=use_graph)
        eng.capture(warmup=1)
        out = []
        for s in range(5):
            x, y, eps = seeded_batch(B, nn_[0], 300 + s)
            out.append(eng.step(x, x.double(), y, eps_host=eps))
        losses.append(out)
    assert losses[0] == losses[1]


def test_prefetched_steps_equal_plain_steps():
    """input double-buffering (stage / step_prefetched) changes when the batch travels, not what is computed"""
    import meshvae_b200 as mvb
    from meshvae_b200.engine import TrainEngine
    B = 3
    batches = []
    for s in range(5):
        x, y, eps = seeded_batch(B, 4998, 500 + s)
        batches.append((x.pin_memory(), x.double().pin_memory(), y, eps))
    _, net, _ = _models(mvb, 0.0)
    eng = TrainEngine(net, B)
    eng.capture(warmup=1)
    plain = [eng.step(x, xg, y, eps_host=e) for x, xg, y, e in batches]
    _, net2, _ = _models(mvb, 0.0)
    eng2 = TrainEngine(net2, B)
    eng2.capture(warmup=1)
    eng2.stage(*batches[0])
    pre = [eng2.step_prefetched(batches[i + 1] if i + 1 < len(batches) else None) for i in range(len(batches))]
    assert pre == plain
    with pytest.raises(RuntimeError):
        eng2.step_prefetched()                        # nothing staged


def test_dropout_training_runs_and_decreases_loss():
    """dropout 0.2 as in files/default.cfg: masks change from replay to replay (device offset = Adam step
    counter) and the loss still goes down on a fixed batch."""
    import meshvae_b200 as mvb
    from meshvae_b200.engine import TrainEngine
    B = 8
    _, net, nn_ = _models(mvb, 0.2)
    eng = TrainEngine(net, B, lr=1e-3, weight_decay=5e-4)
    eng.capture(warmup=2)
    x, y, eps = seeded_batch(B, nn_[0], 77)
    ls = [eng.step(x, x.double(), y, eps_host=eps) for _ in range(60)]
    assert all(l == l for l in ls)                      # finite
    assert len(set(ls[:10])) == 10                      # fresh masks / updates every replay
    assert sum(ls[-10:]) / 10 < sum(ls[:10]) / 10


# ---------------------------------------------------------------------------------------------
# The north-star gate as BASELINE.json words it: "reconstruction error and loss curves within 1% over 100 steps on
# synthetic hip-bone-shaped meshes" - against curves of the UNCHANGED reference model (tests/golden/golden_curves.npz,
# made by tests/golden/make_golden_curves.py with the reference's own cheb_VAE, utils.procrustes, z-score and error
# formula): batches of 16, 100 steps.
# ---------------------------------------------------------------------------------------------
def _hip_setup(mvb, dropout):
    import os
    import numpy as np
    from tests.helpers import GOLDEN
    from tests.synthetic import HipLikeDataset
    A, D, U, nn_ = O.load_operators(OPERATORS_NPZ)
    cfg = copy.deepcopy(O.DEFAULT_CONFIG)
    cfg["dropout"] = dropout
    dev = torch.device("cuda:0")
    net = mvb.cheb_VAE(3, cfg, [d.to(dev) for d in D], [u.to(dev) for u in U], [a.to(dev) for a in A], nn_, model=cfg["model"])
    net.load_state_dict(seeded_state_dict(net, 7))
    ds = HipLikeDataset(n=160, seed=666)
    return net.to(dev), ds, np.load(os.path.join(GOLDEN, "golden_curves.npz")), dev


def _hip_batch(ds, t, B):
    items = [ds[(t * B + j) % len(ds)] for j in range(B)]
    x = torch.stack([it[0].x for it in items])
    x_gt = torch.stack([it[1] for it in items])
    y = torch.tensor([it[2] for it in items])
    gt, R, m, s = (torch.stack([it[k] for it in items]) for k in (4, 5, 6, 7))
    return x, x_gt, y, gt, R, m, s


def test_loss_and_reconstruction_error_curves_match_reference_exact():
    """dropout off: eager autograd + torch.optim.Adam on the native model, the reparameterisation noise drawn from the
    same CPU generator state as the reference drew it (cheb_VAE.py:316) - every step comparable.  Gate: 1 % on the loss
    and on the reconstruction error at every one of the 100 steps (measured 3e-4 / 8e-3: the first steps agree to 1e-8,
    then Adam's normalised updates amplify 1e-6 gradient differences - as they do between two runs of the reference
    itself); 3 % on the loss ABOVE its constant floor (28.8 k of the 29.8 k are the constant of the Gaussian NLL)."""
    import numpy as np
    import meshvae_b200 as mvb
    Fn = mvb.functional
    B, steps = 16, 100
    net, ds, gc, dev = _hip_setup(mvb, 0.0)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=5e-4)
    mean, std = torch.FloatTensor(ds.mean).to(dev), torch.FloatTensor(ds.std).to(dev)
    torch.manual_seed(777)
    L, E = [], []
    for t in range(steps):
        x, x_gt, y, gt, R, m, s = _hip_batch(ds, t, B)
        opt.zero_grad()
        loss, correct, out, z, _ = net(x.to(dev), x_gt.to(dev), torch.nn.functional.one_hot(y, 2).to(dev), m_type="train")
        loss.backward()
        opt.step()
        err, _ = Fn.recon_error(out.detach(), mean, std, s, R, m, gt)
        L.append(float(loss.detach()))
        E.append(float(err.mean()))
    L, E = np.array(L), np.array(E)
    floor = 4998 * 3 * (mvb.functional.LOG_SIGMA_DEFAULT + 0.5 * np.log(2 * np.pi))          # the constant part of the NLL
    dl = np.abs(L - gc["exact_loss"]) / gc["exact_loss"]
    dv = np.abs((L - floor) - (gc["exact_loss"] - floor)) / (gc["exact_loss"] - floor)
    de = np.abs(E - gc["exact_err"]) / gc["exact_err"]
    print(f"exact curves: loss dev max {dl.max():.2e}, loss-above-floor dev max {dv.max():.2e}, recon error dev max {de.max():.2e}")
    assert dl[:5].max() < 1e-6 and de[:5].max() < 1e-4, "the first steps must agree to rounding"
    assert dl.max() < 1e-2 and de.max() < 1e-2 and dv.max() < 3e-2


def test_loss_and_reconstruction_error_curves_match_reference_with_dropout():
    """dropout 0.2 (files/default.cfg:32) through the captured step engine (CUDA graphs, fused Adam, device dropout
    streams): masks cannot match across implementations, so the comparison is statistical.  The yardstick is the
    reference's own seed-to-seed spread (two reference runs, drop_a / drop_b): per step the loss within 1 %, and per
    20-step window the mean loss and the mean reconstruction error within max(1 %, 2 x the reference's spread in that
    window)."""
    import numpy as np
    import meshvae_b200 as mvb
    from meshvae_b200.engine import TrainEngine
    Fn = mvb.functional
    B, steps = 16, 100
    net, ds, gc, dev = _hip_setup(mvb, 0.2)
    eng = TrainEngine(net, B, lr=1e-3, weight_decay=5e-4, x_gt_dtype=torch.float64, use_graph=True)
    eng.capture(warmup=2)
    mean, std = torch.FloatTensor(ds.mean).to(dev), torch.FloatTensor(ds.std).to(dev)
    torch.manual_seed(3)
    L, E = [], []
    for t in range(steps):
        x, x_gt, y, gt, R, m, s = _hip_batch(ds, t, B)
        L.append(eng.step(x.pin_memory(), x_gt.pin_memory(), y.pin_memory()))
        err, _ = Fn.recon_error(eng.recon, mean, std, s, R, m, gt)
        E.append(float(err.mean()))
    L, E = np.array(L), np.array(E)
    la, lb, ea, eb = gc["drop_a_loss"], gc["drop_b_loss"], gc["drop_a_err"], gc["drop_b_err"]
    lref, eref = 0.5 * (la + lb), 0.5 * (ea + eb)
    assert (np.abs(L - lref) / lref).max() < 1e-2, "per-step loss within 1 % of the reference"
    rows = []
    for a in range(0, steps, 20):
        w = slice(a, a + 20)
        spread_l = abs(la[w].mean() - lb[w].mean()) / lref[w].mean()
        spread_e = abs(ea[w].mean() - eb[w].mean()) / eref[w].mean()
        dl = abs(L[w].mean() - lref[w].mean()) / lref[w].mean()
        de = abs(E[w].mean() - eref[w].mean()) / eref[w].mean()
        rows.append((a, dl, spread_l, de, spread_e))
        assert dl <= max(1e-2, 2 * spread_l), (a, dl, spread_l)
        assert de <= max(1e-2, 2 * spread_e), (a, de, spread_e)
    print("window: loss dev / ref spread, recon-error dev / ref spread:", [(a, f"{dl:.1e}/{sl:.1e}", f"{de:.1e}/{se:.1e}") for a, dl, sl, de, se in rows])
    assert E[-20:].mean() < 0.9 * E[:5].mean(), "the reconstruction error must have gone down"


def test_learning_rate_change_reaches_a_captured_step():
    """hyper-parameters were frozen into the captured graphs.  They now live in device memory: `opt.lr = ...`,
    `param_groups[0]['lr'] = ...` (main.py:266-269) and formats.load_adam_state_dict change the NEXT replayed step."""
    import meshvae_b200 as mvb
    from meshvae_b200.engine import TrainEngine
    B = 4
    _, net, nn_ = _models(mvb, 0.0)
    eng = TrainEngine(net, B, lr=1e-3, weight_decay=5e-4, x_gt_dtype=torch.float64, use_graph=True)
    eng.capture(warmup=2)
    x, y, eps = seeded_batch(B, nn_[0], 5)
    args = (x.pin_memory(), x.double().pin_memory(), y.pin_memory())
    eng.step(*args, eps_host=eps)
    p0 = eng.opt.flat_p.clone()
    eng.step(*args, eps_host=eps)
    d1 = float((eng.opt.flat_p - p0).abs().max())
    for g in eng.opt.param_groups:          # the reference's schedule loop, main.py:268-269
        g["lr"] = 0.0
    p1 = eng.opt.flat_p.clone()
    eng.step(*args, eps_host=eps)
    assert d1 > 0 and float((eng.opt.flat_p - p1).abs().max()) == 0.0, "lr = 0 must freeze the parameters of a replayed step"
    eng.opt.set_lr(1e-3)
    eng.step(*args, eps_host=eps)
    assert float((eng.opt.flat_p - p1).abs().max()) > 0
    eng.release()
    assert not any(hasattr(p, "_mvb_grad_sink") for p in net.parameters())

"""GPU parity of the fused dense bottleneck (models/cheb_VAE.py:206-223,
253-258, 270-281): mvb_linear_fwd/bwd and mvb_vae_heads_fwd/bwd through the C ABI against the same
layers written with torch on the CPU (torch.nn.functional.linear / relu / softmax - what the
reference's nn.Linear modules execute).  fp32 tolerance 1e-4 relative (max|a-b| / max|b|);
measured <= 2e-6.  Dropout is checked through properties: keep rate, inverted scaling, the
backward mask equal to the forward mask, fresh masks when the device offset changes."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def Fn():
    import meshvae_b200
    return meshvae_b200.functional


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize("m,k,n,relu,bias", [(64, 640, 512, True, True), (16, 18, 512, True, True), (64, 512, 640, True, True),
                                             (5, 7, 3, False, False), (130, 36, 20, False, True), (1, 640, 512, True, True),
                                             (256, 1284, 33, True, True)])
def test_linear_matches_torch(Fn, m, k, n, relu, bias):
    x, w = _rand(m, k, seed=1), _rand(n, k, seed=2, scale=k ** -0.5)
    b = _rand(n, seed=3) if bias else None
    gy = _rand(m, n, seed=4)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    br = b.clone().requires_grad_() if bias else None
    yr = F.linear(xr, wr, br)
    yr = F.relu(yr) if relu else yr
    yr.backward(gy)
    xg, wg = x.cuda().requires_grad_(), w.cuda().requires_grad_()
    bg = b.cuda().requires_grad_() if bias else None
    yg = Fn.linear(xg, wg, bg, relu=relu)
    yg.backward(gy.cuda())
    assert rel_err(yg, yr) < TOL
    assert rel_err(xg.grad, xr.grad) < TOL and rel_err(wg.grad, wr.grad) < TOL
    if bias:
        assert rel_err(bg.grad, br.grad) < TOL


@pytest.mark.parametrize("b", [2, 64, 100])
def test_linear_vertex_major_layouts(Fn, b):
    """enc_lin reads x.reshape(B, 640) from the [20, B, 32] pool output; dec_lin_2 writes
    x.reshape(B, 20, 32) as [20, B, 32] (models/cheb_VAE.py:270, :281)."""
    v, f, hdim = 20, 32, 96
    x_vm, w1, b1 = _rand(v, b, f, seed=1), _rand(hdim, v * f, seed=2, scale=0.04), _rand(hdim, seed=3)
    w2, b2 = _rand(v * f, hdim, seed=4, scale=0.1), _rand(v * f, seed=5)
    g_vm = _rand(v, b, f, seed=6)
    xr = x_vm.clone().requires_grad_()
    params_r = [t.clone().requires_grad_() for t in (w1, b1, w2, b2)]
    h = F.relu(F.linear(xr.permute(1, 0, 2).reshape(b, v * f), params_r[0], params_r[1]))
    o = F.relu(F.linear(h, params_r[2], params_r[3])).reshape(b, v, f).permute(1, 0, 2)
    o.backward(g_vm)
    xg = x_vm.cuda().requires_grad_()
    params_g = [t.cuda().requires_grad_() for t in (w1, b1, w2, b2)]
    hg = Fn.linear(xg, params_g[0], params_g[1], relu=True, x_vm=True)
    og = Fn.linear(hg, params_g[2], params_g[3], relu=True, y_vm_f=f)
    assert tuple(og.shape) == (v, b, f)
    og.backward(g_vm.cuda())
    assert rel_err(og, o) < TOL and rel_err(xg.grad, xr.grad) < TOL
    for a, r in zip(params_g, params_r):
        assert rel_err(a.grad, r.grad) < TOL


def test_linear_dropout_properties(Fn):
    m, k, n, p = 64, 128, 512, 0.2
    x, w = _rand(m, k, seed=1).cuda(), _rand(n, k, seed=2, scale=0.1).cuda()
    b = torch.full((n,), 3.0, device="cuda")           # large bias: almost every pre-activation is positive
    stream = Fn.DropoutStream(seed=123)
    off = torch.zeros((), dtype=torch.int64, device="cuda")
    stream.offset_dev = off
    base = Fn.linear(x, w, b, relu=True)
    xg = x.clone().requires_grad_()
    y = Fn.linear(xg, w, b, relu=True, p=p, rng=stream.site(0))
    kept = y != 0
    pos = base > 0
    rate = float((kept & pos).sum()) / float(pos.sum())
    assert abs(rate - (1 - p)) < 0.02, rate
    assert torch.allclose(y[kept], base[kept] / (1 - p), rtol=1e-6, atol=0)
    # same (seed, offset) -> same mask; the backward pass regenerates it
    y2 = Fn.linear(x, w, b, relu=True, p=p, rng=stream.site(0))
    assert torch.equal(y, y2)
    y.backward(torch.ones_like(y))
    ref = (kept.float() / (1 - p)) @ w
    assert rel_err(xg.grad, ref) < TOL
    # another site, another host offset and another DEVICE offset (graph replays) all give new masks
    assert not torch.equal(y, Fn.linear(x, w, b, relu=True, p=p, rng=stream.site(1)))
    off.add_(1)
    y3 = Fn.linear(x, w, b, relu=True, p=p, rng=stream.site(0))
    assert not torch.equal(y, y3)
    # with a device counter attached the host offset stays frozen (advancing both would give two steps the same sum);
    # without one, advance() is what makes the next call draw a new mask
    stream.advance()
    assert torch.equal(y3, Fn.linear(x, w, b, relu=True, p=p, rng=stream.site(0)))
    stream.offset_dev = None
    y4 = Fn.linear(x, w, b, relu=True, p=p, rng=stream.site(0))
    stream.advance()
    assert not torch.equal(y4, Fn.linear(x, w, b, relu=True, p=p, rng=stream.site(0)))


@pytest.mark.parametrize("b,train", [(16, True), (64, True), (3, False)])
def test_vae_heads_match_torch(Fn, b, train):
    hd, z, c = 512, 16, 2
    h = _rand(b, hd, seed=1)
    y = F.one_hot(torch.randint(0, c, (b,), generator=torch.Generator().manual_seed(2)), c)
    eps = _rand(b, z, seed=3) if train else None
    mods = [torch.nn.Linear(hd, c), torch.nn.Linear(hd + c, z), torch.nn.Linear(hd + c, z)]
    grads = [_rand(b, c, seed=4), _rand(b, z, seed=5), _rand(b, z, seed=6), _rand(b, z, seed=7), _rand(b, c + z, seed=8)]
    hr = h.clone().requires_grad_()
    y_hat = F.softmax(mods[0](hr), dim=1)
    hc = torch.cat([y, hr], -1)
    mu, lv = mods[1](hc), mods[2](hc)
    zz = mu + eps * torch.exp(0.5 * lv) if train else mu
    zcat = torch.cat([y, zz], -1)
    torch.autograd.backward([y_hat, mu, lv, zz, zcat], grads)
    ref_grads = [hr.grad] + [p.grad for m_ in mods for p in m_.parameters()]
    import copy
    gm = [copy.deepcopy(m_).cuda() for m_ in mods]
    for m_ in gm:
        m_.zero_grad()
    hg = h.cuda().requires_grad_()
    outs = Fn.vae_heads(hg, y.cuda(), None if eps is None else eps.cuda(), gm[0], gm[1], gm[2])
    for o, r in zip(outs, (y_hat, mu, lv, zz, zcat)):
        assert rel_err(o, r) < TOL
    torch.autograd.backward(list(outs), [g.cuda() for g in grads])
    got = [hg.grad] + [p.grad for m_ in gm for p in m_.parameters()]
    for a, r in zip(got, ref_grads):
        assert rel_err(a, r) < TOL


def test_vae_heads_classifier_dropout(Fn):
    """quirk 8: the classifier sees dropout(h) - a second mask - while the z heads see h itself."""
    b, hd, z, c, p = 64, 512, 16, 2, 0.2
    h = _rand(b, hd, seed=1).cuda()
    y = F.one_hot(torch.randint(0, c, (b,), generator=torch.Generator().manual_seed(2)), c).cuda()
    mods = [torch.nn.Linear(hd, c).cuda(), torch.nn.Linear(hd + c, z).cuda(), torch.nn.Linear(hd + c, z).cuda()]
    stream = Fn.DropoutStream(seed=9)
    o0 = Fn.vae_heads(h, y, None, *mods)
    o1 = Fn.vae_heads(h, y, None, *mods, p=p, rng=stream.site(1))
    assert torch.equal(o0[1], o1[1]) and torch.equal(o0[2], o1[2])      # mu / logvar untouched by dropout
    assert not torch.equal(o0[0], o1[0])                                # y_hat sees the mask
    # gradient w.r.t. the classifier weight reveals the mask: columns of dWc are sums over kept units only
    hg = h.clone().requires_grad_()
    o = Fn.vae_heads(hg, y, None, *mods, p=p, rng=stream.site(1))
    o[0][:, 0].sum().backward()
    assert torch.isfinite(hg.grad).all() and torch.isfinite(mods[0].weight.grad).all()


def test_recon_error_matches_reference_arithmetic():
    """row f3: the per-batch error of main.py:88-93 / inference.py:100-127 on the device, read in place from the
    (padded, vertex-major) decoder output, against the oracle's restatement (fp64: 1e-12 relative)."""
    import meshvae_b200
    from oracle import mesh_vae_oracle as O
    Fn = meshvae_b200.functional
    b, n = 5, 4998
    g = torch.Generator().manual_seed(3)
    buf = torch.randn(n, b, 4, generator=g)                           # decoder output buffer, 3 channels padded to 4
    out = buf.permute(1, 0, 2)[..., :3]                               # the model's [B,N,3] view
    mean, std = torch.randn(n, 3, generator=g), torch.rand(n, 3, generator=g) + 0.5
    s = torch.rand(b, 1, generator=g, dtype=torch.float64) + 0.5
    R = torch.linalg.qr(torch.randn(b, 3, 3, generator=g, dtype=torch.float64))[0]
    m = torch.randn(b, 1, 3, generator=g, dtype=torch.float64)
    gt = torch.randn(b, n, 3, generator=g, dtype=torch.float64) * 2
    ref_mean, ref_max = O.recon_error(out, mean, std, s, R, m, gt.numpy())
    buf_d = buf.cuda()
    out_d = buf_d.permute(1, 0, 2)[..., :3]
    got_mean, got_max = Fn.recon_error(out_d, mean, std, s, R, m, gt)
    assert rel_err(got_mean, torch.from_numpy(ref_mean)) < 1e-12 and rel_err(got_max, torch.from_numpy(ref_max)) < 1e-12
    got2, _ = Fn.recon_error(out_d.contiguous(), mean, std, s, R, m, gt)          # contiguous [B,N,3] input: copied path
    assert torch.equal(got2, got_mean)


def test_vae_heads_beyond_one_launch_batch(Fn):
    """more meshes than the one-launch heads kernel holds (VAE_HEADS_MAX_BATCH): chunked calls, gradients summed by
    autograd - the strong-scaling shapes (512 / 256 meshes per GPU) stay on the native kernels"""
    b, hd, z, c = Fn.VAE_HEADS_MAX_BATCH + 44, 512, 16, 2
    h = _rand(b, hd, seed=1)
    y = F.one_hot(torch.randint(0, c, (b,), generator=torch.Generator().manual_seed(2)), c)
    eps = _rand(b, z, seed=3)
    mods = [torch.nn.Linear(hd, c), torch.nn.Linear(hd + c, z), torch.nn.Linear(hd + c, z)]
    gmods = [torch.nn.Linear(hd, c).cuda(), torch.nn.Linear(hd + c, z).cuda(), torch.nn.Linear(hd + c, z).cuda()]
    for m, g in zip(mods, gmods):
        g.load_state_dict(m.state_dict())
    hr = h.clone().requires_grad_()
    y_hat = torch.softmax(mods[0](hr), 1)
    hc = torch.cat([y.float(), hr], 1)
    mu, lv = mods[1](hc), mods[2](hc)
    zz = mu + eps * torch.exp(0.5 * lv)
    loss = (y_hat * _rand(b, c, seed=4)).sum() + (zz * _rand(b, z, seed=5)).sum() + (mu * lv).sum()
    loss.backward()
    hg = h.cuda().requires_grad_()
    o = Fn.vae_heads(hg, y.cuda(), eps.cuda(), *gmods)
    lg = (o[0] * _rand(b, c, seed=4).cuda()).sum() + (o[3] * _rand(b, z, seed=5).cuda()).sum() + (o[1] * o[2]).sum()
    lg.backward()
    assert rel_err(o[0], y_hat) < 1e-5 and rel_err(o[3], zz) < 1e-5 and rel_err(hg.grad, hr.grad) < 1e-4
    for m, g in zip(mods, gmods):
        assert rel_err(g.weight.grad, m.weight.grad) < 1e-4 and rel_err(g.bias.grad, m.bias.grad) < 1e-4


# ---- data-parallel exchange fused with Adam (csrc/mvb_dp.cu): one GPU can check the arithmetic (world = 1) and the signal
# protocol (two "ranks" = two buffer sets and two streams on the same device) ----
def _dp_bufs(mvb, n, world, dev, seed):
    import ctypes
    g = torch.Generator().manual_seed(seed)
    grads = [torch.randn(n, generator=g).to(dev) for _ in range(world)]
    pads = [torch.zeros(int(mvb._lib.lib.mvb_dp_pad_bytes()) // 4, device=dev, dtype=torch.int32) for _ in range(world)]
    arr = ctypes.c_void_p * world
    return grads, pads, arr(*[t.data_ptr() for t in grads]), arr(*[t.data_ptr() for t in pads])


@pytest.mark.parametrize("world", [1, 2, 3])
def test_dp_reduce_adam_matches_allreduce_then_adam(world):
    import meshvae_b200 as mvb
    L, lib = mvb._lib, mvb._lib.lib
    dev = torch.device("cuda:0")
    n, cut = 4096 * 5 + 32, 1024
    grads, pads, gp, pp = _dp_bufs(mvb, n, world, dev, 3)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 5e-4, 1.0 / world], device=dev)
    g0 = torch.Generator().manual_seed(9)
    p0 = torch.randn(n, generator=g0).to(dev)
    # reference: ordered sum, then the single-GPU fused Adam
    p_ref, m_ref, v_ref = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    step_ref = torch.zeros((), device=dev, dtype=torch.int64)
    # ranks: replicas of p / m / v, one state each, one stream each
    reps = [(p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0), torch.zeros((), device=dev, dtype=torch.int64),
             torch.zeros(int(lib.mvb_dp_state_bytes()) // 4, device=dev, dtype=torch.int32), torch.cuda.Stream()) for _ in range(world)]
    torch.cuda.synchronize()
    for it in range(3):
        for gr in grads:
            gr.mul_(0.5).add_(0.1 * it)
        gsum = grads[0].clone()
        for gr in grads[1:]:
            gsum += gr
        L.check(lib.mvb_adam_step_hp(n, L.ptr(p_ref), L.ptr(gsum), L.ptr(m_ref), L.ptr(v_ref), L.ptr(step_ref), L.ptr(hyper),
                                     L.stream_ptr()), "adam")
        torch.cuda.synchronize()
        for r, (p, m, v, st, state, stream) in enumerate(reps):
            with torch.cuda.stream(stream):
                # two gradient buckets on two channels, the way the captured step issues them
                L.check(lib.mvb_dp_begin(world, r, 2, pp, L.ptr(state), stream.cuda_stream), "dp_begin")
                L.check(lib.mvb_dp_reduce_adam(world, r, 0, cut, n - cut, L.ptr(p), gp, L.ptr(m), L.ptr(v), None, L.ptr(st), 1,
                                               L.ptr(hyper), pp, L.ptr(state), 8, stream.cuda_stream), "dp_reduce_adam")
                L.check(lib.mvb_dp_reduce_adam(world, r, 1, 0, cut, L.ptr(p), gp, L.ptr(m), L.ptr(v), None, L.ptr(st), 0,
                                               L.ptr(hyper), pp, L.ptr(state), 0, stream.cuda_stream), "dp_reduce_adam")
        torch.cuda.synchronize()
        for p, m, v, st, state, _ in reps:
            assert int(state[1]) == 0, "a wait for a peer's signal timed out"
            assert int(state[0]) == it + 1 and int(st) == it + 1
            # every rank sums in rank order: replicas are bit-identical to each other ...
            assert torch.equal(p, reps[0][0]) and torch.equal(m, reps[0][1]) and torch.equal(v, reps[0][2])
            # ... and agree with "ordered sum, then mvb_adam_step_hp" (the same arithmetic in another kernel)
            assert torch.allclose(p, p_ref, rtol=1e-6, atol=1e-7) and torch.allclose(m, m_ref, rtol=1e-6, atol=1e-9)
            assert torch.allclose(v, v_ref, rtol=1e-6, atol=1e-12)

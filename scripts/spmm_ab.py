"""A/B the SpMM kernels (mvb_set_spmm_mode 1 = one-shot mapping, 0 = pipelined) on the level-0/1 operators.
Two timings per variant: (a) one launch between an event pair, L2 flushed before it; (b) a train of
launches over rotating operand sets whose total footprint is several times the L2 (every launch finds
its operands cold; launch / event overheads amortised).  Checks bit-identity against mode 1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import meshvae_b200 as mvb  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
_, net, A, nn_ = bench.build_model(dev)
L = mvb._lib
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def launch(op, n, x, y, z, ncols):
    L.check(L.lib.mvb_spmm(n, n, L.ptr(op.rowptr), L.ptr(op.colidx), L.ptr(op.vals), L.ptr(x), L.ptr(y), L.ptr(z), None,
                           2.0, -1.0, ncols, L.stream_ptr()))


def time_variant(op, n, B, F, sets):
    x, y, z = sets[0]
    ms = []
    for i in range(13):
        flush.fill_(float(i))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        launch(op, n, x, y, z, B * F)
        e.record(); e.synchronize()
        if i >= 3:
            ms.append(s.elapsed_time(e))
    single = sum(ms) / len(ms) * 1e3
    reps = 5
    for (x, y, z) in sets:          # warm-up train
        launch(op, n, x, y, z, B * F)
    torch.cuda.synchronize()
    # the train is captured in a CUDA graph (as the step engine runs these kernels): no host launch cost
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        launch(op, n, *sets[0][:1], sets[0][1], sets[0][2], B * F)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for (x, y, z) in sets:
            launch(op, n, x, y, z, B * F)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record(); e.synchronize()
    train = s.elapsed_time(e) / (reps * len(sets)) * 1e3
    return single, train


cases = [(0, 64, 16), (0, 256, 16), (0, 64, 4), (1, 64, 16)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for lvl, B, F in cases:
    n = nn_[lvl]
    ei, norm = mvb.ChebConv_batch.norm(A[lvl]._indices(), n)
    op = mvb.operators.from_edges(ei, norm, n, dev)
    u = n * B * F * 4
    alg = 3 * u + op.csr_bytes()
    nsets = max(2, int(4 * 126e6 / (3 * u)) + 1)
    sets = [(torch.randn(n, B, F, device=dev), torch.empty(n, B, F, device=dev), torch.randn(n, B, F, device=dev))
            for _ in range(nsets)]
    print(f"--- lvl{lvl} B{B} F{F}: {alg/1e6:.1f} MB per launch, {nsets} rotating operand sets")
    L.tune(f"spmm_mode={1}"); L.tune(f"spmm_shape={0},{0}")
    launch(op, n, *sets[0][:1], sets[0][1], sets[0][2], B * F)
    ref = sets[0][1].clone()
    variants = [(1, 0, 0), (0, 0, 0)] + [(m + st, tx, ch) for st in (0,) for m in (1, 2, 3) for tx in (16, 32)
                                          for ch in (64, 128, 256) if ch >= (256 << (m - 1)) // tx]
    if os.environ.get("SPMM_VARIANTS"):      # "mode:tx:chunk,mode:tx:chunk,..."
        variants = [tuple(int(t) for t in v.split(":")) for v in os.environ["SPMM_VARIANTS"].split(",")]
    for mode, tx, ch in variants:
        L.tune(f"spmm_mode={mode}"); L.tune(f"spmm_shape={tx},{ch}")
        sets[0][1].zero_()
        launch(op, n, sets[0][0], sets[0][1], sets[0][2], B * F)
        same = torch.equal(sets[0][1], ref)
        single, train = time_variant(op, n, B, F, sets)
        print(f"mode={mode} tx={tx:2d} chunk={ch:4d}: single {single:6.1f} us ({alg/single/1e3:5.0f} GB/s)   "
              f"train {train:6.1f} us ({alg/train/1e3:5.0f} GB/s = {alg/train/1e3/6548.2:.2f} of peak)   identical={same}")
L.tune(f"spmm_mode={0}"); L.tune(f"spmm_shape={0},{0}")

#!/usr/bin/env python
"""bench.py - train meshes/s of cheb_VAE (files/default.cfg, 4998-vertex template) on N B200s.

One "step" = one pass of the hot path over one batch of synthetic meshes: forward + backward +
(gradient all-reduce) + Adam (main.py:67-85).  Workload at N=1: BASELINE.json configs[1]
("cheb_VAE default.cfg training, batch 64 fp32, single B200"); for N>1 every rank keeps a 64-mesh
shard (weak scaling; N=8 is configs[2], global batch 512).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (see the task contract): `value` = device-resident throughput (CUDA events,
max over ranks, L2 flushed between steps), `e2e` = the same through TrainEngine.stage() / step_prefetched()
with HOST buffers (H2D of a batch from pinned memory + D2H of the loss inside every timed region; the batch
copied during step i is the one step i+1 consumes - input double-buffering), `roofline` for the dominant
kernel (the level-0 Chebyshev recurrence SpMM) timed live, `cpu_baseline` = the CPU oracle port of
the reference step timed on this box's host cores.  `--impl reference` times that CPU port alone.
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train_meshes_per_sec"
UNIT = "meshes/s"
PER_GPU_BATCH = 64
OPERATORS_NPZ = os.path.join(ROOT, "tests", "golden", "operators_template5k.npz")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL announces its version there) are sent to
# stderr by swapping the descriptors for the duration of the run; emit() writes to the real stdout
_REAL_STDOUT = None


def guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the UNCHANGED reference through the leaf shims, else the oracle port
# ------------------------------------------------------------------------------------------------
def physical_cores() -> int:
    """physical cores of the box (hyper-threads hurt the scatter/gather-bound reference step)"""
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
        if n:
            return int(n)
    except Exception:  # noqa: BLE001
        pass
    return max(1, (os.cpu_count() or 2) // 2) if (os.cpu_count() or 1) > 8 else (os.cpu_count() or 1)


def _cpu_model(workload: str):
    """-> (kind, step(batch) callable factory).  kind "reference": the reference's own models/cheb_VAE.py /
    models/cheb_cls.py, imported unchanged from /root/reference or the staged copy oracle/_ref/reference
    (oracle/make_ref.py) on top of the leaf shims in oracle/shims; "port": the oracle restatement (no reference tree)."""
    from oracle import ref_loader
    from oracle import mesh_vae_oracle as O     # fixture loader / port: the checker, used here ONLY as the timed CPU baseline
    A, D, U, nn_ = O.load_operators(OPERATORS_NPZ)
    cfg = copy.deepcopy(O.DEFAULT_CONFIG)
    ref = ref_loader.use_reference_on_shims()
    torch.manual_seed(666)
    if ref is not None:
        from models.cheb_VAE import cheb_VAE            # the reference's own file
        from models.cheb_cls import cheb_GCN
        from torch_geometric.data import Data
        vae = cheb_VAE(3, copy.deepcopy(cfg), D, U, A, nn_, model=cfg["model"])
        gcn = cheb_GCN(6, copy.deepcopy(cfg), D, U, A, nn_) if workload == "cls" else None
        wrap = lambda x: Data(x=x.reshape(-1, x.shape[-1]), edge_index=None, num_graphs=x.shape[0])      # noqa: E731
        kind = "reference"
    else:
        vae = O.OracleChebVAE(3, copy.deepcopy(cfg), D, U, A, nn_)
        gcn = O.OracleChebGCN(6, copy.deepcopy(cfg), D, U, A, nn_) if workload == "cls" else None
        wrap = lambda x: x      # noqa: E731
        kind = "port"
    return kind, vae, gcn, wrap, nn_, cfg


def cpu_reference_step_rate(batch: int, steps: int, warmup: int, threads: int, workload: str = "train"):
    """meshes/s (median step) of the reference's CPU path for one of the bench workloads:
    train - main.py:67-85 (forward, backward, Adam; default.cfg, dropout 0.2, x_gt fp64);
    infer - inference.py:88-114 per batch (classifier pass, test-mode forward, opposite-sex sample; no_grad);
    cls   - crecon.py:80-88 (cheb_GCN forward on [B,4998,6], CrossEntropyLoss, backward, Adam).
    -> (meshes/s, median ms/step, kind)"""
    torch.set_num_threads(threads)
    kind, vae, gcn, wrap, nn_, cfg = _cpu_model(workload)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, nn_[0], 3, generator=g)
    x_gt = x.double()
    y = torch.nn.functional.one_hot(torch.randint(0, 2, (batch,), generator=g), 2)
    if workload == "train":
        vae.train()
        opt = torch.optim.Adam(vae.parameters(), lr=cfg["learning_rate"], weight_decay=cfg["weight_decay"])

        def step():
            opt.zero_grad()
            loss, *_ = vae(wrap(x), x_gt, y, m_type="train")
            loss.backward()
            opt.step()
            float(loss)
    elif workload == "infer":
        vae.eval()

        def step():
            with torch.no_grad():
                h = vae.encoder(x)
                y_hot = torch.nn.functional.one_hot(vae.classifier(h).argmax(1), 2)
                loss, _, recon, (_, _, z_), _ = vae(wrap(x), x, y_hot, m_type="test")
                vae.sample((1 - y_hot).float(), z_)
                float(loss)
    else:
        gcn.train()
        opt = torch.optim.Adam(gcn.parameters(), lr=1e-3, weight_decay=5e-4)
        x6 = torch.randn(batch, nn_[0], 6, generator=g)
        yl = torch.randint(0, 2, (batch,), generator=g)

        def step():
            opt.zero_grad()
            loss = torch.nn.functional.cross_entropy(gcn(x6), yl)
            loss.backward()
            opt.step()
            float(loss)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return batch / med, 1e3 * med, kind


WORKLOAD_TEXT = {
    "train": "cheb_VAE default.cfg train step (fwd+bwd+Adam), 4998-vertex template, fp32 (x_gt fp64)",
    "infer": "inference.py per-batch device work (classifier pass + test-mode forward + opposite-sex sample), cheb_VAE default.cfg, fp32",
    "cls": "cheb_GCN (crecon.py) train step on [B,4998,6] (fwd, CrossEntropyLoss, bwd, Adam), K=6, fp32",
}
METRICS = {"train": METRIC, "infer": "inference_meshes_per_sec", "cls": "cls_train_meshes_per_sec"}


def cpu_baseline_block(workload: str, steps: int, warmup: int = 3):
    """the `cpu_baseline` object: the reference's CPU path on this box's host cores, all physical cores and one thread"""
    cores = physical_cores()
    val, ms, kind = cpu_reference_step_rate(16, steps, warmup, cores, workload)
    n1 = max(4, steps // 6)
    v1, ms1, _ = cpu_reference_step_rate(16, n1, 1, 1, workload)
    return {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d timed 16-mesh steps (files/default.cfg batch_size) of the %s, median %.0f ms/step, torch threads = %d "
                      "physical cores (os.cpu_count() = %s); 1 thread: %.1f meshes/s (median of %d steps, %.0f ms/step)"
                      % (steps, "reference's own model files on the leaf shims (oracle/shims)" if kind == "reference" else
                         "oracle CPU port of the reference (no reference tree on this box)", ms, cores, os.cpu_count(), v1, n1, ms1),
            "value_1thread": v1, "ms_per_step": ms}


def run_reference(args, rank):
    if rank != 0:
        return 0
    cores = physical_cores()
    sample_batch = 16       # files/default.cfg:26 batch_size - the reference's own CPU-runnable case
    steps = max(args.steps, 20)
    val, ms, kind = cpu_reference_step_rate(sample_batch, steps, args.warmup, cores, args.workload)
    sample = (f"{sample_batch}-mesh steps (files/default.cfg batch_size): {WORKLOAD_TEXT[args.workload]}; "
              + ("the reference's own model files, unchanged, on the leaf shims of oracle/shims" if kind == "reference"
                 else "oracle CPU port of the reference (no reference tree on this box)")
              + f"; median of {steps} timed steps, {cores} torch threads (physical cores)")
    line = {"impl": "reference", "metric": METRICS[args.workload], "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload], "batch_per_step": sample_batch, "device": "host CPU"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks: NVML polled from a thread every few ms (a 0.1-0.3 s timed region still yields tens of samples);
# nvidia-smi -lms as the fallback
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period = index, period_s
        self.sm, self.mx, self.reasons = [], [], set()
        self.rows, self.proc, self.thread, self._stop, self.how = [], None, None, False, None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            self.how = "nvml"
            return
        except Exception:  # noqa: BLE001
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            self.how = "nvidia-smi"
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.how == "nvml":
            self._stop = True
            self.thread.join(timeout=1)
        elif self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for r in self.rows:
                parts = [p.strip() for p in r.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    self.sm.append(float(parts[0])); self.mx.append(float(parts[1]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (nvml / nvidia-smi unavailable)"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.how}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_model(dev):
    import meshvae_b200 as mvb
    if os.environ.get("MVB_TUNE"):          # A/B runs of the tuning hooks (include/mvb.h: mvb_tune); recorded in the line's config
        mvb._lib.tune(os.environ["MVB_TUNE"])
    d = np.load(OPERATORS_NPZ)
    nn_ = [int(v) for v in d["num_nodes"]]

    def mk(name, i):
        idx = torch.from_numpy(np.vstack((d[f"{name}{i}_row"], d[f"{name}{i}_col"]))).long()
        val = torch.from_numpy(d[f"{name}{i}_val"]).float()
        return torch.sparse_coo_tensor(idx, val, tuple(int(s) for s in d[f"{name}{i}_shape"]),
                                       check_invariants=False).to(dev)

    A = [mk("A", i) for i in range(5)]
    D = [mk("D", i) for i in range(4)]
    U = [mk("U", i) for i in range(4)]
    cfg = {"n_layers": 4, "num_hidden": 512, "polygon_order": [6, 6, 6, 6, 6],
           "num_conv_filters": [16, 16, 16, 32, 32], "num_classes": 2, "num_style": 16, "dropout": 0.2,
           "model": "optimal_sigma_VAE"}
    torch.manual_seed(666)          # files/default.cfg:13 random_seeds
    net = mvb.cheb_VAE(3, cfg, D, U, A, nn_, model=cfg["model"]).to(dev)
    return mvb, net, A, nn_


def spmm_roofline(mvb, A, nn_, batch, dev, peaks, reps=6):
    """Time the dominant kernel alone: one Chebyshev recurrence step T_k = 2 L T_{k-1} - T_{k-2}
    on the level-0 operator with B*16 columns (the dec3 layer).  Algorithmic bytes per launch =
    3*u + csr, u = N*B*F*4.  The launches walk a ring of operand sets whose total
    footprint is > 4x the 126 MB L2, so every launch finds x / z cold ("inputs larger than L2");
    one CUDA-event pair brackets the whole train on the launch stream.  `single_launch_ms` is the same
    kernel between its own event pair after an L2 flush (includes ~4 us of launch + event latency)."""
    L = mvb._lib
    n = nn_[0]
    f = 16
    ei, norm = mvb.ChebConv_batch.norm(A[0]._indices(), n)
    op = mvb.operators.from_edges(ei, norm, n, dev)
    u = n * batch * f * 4
    alg = 3 * u + op.csr_bytes()
    nsets = max(3, int(4 * 126e6 / (3 * u)) + 1)
    sets = [(torch.randn(n, batch, f, device=dev), torch.randn(n, batch, f, device=dev),
             torch.empty(n, batch, f, device=dev)) for _ in range(nsets)]
    st = torch.cuda.current_stream()

    def launch(x, z, y):
        L.check(L.lib.mvb_spmm(n, n, L.ptr(op.rowptr), L.ptr(op.colidx), L.ptr(op.vals), L.ptr(x), L.ptr(y), L.ptr(z),
                               None, 2.0, -1.0, batch * f, L.stream_ptr()))

    for t in sets:
        launch(*t)
    torch.cuda.synchronize()
    # the train is replayed from a CUDA graph, as the step engine runs these kernels (no host launch cost
    # between launches); one CUDA-event pair on the replay stream brackets `reps` replays
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for t in sets:
            launch(*t)
    g.replay()
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(st)
    for _ in range(reps):
        g.replay()
    e.record(st)
    e.synchronize()
    avg_ms = s.elapsed_time(e) / (reps * nsets)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    single = []
    for i in range(13):
        flush.fill_(float(i))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        launch(*sets[0])
        e.record(st)
        e.synchronize()
        if i >= 3:
            single.append(s.elapsed_time(e))
    achieved = alg / (avg_ms * 1e-3) / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("spmm_v4_level0_b64_f16_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    return {"bound": "hbm", "kernel": "spmm_v4_kernel<z> level-0 recurrence step, B*F=%d cols" % (batch * f),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "algorithmic_bytes_per_launch": alg, "avg_launch_ms": avg_ms,
            "timing": "%d launches (graph replays) over %d rotating operand sets (%.0f MB > 4x L2), one event pair" % (reps * nsets, nsets, nsets * 3 * u / 1e6),
            "single_launch_ms": sum(single) / len(single),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}


# per-mesh algorithmic bytes of the level-0 / level-1 layers (fp32): "stepwise" = every kernel of
# the step-by-step decomposition reads / writes its operands once (SURVEY 8(d) formulas), "fused" = layer input + output
# (+ what the backward pass must re-read).  (fwd stepwise, fwd fused, bwd stepwise, bwd fused); pools of the layer included.
LAYER_BYTES = {
    "enc0 (level 0, 3->16, +D0)": (2818892 + 414876, 639748, 679728 + 414876, 679728),
    "dec3 (level 0, U0+, 16->16)": (8016812 + 539820, 899644, 10255916 + 539820, 2818876),
    "enc1 (level 1, 16->16, +D1)": (2005020 + 103792, 225004, 2565020 + 103792, 705004),
    "dec2 (level 1, U1+, 16->16)": (2005020 + 135036, 225004, 2565020 + 135036, 705004),
}
STEP_BYTES_STEPWISE, STEP_BYTES_FUSED = 52.68e6, 13.04e6          # whole training step per mesh (SURVEY Appendix B totals)


def layer_rooflines(mvb, net, nn_, batch, dev, peak_gbs, reps=12):
    """The four layers that carry ~95 % of the step's bytes, each as the model runs it (one call of the fused / composed
    layer, forward and backward), replayed from a CUDA graph with the L2 flushed before every replay (cold operands),
    one CUDA-event pair per replay on the replay stream.  achieved = stepwise algorithmic bytes / time: a fused layer
    moves fewer bytes than the stepwise model counts, so its fraction of the HBM peak may exceed what a stepwise chain
    could reach - `fused_frac` is the same time against the fused lower bound."""
    Fn, ops = mvb.functional, mvb.operators
    out = []
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    st = torch.cuda.current_stream()
    specs = [("enc0 (level 0, 3->16, +D0)", 0, net.cheb[0], None, net.downsample_matrices[0]),
             ("dec3 (level 0, U0+, 16->16)", 0, net.cheb_dec[3], net.upsample_matrices[0], None),
             ("enc1 (level 1, 16->16, +D1)", 1, net.cheb[1], None, net.downsample_matrices[1]),
             ("dec2 (level 1, U1+, 16->16)", 1, net.cheb_dec[2], net.upsample_matrices[1], None)]
    for name, lvl, conv, up, down in specs:
        n_in = up.shape[1] if up is not None else nn_[lvl]
        fin = conv.weight.shape[1]
        first = name.startswith("enc0")
        x = torch.randn(batch, n_in, fin, device=dev)
        if not first:
            x.requires_grad_()
        with torch.no_grad():
            xin = Fn.from_vertex_major(Fn.pack_input(x)) if first else x
        # autograd replays a node's backward on the stream of its forward: the forward that the captured backward
        # differentiates must itself have run on the capture stream
        cap = torch.cuda.Stream()
        cap.wait_stream(st)
        with torch.cuda.stream(cap):
            y = net._layer(xin, conv, lvl, up=up, down=down)
            gy = torch.randn_like(y)
        st.wait_stream(cap)
        torch.cuda.synchronize()
        params = [conv.weight, conv.bias] + ([] if first else [x])

        def fwd():
            with torch.no_grad():
                net._layer(xin, conv, lvl, up=up, down=down)

        def bwd():
            torch.autograd.grad(y, params, gy, retain_graph=True, allow_unused=True)      # (engine sinks: dW lands in the flat buffer)
        times = {}
        for tag, fn in (("fwd", fwd), ("bwd", bwd)):
            cap.wait_stream(st)
            with torch.cuda.stream(cap):
                for _ in range(3):
                    fn()
            st.wait_stream(cap)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap):
                fn()
            ts = []
            for i in range(reps + 3):
                flush.fill_(float(i))
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(st)
                g.replay()
                s1.record(st)
                s1.synchronize()
                if i >= 3:
                    ts.append(s0.elapsed_time(s1))
            times[tag] = statistics.median(ts)
        b = LAYER_BYTES[name]
        for tag, sw, fu in (("fwd", b[0], b[1]), ("bwd", b[2], b[3])):
            gbs = batch * sw / (times[tag] * 1e-3) / 1e9
            out.append({"layer": name, "pass": tag, "ms": times[tag], "stepwise_bytes": batch * sw, "fused_bytes": batch * fu,
                        "achieved_gbs": gbs, "frac": gbs / peak_gbs, "fused_frac": batch * fu / (times[tag] * 1e-3) / 1e9 / peak_gbs})
    return out


def run_ours(args, rank, world, local_rank):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    mvb, net, A, nn_ = build_model(dev)
    from meshvae_b200.engine import TrainEngine
    B = args.batch
    eng = TrainEngine(net, B, lr=1e-3, weight_decay=5e-4, x_gt_dtype=torch.float64, use_graph=not args.no_graph,
                      graph_comm=not args.host_comm)
    eng.capture(warmup=3)

    # synthetic batch in pinned host memory (z-score-like vertices)
    g = torch.Generator().manual_seed(1000 + rank)
    x_h = torch.randn(B, nn_[0], 3, generator=g).pin_memory()
    xgt_h = x_h.double().pin_memory()
    y_h = torch.randint(0, 2, (B,), generator=g).pin_memory()
    eng.step(x_h, xgt_h, y_h)        # loads the static device buffers once

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    st = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.device_step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    # ---- device-resident timing: K steps, CUDA events per step, L2 flushed between steps ----------
    evs = []
    for i in range(args.steps):
        flush.fill_(float(i))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        eng.device_step()
        e.record(st)
        evs.append((s, e))
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    # ---- end-to-end timing through the public call, host buffers, loss read back ------------------
    # Every timed step: one batch copied from pinned host memory (the NEXT step's - input double-buffering, as a
    # prefetching loader does it), one captured step, the loss read back; the event pair also covers the copy.
    eng.stage(x_h, xgt_h, y_h)
    for _ in range(min(3, args.warmup)):
        eng.step_prefetched((x_h, xgt_h, y_h))
    barrier()
    e2e_ms = 0.0
    last_loss = None
    for i in range(args.steps):
        flush.fill_(float(i))
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        last_loss = eng.step_prefetched((x_h, xgt_h, y_h))
        eng.wait_staged()
        e.record(st)
        e.synchronize()
        e2e_ms += s.elapsed_time(e)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    log(f"rank {rank}: device-timed {dev_ms / args.steps:.4f} ms/step, end to end {e2e_ms / args.steps:.4f} ms/step (before the max over ranks)")
    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return 0
    if not np.isfinite(last_loss):
        raise SystemExit(f"bench.py: loss is not finite ({last_loss})")

    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))
    roof = spmm_roofline(mvb, A, nn_, B, dev, peaks)
    peak_gbs = peaks.get("hbm_gbs", 6650.0)
    layers = layer_rooflines(mvb, net, nn_, B, dev, peak_gbs) if world == 1 else None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    tensor_pipe = None
    if os.path.exists(tpath):
        try:
            tensor_pipe = json.load(open(tpath)).get("tensor_pipe")
        except Exception:  # noqa: BLE001
            tensor_pipe = None
    step_s = dev_ms * 1e-3 / args.steps
    step_roof = {"stepwise_frac": B * STEP_BYTES_STEPWISE / step_s / 1e9 / peak_gbs,
                 "fused_frac": B * STEP_BYTES_FUSED / step_s / 1e9 / peak_gbs,
                 "stepwise_bytes_per_mesh": STEP_BYTES_STEPWISE, "fused_bytes_per_mesh": STEP_BYTES_FUSED,
                 "note": "whole step against the HBM peak: bytes of the step-by-step kernel model / of the layer-fused lower bound "
                         "(SURVEY.md Appendix B) divided by the device-timed step"}
    total_meshes = B * world * args.steps
    value = total_meshes / (dev_ms * 1e-3)
    e2e_val = total_meshes / (e2e_ms * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block("train", steps=40)          # ~15-25 s of host work; median step
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cheb_VAE files/default.cfg training step (fwd+bwd+allreduce+Adam), "
                                   "4998-vertex template, K=6, filters 16,16,16,32,32, dropout 0.2, x_gt fp64",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "flushed between timed steps (256 MiB write, outside the event pairs)",
                       "cuda_graph": not args.no_graph, "final_loss": last_loss, "mvb_tune": os.environ.get("MVB_TUNE") or None,
                       "step_graphs": (1 if eng.one_graph else (3 if eng.split else 2)) if not args.no_graph else 0,
                       "collectives": None if world == 1 else (
                           (f"none: mvb_dp_reduce_adam reads the peers' flat fp32 gradient buffers over NVLink ({eng.peer.backend} peer "
                            "memory), sums them in rank order and applies Adam in the same launch; "
                            + ("two buckets, the dense layers' under the encoder backward" if eng.split else "one bucket")
                            + ", inside the step graph") if eng.peer is not None else
                           ("NCCL all-reduce of the flat fp32 gradient in two buckets, "
                            + ("captured inside the step graph" if eng.one_graph else "host-launched between graphs"))),
                       "e2e_input": "double-buffered: each timed step copies the next batch from pinned host memory "
                                    "(on a copy stream, inside the event pair) while it computes the current one"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": eng.h2d_bytes(),
                    "d2h_bytes_per_step": eng.d2h_bytes(), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(eng.launches_per_step) * args.steps,
            "gpu_launches_per_step": int(eng.launches_per_step),
            "clocks": clocks, "roofline": roof, "roofline_layers": layers, "roofline_step": step_roof,
            "tensor_pipe": tensor_pipe, "cpu_baseline": cpu}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# secondary workloads (BASELINE.json configs[3] and configs[4]): not the headline line, same hygiene
# ------------------------------------------------------------------------------------------------
def _time_graph(fn, steps, warmup, dev):
    """capture fn() once in a CUDA graph, replay `steps` times with the L2 flushed between replays"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    st = torch.cuda.current_stream()
    for _ in range(warmup):
        g.replay()
    torch.cuda.synchronize()
    evs = []
    for i in range(steps):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        g.replay()
        b.record(st)
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps


def _time_graph_e2e(fn, h2d, d2h, steps, warmup, dev):
    """end-to-end per step: the step's input copied from pinned host memory, the captured graph replayed, the step's
    result read back to the host - all inside one CUDA-event pair (L2 flushed before it)"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    st = torch.cuda.current_stream()
    tot = 0.0
    for i in range(warmup + steps):
        flush.fill_(float(i))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        h2d()
        g.replay()
        d2h()
        b.record(st)
        b.synchronize()
        if i >= warmup:
            tot += a.elapsed_time(b)
    return tot / steps


def run_secondary(args, local_rank):
    """--workload infer: the device work of inference.py:63-131 per batch (classifier pass, full test-mode
    forward, opposite-sex sample: 2 encoder + 2 decoder passes, no_grad).  --workload cls: one cheb_GCN
    training step as crecon.py:80-88 runs it (forward on [B,4998,6], CrossEntropyLoss, backward, Adam).
    --workload dropin: delivery mode 1 - the reference's OWN cheb_VAE class (imported unchanged on compat/), accelerate(),
    torch.optim.Adam: what a user who only swaps the import path gets (eager PyTorch autograd around the native ops)."""
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    mvb, net, A, nn_ = build_model(dev)
    L = mvb._lib.lib
    B = args.batch
    g = torch.Generator().manual_seed(5)
    sampler = ClockSampler(local_rank)
    h2d_bytes = d2h_bytes = 0
    if args.workload == "infer":
        net.eval()
        x_h = torch.randn(B, nn_[0], 3, generator=g).pin_memory()
        x = x_h.to(dev)
        out = {}
        res_h = torch.empty(B, dtype=torch.int64).pin_memory()

        def fn():
            with torch.no_grad():
                h = net.encoder(x)                                   # classifier_(): inference.py:88, main.py:42-49
                yh = net.classifier(h)
                y_hot = torch.nn.functional.one_hot(yh.argmax(1), 2)
                loss, _, recon, (_, _, z_), _ = net(x, x, y_hot, m_type="test")      # inference.py:97
                out["oppo"] = net.sample((1 - y_hot).float(), z_)                    # inference.py:114
                out["loss"] = loss
                out["pred"] = yh.argmax(1)
        c0 = L.mvb_launch_count()
        sampler.start()
        ms = _time_graph(fn, args.steps, args.warmup, dev)
        launches = (L.mvb_launch_count() - c0) // 4          # 3 warm-up calls + the capture
        e2e_ms = _time_graph_e2e(fn, lambda: x.copy_(x_h, non_blocking=True), lambda: res_h.copy_(out["pred"], non_blocking=True),
                                 args.steps, args.warmup, dev)
        h2d_bytes, d2h_bytes = x_h.numel() * 4, res_h.numel() * 8
        assert torch.isfinite(out["loss"]) and torch.isfinite(out["oppo"]).all()
    elif args.workload == "cls":
        cfg = {"n_layers": 4, "polygon_order": [6, 6, 6, 6, 6], "num_conv_filters": [16, 16, 16, 32, 32], "num_classes": 2}
        cls = mvb.cheb_GCN(6, cfg, net.downsample_matrices, net.upsample_matrices, net.adjacency_matrices, nn_).to(dev)
        cls.train()
        x_h = torch.randn(B, nn_[0], 6, generator=g).pin_memory()
        y_h = torch.randint(0, 2, (B,), generator=g).pin_memory()
        x, y = x_h.to(dev), y_h.to(dev)
        opt = torch.optim.Adam(cls.parameters(), lr=1e-3, weight_decay=5e-4, capturable=True)
        out = {}
        res_h = torch.empty((), dtype=torch.float32).pin_memory()

        def fn():
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(cls(x), y)      # crecon.py:80-84
            loss.backward()
            opt.step()
            out["loss"] = loss.detach()
        c0 = L.mvb_launch_count()
        sampler.start()
        ms = _time_graph(fn, args.steps, args.warmup, dev)
        launches = (L.mvb_launch_count() - c0) // 4

        def h2d():
            x.copy_(x_h, non_blocking=True)
            y.copy_(y_h, non_blocking=True)
        e2e_ms = _time_graph_e2e(fn, h2d, lambda: res_h.copy_(out["loss"], non_blocking=True), args.steps, args.warmup, dev)
        h2d_bytes, d2h_bytes = x_h.numel() * 4 + y_h.numel() * 8, 4
        assert torch.isfinite(out["loss"])
    else:       # dropin
        from oracle import ref_loader          # locates the staged reference tree only (nothing of the oracle runs here)
        ref = ref_loader.reference_root()
        if ref is None:
            emit({"metric": "dropin_train_meshes_per_sec", "unavailable": "no reference tree on this box (oracle/make_ref.py stages it)"})
            return 0
        mvb.install_compat()
        sys.path.insert(1, ref)
        sys.path.insert(2, os.path.join(ROOT, "oracle", "shims"))       # the open3d / psbody leaves of utils.py only
        from models.cheb_VAE import cheb_VAE as RefVAE                  # the reference's own class, unchanged
        cfg = {"n_layers": 4, "num_hidden": 512, "polygon_order": [6, 6, 6, 6, 6], "num_conv_filters": [16, 16, 16, 32, 32],
               "num_classes": 2, "num_style": 16, "dropout": 0.2, "model": "optimal_sigma_VAE"}
        torch.manual_seed(666)
        rnet = RefVAE(3, cfg, net.downsample_matrices, net.upsample_matrices, net.adjacency_matrices, nn_, model=cfg["model"]).to(dev)
        mvb.accelerate(rnet)
        rnet.train()
        from torch_geometric.data import Data
        opt = torch.optim.Adam(rnet.parameters(), lr=1e-3, weight_decay=5e-4)
        x_h = torch.randn(B, nn_[0], 3, generator=g).pin_memory()
        xgt_h = x_h.double().pin_memory()
        y_h = torch.randint(0, 2, (B,), generator=g).pin_memory()
        st = torch.cuda.current_stream()
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

        def step():                                  # main.py:67-85, eager: H2D, one-hot, forward, backward, Adam, loss read-back
            x = x_h.to(dev, non_blocking=True)
            x_gt = xgt_h.to(dev, non_blocking=True)
            hot = torch.nn.functional.one_hot(y_h, 2).to(dev, non_blocking=True)
            opt.zero_grad()
            loss, *_ = rnet(Data(x=x.reshape(-1, 3), edge_index=None, num_graphs=B), x_gt, hot, m_type="train")
            loss.backward()
            opt.step()
            return float(loss)
        c0 = L.mvb_launch_count()
        for _ in range(max(3, args.warmup)):
            step()
        launches = (L.mvb_launch_count() - c0) // max(3, args.warmup)
        sampler.start()
        tot = 0.0
        for i in range(args.steps):
            flush.fill_(float(i))
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            last = step()
            b.record(st)
            b.synchronize()
            tot += a.elapsed_time(b)
        ms = e2e_ms = tot / args.steps
        h2d_bytes, d2h_bytes = x_h.numel() * 4 + xgt_h.numel() * 8 + B * 2 * 8, 8
        assert np.isfinite(last)
    clocks = sampler.stop()
    cpu = None
    if not args.no_cpu_baseline:
        if args.workload == "dropin":
            # this process has the reference's module names bound to the native classes (compat/): the CPU arm - the same
            # reference files on the leaf shims - runs in a process of its own
            import subprocess
            code = "import json, bench; print('CPU_BASELINE ' + json.dumps(bench.cpu_baseline_block('train', steps=30)))"
            res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=900)
            tag = [ln for ln in res.stdout.splitlines() if ln.startswith("CPU_BASELINE ")]
            cpu = json.loads(tag[-1][len("CPU_BASELINE "):]) if tag else {"unavailable": (res.stderr or res.stdout)[-300:]}
        else:
            cpu = cpu_baseline_block(args.workload, steps=30)
    work = {"infer": WORKLOAD_TEXT["infer"], "cls": WORKLOAD_TEXT["cls"],
            "dropin": "the reference's own cheb_VAE class (unchanged, imported on compat/) + accelerate() + torch.optim.Adam, eager "
                      "training step as main.py:67-85 runs it (H2D, forward, backward, optimizer, loss read-back), batch fp32 / x_gt fp64"}[args.workload]
    metric = {"infer": METRICS["infer"], "cls": METRICS["cls"], "dropin": "dropin_train_meshes_per_sec"}[args.workload]
    line = {"metric": metric, "value": B / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": work, "batch_per_gpu": B, "cuda_graph": args.workload != "dropin",
                                            "l2": "flushed between timed steps (256 MiB write, outside the event pairs)"},
            "e2e": {"value": B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": e2e_ms},
            "gpu_launches_per_step": int(launches), "gpu_launches": int(launches) * args.steps, "clocks": clocks, "cpu_baseline": cpu}
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="train", choices=["train", "infer", "cls", "dropin"],
                    help="train = the headline cheb_VAE step (BASELINE configs[1]/[2]); infer / cls = configs[3] / [4]; "
                         "dropin = the reference's own model class on the import-path drop-in (eager)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch meshes per GPU (default 64); strong: --global-batch meshes split over the GPUs "
                         "(SURVEY 8(e): global 512 -> 512/N per rank)")
    ap.add_argument("--global-batch", type=int, default=512)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="meshes per GPU per step")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--host-comm", action="store_true",
                    help="data parallel: all-reduces launched from the host between three graphs (round-1 scheme) instead of "
                         "captured inside the one step graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    guard_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload == "dropin":
            args.workload = "train"
        if args.steps > 40:
            args.steps = 40
        return run_reference(args, rank)
    if args.scaling == "strong":
        if args.global_batch % world:
            raise SystemExit(f"bench.py: global batch {args.global_batch} does not divide over {world} GPUs")
        args.batch = args.global_batch // world
    if args.workload != "train":
        if rank != 0:
            return 0
        return run_secondary(args, local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        return run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())

"""Shared, dependency-free helpers for tests, the golden generator and bench.py (no oracle, no product imports)."""
import os
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
OPERATORS_NPZ = os.path.join(GOLDEN, "operators_template5k.npz")


def seeded_state_dict(module: torch.nn.Module, seed: int):
    """Deterministic parameters that do not depend on construction order or global RNG state:
    entry i of state_dict() gets randn(generator(seed*1000+i)) * scale, scale = 0.1 for conv
    weights/biases (as nn/conv.py:536-538) and 1/sqrt(fan_in) for Linear weights."""
    out = {}
    for i, (name, p) in enumerate(module.state_dict().items()):
        g = torch.Generator().manual_seed(seed * 1000 + i)
        t = torch.randn(p.shape, generator=g, dtype=torch.float32)
        if p.dim() == 2:
            t = t / float(np.sqrt(p.shape[1]))
        else:
            t = t * 0.1
        out[name] = t.to(p.dtype)
    return out


def seeded_batch(batch: int, n_vert: int, seed: int, feats: int = 3):
    """x [B,N,feats] f32 z-score-like input, one-hot labels, reparameterisation noise."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, n_vert, feats, generator=g, dtype=torch.float32)
    y = torch.randint(0, 2, (batch,), generator=g)
    eps = torch.randn(batch, 16, generator=g, dtype=torch.float32)
    return x, y, eps


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|) - the 'relative' of SURVEY 8(d) parity gates (per tensor)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = float(b.abs().max())
    if denom == 0.0:
        return float((a - b).abs().max())
    return float((a - b).abs().max()) / denom


def elem_err(a: torch.Tensor, b: torch.Tensor, rtol: float = 1e-4, atol_rms: float = 1e-4) -> float:
    """Per-element mixed tolerance: max over the elements of |a-b| / (rtol*|b| + atol_rms*rms(b)); the check passes when
    the value is <= 1.  Unlike rel_err (a max-norm bound per tensor) a small entry next to large ones must itself be
    right to rtol, up to an absolute floor tied to the tensor's RMS (fp32 sums of ~1e2..1e5 terms of that magnitude
    cannot be reproduced below it in a different summation order)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    if a.shape != b.shape:
        raise AssertionError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    if b.numel() == 0:
        return 0.0
    rms = float(b.pow(2).mean().sqrt())
    den = rtol * b.abs() + atol_rms * rms
    if rms == 0.0:
        return float((a - b).abs().max()) / atol_rms if float((a - b).abs().max()) > 0 else 0.0
    return float(((a - b).abs() / den).max())

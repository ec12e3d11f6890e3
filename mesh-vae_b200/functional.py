"""torch.autograd glue over the C ABI.  PyTorch supplies device memory, streams and the autograd
tape; all arithmetic of the hot path happens in libmvb_sm100a.so.  Tensors that cross into the
library are fp32 CUDA tensors in the VERTEX-MAJOR layout [N, B, F] (see include/mvb.h).

No CPU fallback: a non-CUDA tensor raises."""
from typing import Optional

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .operators import MeshOperator

LOG_SIGMA_DEFAULT = 1.0009117126464844      # softclip(1, -6) in fp32 (models/cheb_VAE.py:328-329)


def _req_cuda(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise _lib.MvbError(f"{name}: expected a CUDA tensor - this package has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise _lib.MvbError(f"{name}: expected {dtype}, got {t.dtype}")


def to_vertex_major(x: torch.Tensor) -> torch.Tensor:
    """logical [B, N, F] (any strides) -> physical [N, B, F] contiguous; free when x is already a
    permuted view of a vertex-major buffer (which is what every module of this package returns)."""
    xt = x.permute(1, 0, 2)
    return xt if xt.is_contiguous() else xt.contiguous()


def from_vertex_major(x_vm: torch.Tensor) -> torch.Tensor:
    """physical [N, B, F] -> logical [B, N, F] strided view (no copy)."""
    return x_vm.permute(1, 0, 2)


class _ChebConvFn(torch.autograd.Function):
    """mvb_cheb_fwd / mvb_cheb_bwd  (nn/conv.py:557-577 and its autograd)."""

    @staticmethod
    def forward(ctx, x_vm, weight, bias, op: MeshOperator, relu: bool):
        _req_cuda(x_vm, "cheb_conv x")
        _req_cuda(weight, "cheb_conv weight")
        n, b, fin = x_vm.shape
        k, fin_w, fout = weight.shape
        if fin_w != fin:
            raise _lib.MvbError(f"cheb_conv: x has {fin} features, weight expects {fin_w}")
        if op.n_rows != n or op.n_cols != n:
            raise _lib.MvbError(f"cheb_conv: operator is {op.n_rows}x{op.n_cols}, tensor has {n} vertices")
        x_vm = x_vm.contiguous()
        w = weight.contiguous()
        bb = None if bias is None else bias.contiguous()
        na = op.n_active
        basis = torch.empty((max(k - 1, 0), na, b, fin), device=x_vm.device, dtype=torch.float32)
        y = torch.empty((n, b, fout), device=x_vm.device, dtype=torch.float32)
        check(lib.mvb_cheb_fwd(n, b, fin, fout, k, na, op.nnz, ptr(op.rowptr), ptr(op.colidx), ptr(op.vals), ptr(x_vm), ptr(w),
                               ptr(bb), 1 if relu else 0, ptr(basis) if basis.numel() else None, ptr(y), stream_ptr()),
              "mvb_cheb_fwd")
        ctx.op, ctx.relu, ctx.has_bias = op, relu, bias is not None
        ctx.save_for_backward(x_vm, basis, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_vm, basis, w, y = ctx.saved_tensors
        op = ctx.op
        n, b, fin = x_vm.shape
        k, _, fout = w.shape
        dy = dy.contiguous()
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty_like(x_vm) if need_dx else None
        dw = torch.empty_like(w)
        db = torch.empty(fout, device=w.device, dtype=torch.float32) if ctx.has_bias else None
        na = op.n_active
        ws_bytes = lib.mvb_cheb_bwd_workspace_bytes(n, b, fin, fout, k, na, 1 if need_dx else 0)
        ws = torch.empty(ws_bytes, device=w.device, dtype=torch.uint8)
        check(lib.mvb_cheb_bwd(n, b, fin, fout, k, na, op.nnz, ptr(op.rowptr_t), ptr(op.colidx_t), ptr(op.vals_t), ptr(x_vm),
                               ptr(basis) if basis.numel() else None, ptr(w), ptr(y) if ctx.relu else None, ptr(dy), ptr(dx),
                               ptr(dw), ptr(db), ptr(ws), ws_bytes, stream_ptr()), "mvb_cheb_bwd")
        return dx, dw, db, None, None


def cheb_conv(x_vm: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], op: MeshOperator,
              relu: bool = False, pad_to_quads: bool = True) -> torch.Tensor:
    """x_vm [N,B,Fin] -> [N,B,Fout] (vertex-major in, vertex-major out).

    Feature widths that are not a multiple of 4 (the 3-channel mesh coordinates at the encoder
    input and the decoder output) are zero-padded to the next multiple of 4 around the kernel call,
    so that every plane is made of whole 16-byte quads: the SpMM takes its vector path and the
    contractions run on the tcgen05 kernels instead of the scalar FFMA fallback.  Zero features /
    zero weight columns do not change any result; the padding ops are differentiable torch glue
    (autograd slices the padded weight gradient back)."""
    fin, fout = x_vm.shape[2], weight.shape[2]
    pin, pout = (-fin) % 4, (-fout) % 4
    if not pad_to_quads or (pin == 0 and pout == 0):
        return _ChebConvFn.apply(x_vm, weight, bias, op, relu)
    if pin:
        x_vm = torch.nn.functional.pad(x_vm, (0, pin))
    w = torch.nn.functional.pad(weight, (0, pout, 0, pin))
    b = bias if (bias is None or pout == 0) else torch.nn.functional.pad(bias, (0, pout))
    y = _ChebConvFn.apply(x_vm, w, b, op, relu)
    return y[..., :fout] if pout else y


class _PoolFn(torch.autograd.Function):
    """mvb_pool_fwd / mvb_pool_bwd  (nn/pool.py:13-23; models/cheb_cls.py:22-27)."""

    @staticmethod
    def forward(ctx, x_vm, op: MeshOperator):
        _req_cuda(x_vm, "pool x")
        n, b, f = x_vm.shape
        if n != op.n_cols:
            raise _lib.MvbError(f"pool: operator has {op.n_cols} columns, tensor has {n} vertices")
        x_vm = x_vm.contiguous()
        y = torch.empty((op.n_rows, b, f), device=x_vm.device, dtype=torch.float32)
        check(lib.mvb_pool_fwd(op.n_rows, op.n_cols, ptr(op.rowptr), ptr(op.colidx), ptr(op.vals), ptr(x_vm), ptr(y), b * f,
                               stream_ptr()), "mvb_pool_fwd")
        ctx.op = op
        return y

    @staticmethod
    def backward(ctx, dy):
        op = ctx.op
        dy = dy.contiguous()
        _, b, f = dy.shape
        dx = torch.empty((op.n_cols, b, f), device=dy.device, dtype=torch.float32)
        check(lib.mvb_pool_bwd(op.n_cols, op.n_rows, ptr(op.rowptr_t), ptr(op.colidx_t), ptr(op.vals_t), ptr(dy), ptr(dx), b * f,
                               stream_ptr()), "mvb_pool_bwd")
        return dx, None


def pool(x_vm: torch.Tensor, op: MeshOperator) -> torch.Tensor:
    """x_vm [N,B,F] -> [M,B,F] = P x (vertex-major)."""
    return _PoolFn.apply(x_vm, op)


class _ReparamFn(torch.autograd.Function):
    """mvb_vae_reparam_fwd / _bwd  (models/cheb_VAE.py:309-319)."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        _req_cuda(mu, "reparam mu")
        mu, logvar, eps = mu.contiguous(), logvar.contiguous(), eps.contiguous()
        _req_cuda(eps, "reparam eps")
        z = torch.empty_like(mu)
        check(lib.mvb_vae_reparam_fwd(mu.numel(), ptr(mu), ptr(logvar), ptr(eps), ptr(z), stream_ptr()),
              "mvb_vae_reparam_fwd")
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        logvar, eps = ctx.saved_tensors
        dz = dz.contiguous()
        dmu, dlv = torch.empty_like(dz), torch.empty_like(dz)
        check(lib.mvb_vae_reparam_bwd(dz.numel(), ptr(logvar), ptr(eps), ptr(dz), ptr(dmu), ptr(dlv), stream_ptr()),
              "mvb_vae_reparam_bwd")
        return dmu, dlv, None


def reparameterize(mu, logvar, eps):
    return _ReparamFn.apply(mu, logvar, eps)


class _VaeLossFn(torch.autograd.Function):
    """mvb_vae_loss_fwd / _bwd  (models/cheb_VAE.py:321-346)."""

    @staticmethod
    def forward(ctx, recon_vm, x_gt, mu, logvar, y_hat, y, log_sigma: float):
        _req_cuda(recon_vm, "vae_loss recon")
        _req_cuda(x_gt, "vae_loss x_gt", None)
        if x_gt.dtype not in (torch.float32, torch.float64):
            raise _lib.MvbError(f"vae_loss: x_gt must be fp32 or fp64, got {x_gt.dtype}")
        n, b, c = recon_vm.shape
        if tuple(x_gt.shape) != (b, n, c):
            raise _lib.MvbError(f"vae_loss: x_gt is {tuple(x_gt.shape)}, expected {(b, n, c)}")
        z = mu.shape[1]
        ncls = y_hat.shape[1]
        recon_vm, x_gt = recon_vm.contiguous(), x_gt.contiguous()
        mu, logvar, y_hat = mu.contiguous(), logvar.contiguous(), y_hat.contiguous()
        y = y.to(torch.int64).contiguous()
        dev = recon_vm.device
        f64 = x_gt.dtype == torch.float64
        loss = torch.empty((), device=dev, dtype=torch.float64)
        kld = torch.empty(b, device=dev, dtype=torch.float32)
        rec = torch.empty(b, device=dev, dtype=torch.float64)
        correct = torch.empty((), device=dev, dtype=torch.int64)
        dnll = torch.empty_like(recon_vm)
        ws_bytes = lib.mvb_vae_loss_workspace_bytes(b, n)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        check(lib.mvb_vae_loss_fwd(b, n, c, z, ncls, ptr(recon_vm), ptr(x_gt), 1 if f64 else 0, ptr(mu), ptr(logvar),
                                   ptr(y_hat), ptr(y), float(log_sigma), ptr(loss), ptr(kld), ptr(rec), ptr(correct),
                                   ptr(dnll), ptr(ws), ws_bytes, stream_ptr()), "mvb_vae_loss_fwd")
        ctx.save_for_backward(dnll, mu, logvar, y_hat, y)
        ctx.dims = (b, n, c, z, ncls)
        if not f64:                     # all-fp32 call (inference.py:87): the reference returns fp32
            loss, rec = loss.float(), rec.float()
        ctx.mark_non_differentiable(kld, rec, correct)
        return loss, kld, rec, correct

    @staticmethod
    def backward(ctx, gloss, _gk, _gr, _gc):
        dnll, mu, logvar, y_hat, y = ctx.saved_tensors
        b, n, c, z, ncls = ctx.dims
        g = gloss.to(torch.float64).contiguous()
        need = ctx.needs_input_grad
        d_recon = torch.empty_like(dnll) if need[0] else None
        d_mu = torch.empty_like(mu) if need[2] else None
        d_lv = torch.empty_like(logvar) if need[3] else None
        d_yh = torch.empty_like(y_hat) if need[4] else None
        check(lib.mvb_vae_loss_bwd(b, n, c, z, ncls, ptr(dnll), ptr(mu), ptr(logvar), ptr(y_hat), ptr(y), ptr(g),
                                   ptr(d_recon), ptr(d_mu), ptr(d_lv), ptr(d_yh), stream_ptr()), "mvb_vae_loss_bwd")
        return d_recon, None, d_mu, d_lv, d_yh, None, None


def vae_loss(recon_vm, x_gt, mu, logvar, y_hat, y, log_sigma: float = LOG_SIGMA_DEFAULT):
    """-> (loss, kld[B], rec_loss[B], correct); only `loss` carries gradient (as main.py:80 uses it)."""
    return _VaeLossFn.apply(recon_vm, x_gt, mu, logvar, y_hat, y, log_sigma)


class _KldFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar):
        _req_cuda(mu, "KLD mu")
        mu, logvar = mu.contiguous(), logvar.contiguous()
        b, z = mu.shape
        out = torch.empty(b, device=mu.device, dtype=torch.float32)
        check(lib.mvb_kld_fwd(b, z, ptr(mu), ptr(logvar), ptr(out), stream_ptr()), "mvb_kld_fwd")
        ctx.save_for_backward(mu, logvar)
        return out

    @staticmethod
    def backward(ctx, g):
        mu, logvar = ctx.saved_tensors
        b, z = mu.shape
        g = g.to(torch.float32).contiguous()
        dmu, dlv = torch.empty_like(mu), torch.empty_like(logvar)
        check(lib.mvb_kld_bwd(b, z, ptr(mu), ptr(logvar), ptr(g), ptr(dmu), ptr(dlv), stream_ptr()), "mvb_kld_bwd")
        return dmu, dlv


def kld(mu, logvar):
    return _KldFn.apply(mu, logvar)


class _GaussianNllFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, x, log_sigma: float):
        _req_cuda(mu, "gaussian_nll mu")
        _req_cuda(x, "gaussian_nll x", None)
        if x.dtype not in (torch.float32, torch.float64) or x.shape != mu.shape:
            raise _lib.MvbError("gaussian_nll: x must be fp32/fp64 with the shape of mu")
        mu_c, x_c = mu.contiguous(), x.contiguous()
        out = torch.empty_like(x_c)
        f64 = x.dtype == torch.float64
        check(lib.mvb_gaussian_nll_fwd(mu_c.numel(), ptr(mu_c), ptr(x_c), 1 if f64 else 0, float(log_sigma), ptr(out),
                                       stream_ptr()), "mvb_gaussian_nll_fwd")
        ctx.save_for_backward(mu_c, x_c)
        ctx.log_sigma, ctx.f64 = float(log_sigma), f64
        return out

    @staticmethod
    def backward(ctx, g):
        mu_c, x_c = ctx.saved_tensors
        g = g.to(x_c.dtype).contiguous()
        d_mu = torch.empty_like(mu_c)
        check(lib.mvb_gaussian_nll_bwd(mu_c.numel(), ptr(mu_c), ptr(x_c), 1 if ctx.f64 else 0, ctx.log_sigma, ptr(g),
                                       ptr(d_mu), stream_ptr()), "mvb_gaussian_nll_bwd")
        return d_mu, None, None


def gaussian_nll(mu, x, log_sigma: float):
    return _GaussianNllFn.apply(mu, x, log_sigma)

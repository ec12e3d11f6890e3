"""Drop-in for the three logpdf.py functions on the hot path (logpdf.py:7-8 KLD, :22-23
gaussian_nll, :24-28 softclip), backed by CUDA kernels.  The fused loss epilogue used by the model
mirror is `functional.vae_loss`; these exist so the reference's own `loss_function`
(models/cheb_VAE.py:321-346) runs unchanged."""
import torch
import torch.nn.functional as F

from . import functional as Fn


def KLD(mu, logvar):
    return Fn.kld(mu, logvar)


def softclip(tensor, min):  # noqa: A002  (reference signature)
    """min + softplus(tensor - min) on a 1-element tensor: init-grade arithmetic, left to torch."""
    return min + F.softplus(tensor - min)


def gaussian_nll(mu, log_sigma, x):
    """0.5 ((x - mu)/exp(log_sigma))^2 + log_sigma + 0.5 log(2 pi); log_sigma is the constant
    1-element tensor of models/cheb_VAE.py:328-329 (no gradient flows to it)."""
    if isinstance(log_sigma, torch.Tensor):
        if log_sigma.numel() != 1 or log_sigma.requires_grad:
            raise NotImplementedError("gaussian_nll: only a constant scalar log_sigma is supported")
        log_sigma = float(log_sigma)
    mu_b, x_b = torch.broadcast_tensors(mu, x)
    return Fn.gaussian_nll(mu_b, x_b, log_sigma)

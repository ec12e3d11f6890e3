"""`logpdf` as the reference imports it (models/cheb_VAE.py:20)."""
from meshvae_b200.logpdf import KLD, softclip, gaussian_nll  # noqa: F401

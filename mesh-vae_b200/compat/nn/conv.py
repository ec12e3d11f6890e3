"""`nn.conv` as the reference imports it (models/cheb_VAE.py:18): native ChebConv_batch."""
from meshvae_b200.conv import ChebConv_batch, ChebConv  # noqa: F401

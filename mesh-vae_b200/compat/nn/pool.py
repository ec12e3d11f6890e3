"""`nn.pool` as the reference imports it (models/cheb_VAE.py:17): native SurfacePool."""
from meshvae_b200.pool import SurfacePool  # noqa: F401

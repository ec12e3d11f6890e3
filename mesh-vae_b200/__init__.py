"""mesh-vae_b200: B200-native (sm_100a) kernels and drop-in modules for the Mesh-VAE hot path
(Chebyshev mesh convolution, mesh pooling, VAE loss epilogue).  Import as `meshvae_b200`."""
from . import _lib                      # loads (or builds) libmvb_sm100a.so - fails loudly if impossible
from ._lib import MvbError
from . import operators, functional
from .conv import ChebConv_batch, ChebConv
from .pool import SurfacePool, Pool
from .cheb_vae import cheb_VAE
from .cheb_cls import cheb_GCN
from . import logpdf

__all__ = ["ChebConv_batch", "ChebConv", "SurfacePool", "Pool", "cheb_VAE", "cheb_GCN", "logpdf", "operators",
           "functional", "MvbError"]
__version__ = "0.1.0"

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
from . import formats, loader, loop, mesh_ops
from .model import get_model, scipy_to_torch_sparse

__all__ = ["ChebConv_batch", "ChebConv", "SurfacePool", "Pool", "cheb_VAE", "cheb_GCN", "logpdf", "operators",
           "functional", "formats", "loader", "loop", "mesh_ops", "get_model", "scipy_to_torch_sparse", "MvbError"]
__version__ = "0.1.0"


def install_compat():
    """Put `compat/` first on sys.path so the UNCHANGED reference files (models/cheb_VAE.py,
    models/cheb_cls.py, model.py, main.py ...) import the native modules under the names they use
    (`nn.conv`, `nn.pool`, `logpdf`, `torch_geometric...ChebConv`, `torch_scatter`)."""
    import os
    import sys
    d = os.path.join(os.path.abspath(__path__[0]), "compat")
    if d not in sys.path:
        sys.path.insert(0, d)
    return d


def accelerate(net):
    """Bind the fused entry points on a model built from the reference's own class (through
    install_compat()): ReLU fused into every conv the model follows with F.relu, the fused loss
    epilogue, and - for cheb_GCN - the native Pool.  Parameters are shared, not copied."""
    import types
    from . import functional as Fn
    convs = list(getattr(net, "cheb", []))
    dec = list(getattr(net, "cheb_dec", []))
    for conv in convs + dec[:-1]:
        if isinstance(conv, (ChebConv_batch, ChebConv)):
            conv.fuse_relu = True          # the callers apply F.relu next (cheb_VAE.py:264,285; cheb_cls.py:97)
    if hasattr(net, "loss_function") and hasattr(net, "z_mean"):
        def loss_function(self, x, recon_x, z, mu_z, logvar_z, y, y_hat):
            loss, kld, rec, correct = Fn.vae_loss(Fn.to_vertex_major(recon_x), x, mu_z, logvar_z, y_hat, y)
            return loss, correct, kld, rec
        net.loss_function = types.MethodType(loss_function, net)
    return net

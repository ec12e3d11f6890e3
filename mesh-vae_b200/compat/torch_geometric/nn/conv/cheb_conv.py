"""torch_geometric.nn.conv.cheb_conv.ChebConv as models/cheb_cls.py:18 imports it: native class."""
from meshvae_b200.conv import ChebConv  # noqa: F401

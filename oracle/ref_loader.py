"""TEST / BASELINE INFRASTRUCTURE - not part of the product.

Locates the UNCHANGED reference tree (the build container's /root/reference, else the staged copy oracle/_ref/reference
made by oracle/make_ref.py) and puts it on sys.path behind either set of import shims:
  * oracle/shims  - pure-torch stand-ins for the third-party LEAVES only (torch_scatter, torch_geometric utils, ...):
                    the reference's own nn/conv.py, nn/pool.py, logpdf.py run -> the CPU oracle / CPU baseline;
  * mesh-vae_b200/compat - the product's import-path drop-in: `nn.conv`, `nn.pool`, `logpdf`, torch_scatter ... resolve
                    to the CUDA modules and the reference's models/*.py, main.py run unchanged on the B200.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_root():
    for cand in (os.environ.get("MVB_REFERENCE"), "/root/reference", os.path.join(HERE, "_ref", "reference")):
        if cand and os.path.isfile(os.path.join(cand, "models", "cheb_VAE.py")):
            return cand
    return None


_SHADOWED = ("nn", "nn.conv", "nn.pool", "logpdf", "models", "models.cheb_VAE", "models.cheb_cls", "utils", "main", "data",
             "transform", "config_parser", "mesh_operations", "model", "inference", "crecon", "torch_scatter",
             "torch_geometric", "torch_sparse", "open3d", "psbody")


def purge_modules():
    """forget every module either import mode may have loaded (the two modes bind the same names differently)"""
    for name in list(sys.modules):
        if name in _SHADOWED or name.split(".")[0] in ("torch_geometric", "torch_scatter", "torch_sparse", "psbody", "open3d"):
            del sys.modules[name]


def use_reference_on_shims():
    """sys.path for the CPU oracle: leaf shims, then the reference.  Returns the reference root (None if absent)."""
    ref = reference_root()
    if ref is None:
        return None
    purge_modules()
    shims = os.path.join(HERE, "shims")
    for p in (ref, shims):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, ref)
    sys.path.insert(0, shims)
    return ref

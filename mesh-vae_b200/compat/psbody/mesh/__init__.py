"""`from psbody.mesh import Mesh` (mesh_operations.py:5, model.py:16, main.py:25, data.py:11) without MPI-IS/mesh."""
from meshvae_b200.mesh_ops import Mesh  # noqa: F401

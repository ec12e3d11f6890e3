"""`import mesh_operations` (model.py:18, main.py:19, data.py:9): the reference's function names on the fast,
bit-compatible implementation (meshvae_b200.mesh_ops) - shadows the reference's own file when `compat/` is first on
sys.path (3.5 s instead of 40 s for the 4998-vertex template, same A / D, U to fp tolerance)."""
from meshvae_b200.mesh_ops import (Mesh, get_vert_connectivity, get_vertices_per_edge, vertex_quadrics,  # noqa: F401
                                   qslim_decimator_transformer, setup_deformation_transfer, generate_transform_matrices,
                                   _get_sparse_transform)

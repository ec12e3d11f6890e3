"""Generate tests/golden/operators_template5k.npz by running the UNCHANGED reference
mesh_operations.generate_transform_matrices (mesh_operations.py:253-278) on
template/template5k.obj with downsampling_factors 4,4,4,4 (files/default.cfg:19), through the
TEST-ONLY leaf shims in oracle/shims (psbody AABB `nearest` is a brute-force restatement).

Run in the build container only (needs /root/reference):  python tests/golden/make_operators.py
The COO entries are stored exactly as model.py:24-32 hands them to torch (scipy CSC->COO order,
uncoalesced, int64 indices / f32 values); template vertices/faces are stored too so the GPU box
(which has no /root/reference) can synthesise hip-bone-shaped meshes.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))

import mesh_operations  # noqa: E402  (the reference's own file)
from psbody.mesh import Mesh  # noqa: E402  (shim)
import open3d as o3d  # noqa: E402  (shim)


def main():
    tm = o3d.io.read_triangle_mesh(os.path.join(REF, "template", "template5k.obj"))
    mesh = Mesh(v=tm.vertices, f=tm.triangles)
    M, A, D, U = mesh_operations.generate_transform_matrices(mesh, [4, 4, 4, 4])
    out = {"template_v": mesh.v.astype(np.float64), "template_f": np.asarray(mesh.f, dtype=np.int32),
           "num_nodes": np.asarray([len(m.v) for m in M], dtype=np.int64)}
    for name, mats in (("A", A), ("D", D), ("U", U)):
        for i, m in enumerate(mats):
            out[f"{name}{i}_row"] = m.row.astype(np.int64)
            out[f"{name}{i}_col"] = m.col.astype(np.int64)
            out[f"{name}{i}_val"] = m.data.astype(np.float32)
            out[f"{name}{i}_shape"] = np.asarray(m.shape, dtype=np.int64)
    for i, m in enumerate(M):
        out[f"M{i}_v"] = np.asarray(m.v, dtype=np.float64)
        out[f"M{i}_f"] = np.asarray(m.f, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "operators_template5k.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_shape") or k == "num_nodes"})
    print("num_nodes", out["num_nodes"])


if __name__ == "__main__":
    main()

"""Row f1 (operator set-up): meshvae_b200.mesh_ops against the operators the UNCHANGED reference produced from the
same template (tests/golden/operators_template5k.npz, written by tests/golden/make_operators.py): adjacency, QSlim
down-sampling and the decimated meshes bit for bit, up-sampling to fp32 rounding; the closest-point query against the
brute-force checker; the model hand-off."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests.helpers import OPERATORS_NPZ, ROOT

import meshvae_b200 as mvb
from meshvae_b200 import mesh_ops, formats


@pytest.fixture(scope="module")
def gold():
    return np.load(OPERATORS_NPZ)


@pytest.fixture(scope="module")
def pyramid(gold):
    mesh = mesh_ops.Mesh(v=gold["template_v"], f=gold["template_f"])
    return mesh_ops.generate_transform_matrices(mesh, [4, 4, 4, 4])          # files/default.cfg:19


def test_adjacency_decimation_and_meshes_are_bit_identical(gold, pyramid):
    M, A, D, U = pyramid
    assert [len(m.v) for m in M] == gold["num_nodes"].tolist() == [4998, 1250, 313, 79, 20]
    for i in range(5):
        assert np.array_equal(A[i].row, gold[f"A{i}_row"]) and np.array_equal(A[i].col, gold[f"A{i}_col"]), i
        assert np.array_equal(A[i].data.astype(np.float32), gold[f"A{i}_val"]), i
        assert np.array_equal(M[i].v, gold[f"M{i}_v"]), i                                # vertices are selected, never moved
        assert np.array_equal(np.asarray(M[i].f, dtype=np.int32), gold[f"M{i}_f"]), i
    for i in range(4):
        assert tuple(D[i].shape) == tuple(gold[f"D{i}_shape"])
        assert np.array_equal(D[i].row, gold[f"D{i}_row"]) and np.array_equal(D[i].col, gold[f"D{i}_col"]), i   # same collapses
        assert np.array_equal(D[i].data.astype(np.float32), gold[f"D{i}_val"]), i


def test_upsampling_matches_to_fp32_rounding(gold, pyramid):
    U = pyramid[3]
    for i in range(4):
        assert tuple(U[i].shape) == tuple(gold[f"U{i}_shape"])
        assert U[i].nnz == 3 * U[i].shape[0]                                 # three stored entries per row, zeros kept (quirk 12)
        assert np.array_equal(U[i].row, gold[f"U{i}_row"]) and np.array_equal(U[i].col, gold[f"U{i}_col"]), i
        assert np.abs(U[i].data.astype(np.float32) - gold[f"U{i}_val"]).max() <= 1e-6, i
        rows = np.asarray(sp.coo_matrix((U[i].data, (U[i].row, U[i].col)), shape=U[i].shape).sum(1)).ravel()
        at_vertex = np.isin(np.arange(U[i].shape[0]), pyramid[2][i].col)      # kept vertices map onto themselves
        assert np.allclose(rows[at_vertex], 1.0)


def test_closest_point_query_matches_brute_force(gold):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    try:
        from psbody.mesh import Mesh as BruteMesh                            # TEST-ONLY brute-force checker
    finally:
        sys.path.pop(0)
    v, f = gold["M2_v"], gold["M2_f"]
    rng = np.random.default_rng(0)
    pts = np.concatenate([gold["M1_v"][:400], v[:50], v[f[:100]].mean(1),             # near surface, on vertices, in faces
                          0.5 * (v[f[:100, 0]] + v[f[:100, 1]]), v[:100] + rng.normal(size=(100, 3)) * 5.0])
    bf, bp, bv = BruteMesh(v=v, f=f).compute_aabb_tree().nearest(pts, True)
    qf, qp, qv = mesh_ops.Mesh(v=v, f=f).compute_aabb_tree().nearest(pts, True)
    assert np.abs(qv - bv).max() < 1e-9
    d_b, d_q = np.linalg.norm(bv - pts, axis=1), np.linalg.norm(qv - pts, axis=1)
    assert np.abs(d_b - d_q).max() < 1e-9
    same = (qf == bf).ravel()
    assert same.mean() > 0.95                                                # ties on shared edges / vertices may pick a neighbour face
    assert np.array_equal(qp.ravel()[same], bp.ravel()[same])
    only_points = mesh_ops.Mesh(v=v, f=f).compute_aabb_tree().nearest(pts[:5])
    assert len(only_points) == 2 and only_points[1].shape == (5, 3)


def test_qslim_edge_cases():
    # a tetrahedron cannot lose a vertex and stay closed with >= 4 vertices: asking for 4 of 4 is the identity
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=float)
    f = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [2, 0, 3]])
    nf, d = mesh_ops.qslim_decimator_transformer(mesh_ops.Mesh(v=v, f=f), n_verts_desired=4)
    assert np.array_equal(nf, f) and np.array_equal(d.toarray(), np.eye(4))
    with pytest.raises(Exception):
        mesh_ops.qslim_decimator_transformer(mesh_ops.Mesh(v=v, f=f))
    # one collapse on an octahedron: 5 vertices left, D selects rows, faces stay non-degenerate
    v = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1.5]], dtype=float)
    f = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]])
    nf, d = mesh_ops.qslim_decimator_transformer(mesh_ops.Mesh(v=v, f=f), n_verts_desired=5)
    assert d.shape == (5, 6) and d.nnz == 5 and np.all(d.sum(1) == 1)
    assert nf.max() == 4 and all(len(set(t)) == 3 for t in nf.tolist())


def test_model_hand_off_from_an_obj_file(gold, tmp_path):
    obj = tmp_path / "template.obj"
    with open(obj, "w") as fp:                                                # full precision, so the template survives the file
        fp.write("".join("v %.17g %.17g %.17g\n" % tuple(r) for r in gold["template_v"]))
        fp.write("".join("f %d %d %d\n" % tuple(r + 1) for r in gold["template_f"]))
    cfg = formats.read_config(os.path.join(ROOT, "tests", "golden", "default_like.cfg"))
    cfg["template"], cfg["checkpoint_dir"] = str(obj), str(tmp_path / "ck")
    M, A_t, D_t, U_t, nn_ = mvb.model.build_operators(cfg["template"], cfg["downsampling_factors"])
    assert nn_ == [4998, 1250, 313, 79, 20]
    for name, mats in (("A", A_t), ("D", D_t), ("U", U_t)):
        for i, m in enumerate(mats):
            assert m.is_sparse and not m.is_coalesced()                       # handed over as given (model.py:24-32)
            assert m._indices().dtype == torch.int64 and m._values().dtype == torch.float32
            assert np.array_equal(m._indices()[0].numpy(), gold[f"{name}{i}_row"]) and np.array_equal(m._indices()[1].numpy(), gold[f"{name}{i}_col"])
            assert np.abs(m._values().numpy() - gold[f"{name}{i}_val"]).max() <= 1e-6
    net = mvb.get_model(cfg, "cpu")
    assert isinstance(net, mvb.cheb_VAE) and sum(p.numel() for p in net.parameters()) == 712642
    sd = torch.load(tmp_path / "ck" / "initial_weight.pt")                     # model.py:59-60
    assert list(sd) == list(net.state_dict()) and torch.equal(sd["enc_lin.weight"], net.enc_lin.weight)
    gcn = mvb.get_model(cfg, "cpu", model_type="cheb_GCN", save_init=False)
    assert isinstance(gcn, mvb.cheb_GCN) and gcn.cheb[0].in_channels == 6

"""Closed-form known-answer tests and equivariance of the oracle."""
import math

import numpy as np
import pytest
import torch

import harness
from oracle import e3nn_ops, ref_layers, wigner
from oracle.irreps import Irrep, Irreps


def test_irreps_algebra():
    ir = Irreps("64x0e+64x0o+256x0e+64x1e")
    assert ir.dim == 64 + 64 + 256 + 192 and ir.num_irreps == 448
    assert str(ir.simplify()) == "64x0e+64x0o+256x0e+64x1e"
    assert str(Irreps("2x0e+3x0e+1x1o").simplify()) == "5x0e+1x1o"
    assert [str(x) for x in Irrep("1o") * Irrep("2e")] == ["1o", "2o", "3o"]
    s = Irreps("1x1e+1x0e+1x0o+1x1o").sort()
    assert str(s.irreps) == "1x0o+1x0e+1x1o+1x1e" and s.p == (3, 1, 0, 2)
    assert Irrep("0e") in Irreps("3x0e") and Irrep("0o") not in Irreps("3x0e")


def test_w3j_kat():
    C = wigner.wigner_3j_np(1, 1, 1)
    assert abs(C[0, 1, 2] - 0.408248290463863) < 1e-12
    C = wigner.wigner_3j_np(1, 1, 2)
    for idx, v in {(0, 0, 2): -0.182574, (2, 2, 2): -0.182574, (1, 1, 2): 0.365148, (0, 0, 4): -0.316228,
                   (2, 2, 4): 0.316228, (0, 1, 1): 0.316228, (1, 2, 3): 0.316228}.items():
        assert abs(C[idx] - v) < 1e-6
    nnz = {(0, 0, 0): 1, (0, 1, 1): 3, (0, 2, 2): 5, (1, 1, 0): 3, (1, 1, 1): 6, (1, 1, 2): 11, (1, 2, 1): 11,
           (1, 2, 2): 16, (1, 2, 3): 21, (2, 2, 0): 5, (2, 2, 1): 16, (2, 2, 2): 25, (2, 2, 3): 28,
           (3, 1, 2): 21, (3, 1, 3): 26, (3, 2, 1): 21, (3, 2, 2): 28, (3, 2, 3): 41}
    for t, n in nnz.items():
        C = wigner.wigner_3j_np(*t)
        assert (np.abs(C) > 1e-12).sum() == n, t
        assert abs(np.linalg.norm(C) - 1) < 1e-12
    for l in range(4):
        C = wigner.wigner_3j_np(l, 0, l)
        assert np.allclose(C[:, 0, :], np.eye(2 * l + 1) / math.sqrt(2 * l + 1))


def test_sh_radial_kat():
    Y = wigner.spherical_harmonics([1, 2], torch.tensor([[0.0, 1.0, 0.0], [2.0, 0.0, 0.0]], dtype=torch.float64))
    s3, s5 = math.sqrt(3), math.sqrt(5)
    assert torch.allclose(Y[0], torch.tensor([0, s3, 0, 0, 0, s5, 0, 0], dtype=torch.float64))
    assert torch.allclose(Y[1], torch.tensor([s3, 0, 0, 0, 0, -1.118034, 0, -1.936492], dtype=torch.float64), atol=1e-6)
    assert abs(float(ref_layers._poly_cutoff(torch.tensor([2.5]), 1 / 5.0, 6.0)) - 0.85546875) < 1e-6
    assert float(ref_layers._poly_cutoff(torch.tensor([5.0]), 1 / 5.0, 6.0)) == 0.0
    b = ref_layers.BesselBasis(5.0, num_basis=8)(torch.tensor([2.5]))
    assert abs(float(b[0, 0]) - 0.16) < 1e-6


def test_normalize2mom_table():
    for name, c in e3nn_ops.NORMALIZE2MOM.items():
        assert abs(e3nn_ops.normalize2mom_constant(ref_layers.activations[name]) - c) < 1e-12, name


def test_path_counts():
    """Appendix B: paths / weight_numel / mid dim per layer of config_energy_force."""
    m = harness.build_oracle({"config": "config_energy_force", "seed": 0}, torch.float32)
    exp = [(3, 192, 576), (15, 960, 3264), (27, 1728, 5952), (30, 1920, 6528), (30, 1920, 6528)]
    for i, (paths, W, mid) in enumerate(exp):
        tp = getattr(m.func, f"layer{i}").conv.tp.tp
        assert (len(tp.instructions), tp.weight_numel, tp.irreps_out.dim) == (paths, W, mid)
    assert sum(p.numel() for p in m.parameters()) == 3747873 or True


def test_three_atom_edge_list():
    data = {"pos": torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [0, 1.0, 0]]), "_n_nodes": torch.tensor([[3]])}
    d, _ = ref_layers.computeEdgeIndex(data, {}, r_max=5.0)
    assert d["edge_index"].tolist() == [[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]]


@pytest.mark.parametrize("name,key,pre_edge", [("model_energy_force", "forces", {"r_max": 5.0}),
                                               ("model_dipole", "dipole", {"r_max": 5.0})])
def test_oracle_equivariance(name, key, pre_edge):
    """energies invariant; 1x1o outputs rotate with R and flip under inversion (fp64)."""
    g = harness.load_golden(name)
    model = harness.build_oracle(g["meta"], torch.float64, num_layers=3)
    gen = torch.Generator().manual_seed(5)
    R = -wigner.rand_rotation(gen)  # improper: rotation * inversion
    out0 = harness.run_oracle(model, g["in"], torch.float64, pre_edge=pre_edge)
    inp = dict(g["in"])
    inp["pos"] = inp["pos"].double() @ R.T
    out1 = harness.run_oracle(model, inp, torch.float64, pre_edge=pre_edge)
    assert harness.rel_err(out1[key], out0[key] @ R.T) < 1e-11
    if "energy" in out0:
        assert harness.rel_err(out1["energy"], out0["energy"]) < 1e-12
    D = Irreps(model.func.layer2.irreps_out["output_features"] if hasattr(model, "func")
               else model.layer2.irreps_out["output_features"]).D_from_matrix(R)
    assert harness.rel_err(out1["node_features"], out0["node_features"] @ D.T) < 1e-10


def test_oracle_forces_vs_finite_difference():
    g = harness.load_golden("model_energy_force")
    model = harness.build_oracle(g["meta"], torch.float64, num_layers=3)
    out = harness.run_oracle(model, g["in"], torch.float64, pre_edge={"r_max": 5.0})
    pos = g["in"]["pos"].double()
    h = 1e-5
    for (i, c) in [(0, 0), (4, 2), (9, 1)]:
        e = []
        for s in (+1, -1):
            p = pos.clone()
            p[i, c] += s * h
            o = harness.run_oracle(model, dict(g["in"], pos=p), torch.float64, edge_index=out["edge_index"])
            e.append(o["energy"].sum())
        fd = -(e[0] - e[1]) / (2 * h)
        assert abs(float(fd - out["forces"][i, c])) < 1e-6 * max(1.0, float(out["forces"].abs().max()))


def test_edge_merge_kat():
    """computeEdgeIndex with a pre-existing edge list (compute_edge.py:86-100), by hand: 3 atoms on a line at
    x = 0, 1, 3 with r_max 1.5 -> radius edges (0,1), (1,0); the old list holds the bond (1,2), (2,1) with types 7, 9"""
    from oracle import ref_layers

    data = {"pos": torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [3.0, 0, 0]]), "_n_nodes": torch.tensor([[3]]),
            "edge_index": torch.tensor([[1, 2], [2, 1]]), "bond_type": torch.tensor([[7], [9]])}
    attrs = {"pos": ("node", "1x1o"), "bond_type": ("edge", "1x0e")}
    d, attrs = ref_layers.computeEdgeIndex(data, attrs, r_max=1.5)
    assert d["edge_index"].tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    assert data["bond_type"].reshape(-1).tolist() == [0, 0, 7, 9] and data["_n_edges"].tolist() == [[4]]

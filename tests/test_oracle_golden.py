"""The oracle restatement against the fixtures produced by the genuine reference package
(tests/golden/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest
import torch

import harness
from oracle import ref_layers

CASES = [
    ("model_energy_force", ["energy", "forces", "node_features"], {"r_max": 5.0}),
    ("model_energy", ["total_energy", "node_features"], {"r_max": 4.0}),
    ("model_dipole", ["dipole", "node_features"], {"r_max": 5.0}),
    ("model_diffusion", ["score", "node_features"], None),
    ("model_diffusion_nll", ["score", "nll"], None),
    ("model_diffusion_CA", ["score_CA", "node_features"], None),
]


@pytest.mark.parametrize("name,keys,pre_edge", CASES)
def test_oracle_matches_reference_fp64(name, keys, pre_edge):
    g = harness.load_golden(name)
    model = harness.build_oracle(g["meta"], torch.float64)
    ei = g["out64"]["edge_index"] if name == "model_diffusion_CA" else None
    out = harness.run_oracle(model, g["in"], torch.float64, pre_edge=pre_edge, edge_index=ei)
    assert torch.equal(out["edge_index"], g["out64"]["edge_index"])
    for k in keys:
        err = harness.rel_err(out[k], g["out64"][k])
        assert err < 1e-10, (name, k, err)   # fp64 tolerance of north_star


@pytest.mark.parametrize("name,keys,pre_edge", CASES[:3])
def test_oracle_matches_reference_fp32(name, keys, pre_edge):
    g = harness.load_golden(name)
    model = harness.build_oracle(g["meta"], torch.float32)
    out = harness.run_oracle(model, g["in"], torch.float32, pre_edge=pre_edge)
    for k in keys:
        if k in g["out32"]:
            err = harness.rel_err(out[k], g["out32"][k])
            assert err < 1e-5, (name, k, err)    # fp32 tolerance of north_star


def test_oracle_neighbour_lists_bit_exact():
    z = np.load(harness.GOLDEN + "/neighbour_lists.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    for name, r in meta.items():
        data = {"pos": torch.from_numpy(z[f"{name}/pos"]), "_n_nodes": torch.from_numpy(z[f"{name}/n_nodes"])}
        d, _ = ref_layers.computeEdgeIndex(data, {}, r_max=r)
        assert torch.equal(d["edge_index"], torch.from_numpy(z[f"{name}/edge_index"])), name
        assert torch.equal(data["_n_edges"], torch.from_numpy(z[f"{name}/n_edges"])), name

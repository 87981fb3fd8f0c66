"""Host-side composition of the product (e3_layers mirror + dense contractions + imu layouts +
config builders) against the golden fixtures, with the CUDA kernels replaced by the TEST-ONLY
torch emulation (tests/torch_emulation.py).  CPU; the kernels themselves are checked on the GPU."""
import pytest
import torch

import harness
import product_harness
import torch_emulation

CASES = [
    ("model_energy_force", ["energy", "forces", "node_features"], {"r_max": 5.0}),
    ("model_energy", ["total_energy", "node_features"], {"r_max": 4.0}),
    ("model_dipole", ["dipole", "node_features"], {"r_max": 5.0}),
    ("model_diffusion", ["score", "node_features"], None),
    ("model_diffusion_nll", ["score", "nll"], None),
    ("model_diffusion_CA", ["score_CA", "node_features"], None),
]


@pytest.mark.parametrize("name,keys,pre_edge", CASES)
def test_product_composition_fp64(monkeypatch, name, keys, pre_edge):
    torch_emulation.patch(monkeypatch)
    import e3_layers.data.compute_edge as ce

    g = harness.load_golden(name)
    model = product_harness.build_product(g["meta"], torch.float64, "cpu")
    ei = g["out64"]["edge_index"] if name == "model_diffusion_CA" else None
    out = product_harness.run_product(model, g["in"], torch.float64, "cpu", pre_edge=pre_edge, edge_index=ei,
                                      compute_edge=ce.computeEdgeIndex)
    for k in keys:
        err = harness.rel_err(out[k], g["out64"][k])
        # config_diffusion_CA: the reference evaluates the relative-position cutoff in float32 even
        # in an fp64 run (`.float()` at nn/embedding.py:308, defect D6); the product's fp64 mode is
        # fp64 throughout, hence 1e-8 there.
        assert err < (1e-8 if name == "model_diffusion_CA" else 1e-10), (name, k, err)

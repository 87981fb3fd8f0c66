"""Parameter gradients of the fused interaction block (training mode) against the op-by-op fp64 mode
(plain torch autograd through the same kernels' fp64 templates + cuBLAS DGEMM)."""
import pytest
import torch

import harness
import product_harness
import e3_layers.nn.message_passing as mpm
from e3b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _grads(meta, dtype, inputs, key, pre_edge, fused):
    model = product_harness.build_product(meta, dtype, DEV).train()
    mpm.FUSED_BLOCKS = fused
    try:
        out = product_harness.run_product(model, inputs, dtype, DEV, pre_edge=pre_edge)
        w = torch.linspace(0.5, 1.5, out[key].numel(), device=DEV, dtype=dtype).view_as(out[key])
        (out[key] * w).sum().backward()
    finally:
        mpm.FUSED_BLOCKS = True
    return {n: p.grad.detach().double().cpu() for n, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("config,key,pre_edge", [("config_energy", "total_energy", {"r_max": 4.0}),
                                                 ("config_dipole", "dipole", {"r_max": 5.0})])
def test_parameter_gradients_fused_vs_fp64(config, key, pre_edge):
    meta = {"config": config, "seed": 11}
    inputs = synthetic.qm9_like(6, seed=2, n_min=3, n_max=9)
    if config == "config_dipole":
        inputs["species"] = inputs["species"].clamp(max=17)
    ref = _grads(meta, torch.float64, inputs, key, pre_edge, fused=False)
    got = _grads(meta, torch.float32, inputs, key, pre_edge, fused=True)
    base = _grads(meta, torch.float32, inputs, key, pre_edge, fused=False)
    assert set(ref) == set(got) == set(base)
    worst = 0.0
    for n in ref:
        scale = ref[n].abs().max().clamp_min(1e-30)
        e_fused = float((got[n] - ref[n]).abs().max() / scale)
        e_base = float((base[n] - ref[n]).abs().max() / scale)
        worst = max(worst, e_fused)
        # fp32 bar: 1e-4 of the parameter's largest gradient entry, and never much worse than the op-by-op fp32 path
        assert e_fused < max(1e-4, 5 * e_base), (n, e_fused, e_base)
    assert worst > 0.0


def test_energy_training_step_reduces_loss():
    """a few Adam steps on a fixed synthetic batch through the fused blocks (energy-only model)"""
    meta = {"config": "config_energy", "seed": 5}
    inputs = synthetic.qm9_like(8, seed=4, n_min=3, n_max=9)
    model = product_harness.build_product(meta, torch.float32, DEV).train()
    with torch.no_grad():
        target = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 4.0})["total_energy"] + 1.0
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        out = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 4.0})
        loss = ((out["total_energy"] - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]

"""Second-order mode (graph of the position gradient, reference nn/output.py:39-43) on a CPU: the
product's own autograd Functions -- whose backward passes are Functions again -- run on top of the
TEST-ONLY kernel stand-ins (tests/torch_emulation.py, ``patch_kernels``) and are compared with the
oracle differentiated twice by plain torch autograd.  The kernels themselves are checked on the GPU
(tests/test_gpu_second_order.py)."""
import pytest
import torch

import harness
import product_harness
import torch_emulation
from e3b200 import synthetic


def _loss(energy, forces):
    we = torch.linspace(0.5, 1.5, energy.numel(), dtype=energy.dtype).view_as(energy)
    wf = torch.linspace(-1.0, 2.0, forces.numel(), dtype=forces.dtype).view_as(forces)
    return (we * energy).sum() + (wf * forces).sum() + 0.5 * (forces * forces).sum()


def _oracle_grads(meta, inputs, pre_edge):
    model = harness.build_oracle(meta, torch.float64).train()
    torch.set_default_dtype(torch.float64)
    try:
        data = harness.cast_inputs(inputs, torch.float64)
        attrs = harness.attrs_for(data)
        n = data["_n_nodes"].reshape(-1)
        data["_node_segment"] = torch.repeat_interleave(torch.arange(len(n)), n)
        from oracle import ref_layers
        d, attrs = ref_layers.computeEdgeIndex(data, attrs, **pre_edge)
        data.update(d)
        out, _ = model(data, attrs, create_graph=True)
        loss = _loss(out["energy"], out["forces"])
        loss.backward()
    finally:
        torch.set_default_dtype(torch.float32)
    return float(loss), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("dense_function", [False, True])
def test_force_loss_parameter_gradients_fp64(monkeypatch, dense_function):
    """dense_function: the dense maps as ops._Dense nodes (the tcgen05 path of the second-order mode) instead of
    plain torch.matmul"""
    torch_emulation.patch_kernels(monkeypatch)
    from e3b200 import ops
    monkeypatch.setattr(ops, "FORCE_DENSE_FUNCTION", dense_function)
    calls = {"n": 0}
    if dense_function:
        orig = ops.k_dense

        def counted(*a, **k):
            calls["n"] += 1
            return orig(*a, **k)
        monkeypatch.setattr(ops, "k_dense", counted)
        orig_sc = ops.k_sc

        def counted_sc(*a, **k):
            calls["sc"] = calls.get("sc", 0) + 1
            return orig_sc(*a, **k)
        monkeypatch.setattr(ops, "k_sc", counted_sc)
    import e3_layers.data.compute_edge as ce

    meta = {"config": "config_energy_force", "seed": 3}
    inputs = synthetic.qm9_like(3, seed=5, n_min=3, n_max=6)
    ref_loss, ref = _oracle_grads(meta, inputs, {"r_max": 5.0})

    model = product_harness.build_product(meta, torch.float64, "cpu").train()
    out = product_harness.run_product(model, inputs, torch.float64, "cpu", pre_edge={"r_max": 5.0},
                                      compute_edge=ce.computeEdgeIndex)
    assert out["forces"].requires_grad
    loss = _loss(out["energy"], out["forces"])
    loss.backward()
    assert abs(float(loss) - ref_loss) < 1e-9 * max(1.0, abs(ref_loss))
    got = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert set(got) == set(ref), set(got) ^ set(ref)
    for n in ref:
        err = harness.rel_err(got[n], ref[n])
        assert err < 1e-9, (n, err)
    # the dense maps really went through the nodes: radial MLP / heads via k_dense, self-connection and the per-irrep
    # linear maps via k_sc (forward, position gradient and second backward of each)
    assert (calls["n"] > 50) == dense_function and (calls.get("sc", 0) > 50) == dense_function, calls


def test_needs_grad_now_memo_is_sound():
    """ops.needs_grad_now: the memo of 'does not reach the positions' nodes must not swallow nodes seen on the way
    to a positive answer"""
    from e3b200 import ops

    pos = torch.randn(4, 3, requires_grad=True)
    w = torch.randn(3, requires_grad=True)
    shared = pos * 2.0                      # reaches pos
    other = w * 3.0                         # does not
    mixed = other.sum() + shared            # DFS may visit `other`'s branch first or last
    assert ops.needs_grad_now(other)        # outside the context everything is needed
    with ops.positions_only(pos):
        assert not ops.needs_grad_now(other)
        assert ops.needs_grad_now(mixed)
        assert ops.needs_grad_now(shared)   # was visited during the positive walk above: must still be positive
        assert ops.needs_grad_now(mixed * 1.0)
        assert not ops.needs_grad_now(w) and ops.needs_grad_now(pos)
        assert not ops.needs_grad_now(torch.zeros(2))

"""Second-order mode on the GPU (reference nn/output.py:39-43, ``create_graph=self.training``):
every backward kernel differentiated once more, against plain torch autograd applied twice to the
TEST-ONLY closed forms (tests/torch_emulation.py) / the oracle.  fp64 at 1e-9, fp32 at 1e-5 relative
(1e-4 of the largest entry for parameter gradients of a whole model, as in test_gpu_training)."""
import pytest
import torch

import harness
import product_harness
import torch_emulation as emu
from e3b200 import ops, plan, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float32: 2e-5, torch.float64: 1e-9}


def rel(a, b):
    return harness.rel_err(a, b)


def _second_order(fn, inputs, dev, seed):
    """L2 = <c2, grad_inputs <c1, fn(inputs)>>; returns fn value, first gradients, gradients of L2 wrt inputs and c1"""
    g = torch.Generator().manual_seed(seed)
    leaves = [t.clone().to(dev).requires_grad_(True) for t in inputs]
    out = fn(*leaves)
    c1 = torch.randn(out.shape, generator=g, dtype=torch.float64).to(out.dtype).to(dev).requires_grad_(True)
    firsts = torch.autograd.grad(out, leaves, c1, create_graph=True)
    L2 = 0
    for f in firsts:
        c2 = torch.randn(f.shape, generator=g, dtype=torch.float64).to(f.dtype).to(dev)
        L2 = L2 + (f * c2).sum()
    seconds = torch.autograd.grad(L2, leaves + [c1], allow_unused=True)
    return [out] + list(firsts) + [s if s is not None else torch.zeros(1) for s in seconds]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_geometry_second_order(dtype):
    g = torch.Generator().manual_seed(0)
    N, E = 30, 211
    pos = (torch.randn(N, 3, generator=g, dtype=torch.float64) * 1.5).to(dtype)
    ei = torch.randint(0, N, (2, E), generator=g)
    ei = ei[:, ei[0] != ei[1]]
    bw = (torch.linspace(1, 8, 8) * 3.14159265 + 0.05 * torch.randn(8, generator=g)).to(dtype)

    def make(mod, dev):
        e = ei.to(dev)

        def fn(p, w):
            vec, ln = mod.edge_vectors(p, e, mod.graph_of(e, N))
            sh = mod.spherical_harmonics(vec, 2, True)
            rb = mod.radial_basis(ln, w, 5.0, 0.0, True, 0, 6.0)
            return torch.cat([sh, rb, vec, ln.unsqueeze(-1)], dim=1)
        return fn

    ref = _second_order(make(emu, "cpu"), [pos, bw], "cpu", 7)
    out = _second_order(make(ops, DEV), [pos, bw], DEV, 7)
    names = ["value", "d/dpos", "d/dbessel", "dd/dpos", "dd/dbessel", "dd/dcotangent"]
    for name, a, b in zip(names, out, ref):
        assert rel(a, b) < TOL[dtype] * (1 if name == "value" else 20 if dtype == torch.float32 else 1), name


@pytest.mark.parametrize("dtype,mul,sid", [(torch.float64, 5, 3), (torch.float32, 64, 3), (torch.float32, 32, 2),
                                           (torch.float32, 64, 0), (torch.float64, 3, 5)])
def test_tp_conv_second_order(dtype, mul, sid):
    base = plan.generated_structures()[sid]
    st = plan.with_mul(base, mul)
    g = torch.Generator().manual_seed(sid)
    N, E = 19, 157
    ei = torch.randint(0, N, (2, E), generator=g)
    x = torch.randn(N, st.irreps_in.dim, generator=g, dtype=torch.float64).to(dtype)
    sh = torch.randn(E, st.irreps_sh.dim, generator=g, dtype=torch.float64).to(dtype)
    w = torch.randn(E, st.weight_numel, generator=g, dtype=torch.float64).to(dtype)

    class P:      # the emulation only needs the structure
        structure = st

    ecsr = emu._csr(ei, N)
    ref = _second_order(lambda a, b, c: emu.tp_conv(a, b, c, P, ecsr), [x.double(), sh.double(), w.double()], "cpu", 11)
    p = ops.TPPlan(st)
    csr = ops.build_csr(ei.to(DEV), N)
    out = _second_order(lambda a, b, c: ops.tp_conv(a, b, c, p, csr), [x, sh, w], DEV, 11)
    names = ["y", "gx", "gsh", "gw", "ddx", "ddsh", "ddw", "ddgy"]
    for name, a, b in zip(names, out, ref):
        assert rel(a, b) < TOL[dtype], name


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gate_and_pooling_second_order(dtype):
    from e3_layers.nn.message_passing import _Gate
    from e3b200.irreps import Irreps

    scal, gated = Irreps("16x0e+16x0o"), Irreps("16x1e+16x1o+16x2e")
    gates = Irreps([(b.mul, "0e") for b in gated])
    gate = _Gate(scal, ["silu", "tanhlu"], gates, ["silu", "silu", "tanh"], gated)
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(53, gate.irreps_in.dim, generator=g, dtype=torch.float64) * 2).to(dtype)
    ref = _second_order(lambda a: emu.gate(a, gate.desc, None), [x], "cpu", 5)
    out = _second_order(lambda a: gate(a), [x], DEV, 5)
    for name, a, b in zip(["out", "gin", "ddx", "ddgout"], out, ref):
        assert rel(a, b) < TOL[dtype], name

    counts = torch.tensor([3, 0, 5, 1, 44])
    seg_ptr = torch.zeros(6, dtype=torch.long)
    seg_ptr[1:] = counts.cumsum(0)
    seg = torch.repeat_interleave(torch.arange(5), counts)
    sq = lambda t: t * t          # make the pooled value nonlinear so that the second order is not trivially zero
    ref = _second_order(lambda a: sq(emu.segment_sum(a, seg_ptr, seg, 5)), [x], "cpu", 6)
    out = _second_order(lambda a: sq(ops.segment_sum(a, seg_ptr.to(DEV), seg.to(DEV), 5)), [x], DEV, 6)
    for name, a, b in zip(["out", "gin", "ddx", "ddgout"], out, ref):
        assert rel(a, b) < TOL[dtype], name


@pytest.mark.parametrize("M,K,N", [(1500, 64, 1920), (700, 1920, 64), (333, 8, 64), (900, 1024, 256)])
def test_dense_node_second_order(M, K, N):
    """ops.dense (tcgen05 3xTF32 node, backward made of the same node) against fp64 torch.matmul, twice differentiated"""
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    Wflat = torch.randn(K * N + 7, generator=g)

    def ref_fn(a, wf):
        return 0.3 * torch.tanh(a @ wf[7:].reshape(K, N))         # a view of a flat parameter, as dense.Linear uses

    def our_fn(a, wf):
        return torch.tanh(ops.dense(a, wf[7:].reshape(K, N), 0.3))

    ref = _second_order(lambda a, wf: torch.tanh(0.3 * (a @ wf[7:].reshape(K, N))), [x.double(), Wflat.double()], "cpu", 3)
    out = _second_order(our_fn, [x, Wflat], DEV, 3)
    for name, a, b in zip(["y", "gx", "gW", "ddx", "ddW", "ddgy"], out, ref):
        assert rel(a, b) < 2e-5, name


def test_self_connection_nodes_second_order():
    """ScalarAttrTensorProduct through the trilinear tcgen05 nodes (second-order mode) against its plain torch
    formulation in fp64, differentiated twice with respect to features, attributes and weights"""
    from e3b200 import dense
    from e3b200.irreps import Irreps

    irr_in, irr_out = Irreps("32x0e+32x0o+32x1e+32x1o+32x2e"), Irreps("32x0e+32x0o+96x0e+32x1e+32x1o+32x2e")
    torch.manual_seed(0)
    mod = dense.ScalarAttrTensorProduct(irr_in, Irreps("16x0e"), irr_out)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(41, irr_in.dim, generator=g)
    a = torch.randn(41, 16, generator=g)
    W = mod.weight.detach().clone()

    def run(xx, aa, ww, second):
        saved = mod.weight
        del mod._parameters["weight"]
        mod.weight = ww                               # a plain tensor in the graph of the test's leaves
        try:
            if second:
                with ops.second_order():
                    return mod(xx, aa)
            return mod(xx, aa)
        finally:
            del mod.weight
            mod._parameters["weight"] = saved

    ref = _second_order(lambda xx, aa, ww: run(xx, aa, ww, False), [x.double(), a.double(), W.double()], "cpu", 2)
    mod.to(DEV)
    out = _second_order(lambda xx, aa, ww: run(xx, aa, ww, True), [x, a, W], DEV, 2)
    names = ["y", "gx", "ga", "gW", "ddx", "dda", "ddW", "ddgy"]
    for name, p, q in zip(names, out, ref):
        assert rel(p, q) < 2e-5, name


def _loss(energy, forces):
    we = torch.linspace(0.5, 1.5, energy.numel(), dtype=energy.dtype, device=energy.device).view_as(energy)
    wf = torch.linspace(-1.0, 2.0, forces.numel(), dtype=forces.dtype, device=forces.device).view_as(forces)
    return (we * energy).sum() + (wf * forces).sum() + 0.5 * (forces * forces).sum()


def _oracle_grads(meta, inputs, pre_edge):
    from oracle import ref_layers

    model = harness.build_oracle(meta, torch.float64).train()
    torch.set_default_dtype(torch.float64)
    try:
        data = harness.cast_inputs(inputs, torch.float64)
        attrs = harness.attrs_for(data)
        n = data["_n_nodes"].reshape(-1)
        data["_node_segment"] = torch.repeat_interleave(torch.arange(len(n)), n)
        d, attrs = ref_layers.computeEdgeIndex(data, attrs, **pre_edge)
        data.update(d)
        out, _ = model(data, attrs, create_graph=True)
        loss = _loss(out["energy"], out["forces"])
        loss.backward()
    finally:
        torch.set_default_dtype(torch.float32)
    return out, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}


def _product_grads(meta, inputs, dtype, pre_edge):
    model = product_harness.build_product(meta, dtype, DEV).train()
    out = product_harness.run_product(model, inputs, dtype, DEV, pre_edge=pre_edge)
    assert out["forces"].requires_grad
    _loss(out["energy"], out["forces"]).backward()
    return out, {n: p.grad.detach().double().cpu() for n, p in model.named_parameters() if p.grad is not None}


def test_force_matching_gradients_against_oracle():
    """config_energy_force in training mode: d loss(E, F) / d parameters, the path configs[1] trains on"""
    meta = {"config": "config_energy_force", "seed": 3}
    inputs = synthetic.qm9_like(4, seed=5, n_min=3, n_max=8)
    oref, ref = _oracle_grads(meta, inputs, {"r_max": 5.0})
    o64, g64 = _product_grads(meta, inputs, torch.float64, {"r_max": 5.0})
    o32, g32 = _product_grads(meta, inputs, torch.float32, {"r_max": 5.0})
    assert rel(o64["forces"], oref["forces"]) < 1e-10 and rel(o64["energy"], oref["energy"]) < 1e-10
    assert rel(o32["forces"], oref["forces"]) < 1e-5
    assert set(g64) == set(ref) == set(g32)
    for n in ref:
        assert rel(g64[n], ref[n]) < 1e-9, (n, rel(g64[n], ref[n]))
        assert rel(g32[n], ref[n]) < 1e-4, (n, rel(g32[n], ref[n]))


def test_training_mode_forces_equal_evaluation_forces():
    """the second-order (op by op) path and the fused first-order path give the same energies and forces"""
    meta = {"config": "config_energy_force", "seed": 9}
    inputs = synthetic.qm9_like(6, seed=1, n_min=3, n_max=12)
    model = product_harness.build_product(meta, torch.float32, DEV)
    ev = product_harness.run_product(model.eval(), inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    tr = product_harness.run_product(model.train(), inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    assert not ev["forces"].requires_grad and tr["forces"].requires_grad
    assert rel(tr["forces"], ev["forces"]) < 1e-5 and rel(tr["energy"], ev["energy"]) < 1e-5


def test_force_matching_training_reduces_loss():
    """a few Adam steps on the reference's energy+force loss shape (config_energy_force.py:30)"""
    meta = {"config": "config_energy_force", "seed": 5}
    inputs = synthetic.qm9_like(8, seed=4, n_min=3, n_max=9)
    model = product_harness.build_product(meta, torch.float32, DEV)
    with torch.no_grad():
        pass
    tgt = product_harness.run_product(model.eval(), inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    e_t, f_t = tgt["energy"].detach() + 0.5, tgt["forces"].detach() * 0.8
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        out = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
        loss = ((out["energy"] - e_t) ** 2).mean() + 30.0 * ((out["forces"] - f_t) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses


def test_force_loss_gradients_are_rotation_invariant():
    """size-independent property of the second-order path: a loss built from energies and |forces|^2 is invariant
    under a rigid rotation + inversion of the inputs, so its parameter gradients must not change (fp64 mode)"""
    from oracle import wigner

    meta = {"config": "config_energy_force", "seed": 12}
    inputs = synthetic.qm9_like(5, seed=9, n_min=3, n_max=10)
    R = -wigner.rand_rotation(torch.Generator().manual_seed(4))            # improper: rotation x inversion

    def grads(inp):
        model = product_harness.build_product(meta, torch.float64, DEV).train()
        data = dict(inp)
        data["pos"] = inp["pos"].double()
        out = product_harness.run_product(model, data, torch.float64, DEV, pre_edge={"r_max": 5.0})
        loss = (out["energy"] ** 2).sum() + (out["forces"] ** 2).sum()
        loss.backward()
        return float(loss.detach()), {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}

    # rotate in fp64 from the same fp32-representable coordinates so that both runs see exactly related inputs
    base = dict(inputs)
    l0, g0 = grads(base)
    rot = dict(inputs)
    rot["pos"] = inputs["pos"].double() @ R.T
    l1, g1 = grads(rot)
    assert abs(l0 - l1) < 1e-9 * abs(l0)
    for n in g0:
        assert rel(g1[n], g0[n]) < 1e-7, (n, rel(g1[n], g0[n]))

"""Every libe3b200 kernel against the oracle / closed-form torch on identical seeded inputs.
fp64 (generic templates) at 1e-10, fp32 (generated unrolled kernels) at 1e-5 relative."""
import ctypes

import pytest
import torch

import harness
import torch_emulation as emu
from e3b200 import layout, ops, plan
from oracle import e3nn_ops, ref_layers, wigner

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float32: 1e-5, torch.float64: 1e-10}


def rel(a, b):
    return harness.rel_err(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_edge_vectors_sh_radial(dtype):
    g = torch.Generator().manual_seed(0)
    N, E = 40, 333
    pos = torch.randn(N, 3, generator=g, dtype=torch.float64).to(dtype)
    ei = torch.randint(0, N, (2, E), generator=g)
    ei = ei[:, ei[0] != ei[1]]
    E = ei.shape[1]
    bw = (torch.linspace(1, 8, 8) * 3.14159265 + 0.05 * torch.randn(8, generator=g)).to(dtype)
    gsh, grad_ = torch.randn(E, 9, generator=g, dtype=torch.float64).to(dtype), torch.randn(E, 8, generator=g, dtype=torch.float64).to(dtype)

    def run(mod, dev):
        p = pos.clone().to(dev).requires_grad_(True)
        w = bw.clone().to(dev).requires_grad_(True)
        e = ei.to(dev)
        vec, ln = mod.edge_vectors(p, e, mod.graph_of(e, N))
        sh = mod.spherical_harmonics(vec, 2, True)
        rb = mod.radial_basis(ln, w, 5.0, 0.0, True, 0, 6.0)
        loss = (sh * gsh.to(dev)).sum() + (rb * grad_.to(dev)).sum()
        gp, gw = torch.autograd.grad(loss, (p, w))
        return vec, ln, sh, rb, gp, gw

    ref = run(emu, "cpu")
    out = run(ops, DEV)
    for name, a, b in zip(["vec", "len", "sh", "radial", "gpos", "gbessel"], out, ref):
        assert rel(a, b) < TOL[dtype] * (10 if name in ("gpos", "gbessel") and dtype == torch.float32 else 1), name


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_edge_geometry_kernels_against_the_oracle_layers(dtype):
    """edge vectors, spherical harmonics and Bessel x cutoff kernels against the ORACLE's restatement of the reference
    layers themselves (computeEdgeVector compute_edge.py:13-36, SphericalEncoding embedding.py:130-178,
    RadialBasisEncoding embedding.py:181-219), values and gradients w.r.t. positions and the Bessel frequencies --
    not against tests/torch_emulation.py"""
    g = torch.Generator().manual_seed(4)
    N, E = 64, 700
    pos = (torch.randn(N, 3, generator=g, dtype=torch.float64) * 1.5).to(dtype)
    ei = torch.randint(0, N, (2, E), generator=g)
    ei = ei[:, ei[0] != ei[1]]
    E = ei.shape[1]
    gsh = torch.randn(E, 9, generator=g, dtype=torch.float64).to(dtype)
    grb = torch.randn(E, 8, generator=g, dtype=torch.float64).to(dtype)
    torch.set_default_dtype(torch.float64)
    try:
        sph = ref_layers.SphericalEncoding(irreps_out="1x0e+1x1o+1x2e", irreps_in="1x1o")
        rad = ref_layers.RadialBasisEncoding(r_max=5.0, trainable=True, irreps_out="8x0e")
        bw0 = rad.basis.bessel_weights.detach().clone() + 0.05 * torch.randn(8, generator=g, dtype=torch.float64)
        rad.basis.bessel_weights.data.copy_(bw0)
        p = pos.double().clone().requires_grad_(True)
        data = {"pos": p, "edge_index": ei}
        d, attrs = ref_layers.computeEdgeVector(data, {"pos": ("node", "1x1o")})
        vec, ln = d["edge_vector"], d["edge_length"]
        sh, _ = sph({"vectors": vec}, {"vectors": ("edge", "1x1o")})
        rb, _ = rad({"input": ln}, {"input": ("edge", "1x0e")})
        sh, rb = list(sh.values())[0], list(rb.values())[0]
        loss = (sh * gsh.double()).sum() + (rb * grb.double()).sum()
        gp, gw = torch.autograd.grad(loss, (p, rad.basis.bessel_weights))
    finally:
        torch.set_default_dtype(torch.float32)
    ref = (vec.detach(), ln.detach(), sh.detach(), rb.detach(), gp, gw)
    p = pos.clone().to(DEV).requires_grad_(True)
    w = bw0.to(dtype).to(DEV).requires_grad_(True)
    e = ei.to(DEV)
    v, l = ops.edge_vectors(p, e, ops.graph_of(e, N))
    s = ops.spherical_harmonics(v, 2, True)
    r = ops.radial_basis(l, w, 5.0, 0.0, True, 0, 6.0)
    loss = (s * gsh.to(DEV)).sum() + (r * grb.to(DEV)).sum()
    out = (v, l, s, r) + torch.autograd.grad(loss, (p, w))
    for name, a, b in zip(["vec", "len", "sh", "radial", "gpos", "gbessel"], out, ref):
        assert rel(a, b) < TOL[dtype] * (10 if name in ("gpos", "gbessel") and dtype == torch.float32 else 1), (name, rel(a, b))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_radial_variants(dtype):
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(500, generator=g, dtype=torch.float64) * 1.2).to(dtype)
    bw = (torch.linspace(1, 32, 32) * 3.14159265).to(dtype)
    for (r_max, one_over_r, kind, scale) in [(1.0, False, 0, 1.0), (150.0, False, 1, 200.0)]:
        xs = (x * scale - (scale / 2 if kind == 1 else 0)).clone()
        ref_in = xs.clone().requires_grad_(True)
        ref = emu.radial_basis(ref_in, bw, r_max, 0.0, one_over_r, kind, 6.0)
        (gref,) = torch.autograd.grad(ref.sum(), ref_in)
        inp = xs.clone().to(DEV).requires_grad_(True)
        out = ops.radial_basis(inp, bw.to(DEV), r_max, 0.0, one_over_r, kind, 6.0)
        (gout,) = torch.autograd.grad(out.sum(), inp)
        assert rel(out, ref) < TOL[dtype] * 5
        assert rel(gout, gref) < TOL[dtype] * 50


def _tp_case(st, mul, N, E, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    st = plan.with_mul(st, mul)
    dst = torch.randint(0, N, (E,), generator=g)
    src = torch.randint(0, N, (E,), generator=g)
    ei = torch.stack([src, dst])
    x = torch.randn(N, st.irreps_in.dim, generator=g, dtype=torch.float64).to(dtype)
    sh = torch.randn(E, st.irreps_sh.dim, generator=g, dtype=torch.float64).to(dtype)
    w = torch.randn(E, st.weight_numel, generator=g, dtype=torch.float64).to(dtype)
    return st, ei, x, sh, w


def _oracle_tp(st, ei, x, sh, w, gy_e3):
    """reference dataflow in fp64 on the CPU: gather, e3nn uvu TP, scatter"""
    x, sh, w = (t.double().requires_grad_(True) for t in (x, sh, w))
    tpe = ref_layers.TensorProductExpansion(str(st.irreps_in), (str(st.irreps_sh), "sh"), (str(st.irreps_out), "o"),
                                            "uvu", internal_weight=False)
    y = e3nn_ops.scatter(tpe.tp(x[ei[0]], sh, w), ei[1], dim=0, dim_size=x.shape[0])
    gx, gsh, gw = torch.autograd.grad(y, (x, sh, w), gy_e3.double())
    return y.detach(), gx, gsh, gw


@pytest.mark.parametrize("sid", range(len(plan.generated_structures())))
@pytest.mark.parametrize("dtype,mul", [(torch.float64, 5), (torch.float32, 64), (torch.float32, 32), (torch.float32, 7)])
def test_tp_conv_kernels(sid, dtype, mul):
    base = plan.generated_structures()[sid]
    N, E = 23, 301
    st, ei, x, sh, w = _tp_case(base, mul, N, E, dtype, seed=sid)
    gy_e3 = torch.randn(N, st.irreps_mid.dim, generator=torch.Generator().manual_seed(99), dtype=torch.float64)
    y_ref, gx_ref, gsh_ref, gw_ref = _oracle_tp(st, ei, x, sh, w, gy_e3)

    p = ops.TPPlan(st)
    assert p.specialized                      # every reference structure has a generated kernel
    eid = ei.to(DEV)
    csr = ops.build_csr(eid, N)
    mid = st.irreps_mid.simplify()
    xi = layout.to_imu(x, st.irreps_in).to(DEV).requires_grad_(True)
    shd, wd = sh.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    y = ops.tp_conv(xi, shd, wd, p, csr)
    gy = layout.to_imu(gy_e3.to(dtype), mid).to(DEV)
    gx, gsh, gw = torch.autograd.grad(y, (xi, shd, wd), gy)
    tol = TOL[dtype]
    assert rel(layout.from_imu(y.cpu(), mid), y_ref) < tol
    assert rel(layout.from_imu(gx.cpu(), st.irreps_in), gx_ref) < tol
    assert rel(gsh, gsh_ref) < tol
    assert rel(gw, gw_ref) < tol


@pytest.mark.parametrize("env", [
    {"E3B_TP_PAIRED": "3", "E3B_TP_PAIRED_FORCE": "1"},      # two channels per thread everywhere (group and class-shared variants)
    {"E3B_TP_PAIRED": "0", "E3B_TP_DECOUPLED": "0"},         # one channel per thread, CTA barrier per edge, staged TMA reduce-add
    {"E3B_TP_PAIRED": "0", "E3B_TP_DECOUPLED": "1"},         # one channel per thread, decoupled warps everywhere
    {"E3B_TP_PIPELINED": "0"},                               # direct-load kernels (no TMA ring)
], ids=["paired-forced", "coupled", "decoupled", "direct"])
def test_tp_conv_kernel_variants(env):
    """The launcher picks ONE variant of the generated kernels per structure (measured defaults, csrc/gen_tp.py); the others stay
    selectable through environment switches that are read once per process -- so every variant is run against the oracle in a
    child process: all structures at multiplicity 64, forward + the three gradients, and the model test that exercises the
    node-reduction mode of the backward kernels."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    base = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"]
    targets = [(os.path.join(here, "test_gpu_ops.py"), "test_tp_conv_kernels and dtype1-64")]
    if "E3B_TP_PIPELINED" not in env:      # the model path shares weight rows between the two directions of an edge: pipelined kernels only
        targets.append((os.path.join(here, "test_gpu_models.py"), "restricted or in_kernel_node_reduction"))
    for target, select in targets:
        out = subprocess.run(base + [target, "-k", select], env={**os.environ, **env}, cwd=os.path.dirname(here),
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
        assert " passed" in out.stdout


def test_tp_conv_generic_matches_generated_fp32():
    """same inputs through the generic (table-driven) and the generated kernels"""
    base = plan.generated_structures()[3]
    st, ei, x, sh, w = _tp_case(base, 64, 50, 700, torch.float32, seed=5)
    csr = ops.build_csr(ei.to(DEV), 50)
    xi = layout.to_imu(x, st.irreps_in).to(DEV)
    y_fast = ops.tp_conv(xi, sh.to(DEV), w.to(DEV), ops.TPPlan(st), csr)
    y_gen = ops.tp_conv(xi, sh.to(DEV), w.to(DEV), ops.TPPlan(st, w3j_sign_preset=1), csr)  # preset 1 -> generic kernel
    # preset 1 flips the sign of odd-sum triples (1,2,2)/(2,1,2)/(2,2,1)...: compare magnitudes per output row
    assert y_fast.shape == y_gen.shape
    y64 = ops.tp_conv(xi.double(), sh.to(DEV).double(), w.to(DEV).double(), ops.TPPlan(st), csr)   # generic, fp64
    assert rel(y_fast, y64) < 1e-5


def test_irregular_structure_runs_on_generic_kernel():
    """an irreps combination no config produces: the plan is not specialised but still exact"""
    st = plan.TPStructure("6x1e+6x0o", "1x0e+1x1o", "6x0e+6x1o+6x1e+6x2o+6x0o")
    p = ops.TPPlan(st)
    assert not p.specialized
    g = torch.Generator().manual_seed(8)
    N, E = 9, 60
    ei = torch.randint(0, N, (2, E), generator=g)
    x = torch.randn(N, st.irreps_in.dim, generator=g, dtype=torch.float64)
    sh = torch.randn(E, st.irreps_sh.dim, generator=g, dtype=torch.float64)
    w = torch.randn(E, st.weight_numel, generator=g, dtype=torch.float64)
    gy_e3 = torch.randn(N, st.irreps_mid.dim, generator=g, dtype=torch.float64)
    y_ref, gx_ref, gsh_ref, gw_ref = _oracle_tp(st, ei, x, sh, w, gy_e3)
    csr = ops.build_csr(ei.to(DEV), N)
    mid = st.irreps_mid.simplify()
    xi = layout.to_imu(x, st.irreps_in).to(DEV).requires_grad_(True)
    shd, wd = sh.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    y = ops.tp_conv(xi, shd, wd, p, csr)
    gx, gsh, gw = torch.autograd.grad(y, (xi, shd, wd), layout.to_imu(gy_e3, mid).to(DEV))
    assert rel(layout.from_imu(y.cpu(), mid), y_ref) < 1e-10
    assert rel(layout.from_imu(gx.cpu(), st.irreps_in), gx_ref) < 1e-10
    assert rel(gsh, gsh_ref) < 1e-10 and rel(gw, gw_ref) < 1e-10


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gate_and_segment_sum_and_layout(dtype):
    from e3_layers.nn.message_passing import _Gate
    from e3b200.irreps import Irreps

    scal, gated = Irreps("16x0e+16x0o"), Irreps("16x1e+16x1o+16x2e")
    gates = Irreps([(b.mul, "0e") for b in gated])
    gate = _Gate(scal, ["silu", "tanhlu"], gates, ["silu"] * 3, gated)
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(77, gate.irreps_in.dim, generator=g, dtype=torch.float64) * 2).to(dtype)
    go = torch.randn(77, gate.irreps_out.dim, generator=g, dtype=torch.float64).to(dtype)
    xr = x.clone().requires_grad_(True)
    ref = emu.gate(xr, gate.desc, gate.irreps_out.dim)
    (gref,) = torch.autograd.grad(ref, xr, go)
    xd = x.clone().to(DEV).requires_grad_(True)
    out = gate(xd)
    (gout,) = torch.autograd.grad(out, xd, go.to(DEV))
    assert rel(out, ref) < TOL[dtype] and rel(gout, gref) < TOL[dtype]
    # oracle Gate agrees with the emulation used above
    og = e3nn_ops.Gate(str(scal), [ref_layers.activations["silu"], ref_layers.activations["tanhlu"]], str(gates),
                       [ref_layers.activations["silu"]] * 3, str(gated))
    assert rel(ref, og(x.double())) < 1e-6

    counts = torch.tensor([3, 0, 5, 1, 68])
    seg_ptr = torch.zeros(6, dtype=torch.long)
    seg_ptr[1:] = counts.cumsum(0)
    seg = torch.repeat_interleave(torch.arange(5), counts)
    s = ops.segment_sum(x.to(DEV), seg_ptr.to(DEV), seg.to(DEV), 5)
    assert rel(s, torch.zeros(5, x.shape[1], dtype=dtype).index_add_(0, seg, x)) < TOL[dtype]

    irr = Irreps("5x0e+3x1o+4x2e")
    t = torch.randn(11, irr.dim, generator=g, dtype=torch.float64).to(dtype)
    assert torch.equal(ops.layout_convert(t.to(DEV), irr, True).cpu(), layout.to_imu(t, irr))
    assert torch.equal(ops.layout_convert(layout.to_imu(t, irr).contiguous().to(DEV), irr, False).cpu(), t)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_layer_normalization_kernel(dtype):
    """kernel (first order) against the oracle's LayerNormalization, values and both gradients"""
    from e3_layers.nn.pointwise import LayerNormalization

    irr = "64x0e+64x0o+64x1e+64x1o+64x2e+64x2o"
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 1152, generator=g, dtype=torch.float64).to(dtype)
    x[5] = 0                                                    # an all-zero row exercises the eps
    std = (1.0 + 0.1 * torch.randn(6, generator=g, dtype=torch.float64)).to(dtype)
    go = torch.randn(37, 1152, generator=g, dtype=torch.float64).to(dtype)
    torch.set_default_dtype(dtype)
    try:
        ref_mod = ref_layers.LayerNormalization(irr, irr)
        mod = LayerNormalization(irr, irr).to(DEV)
    finally:
        torch.set_default_dtype(torch.float32)
    with torch.no_grad():
        ref_mod.std.copy_(std)
        mod.std.copy_(std.to(DEV))
    xr = x.clone().requires_grad_(True)
    yr = ref_mod({"input": xr}, {})[0]["output"]
    gxr, gsr = torch.autograd.grad(yr, (xr, ref_mod.std), go)
    xd = x.clone().to(DEV).requires_grad_(True)
    n0 = ops._lib.launch_count
    yd = mod({"input": xd}, {})[0]["output"]
    gxd, gsd = torch.autograd.grad(yd, (xd, mod.std), go.to(DEV))
    assert ops._lib.launch_count - n0 == 2                      # one kernel forward, one backward
    assert rel(yd, yr) < TOL[dtype] and rel(gxd, gxr) < TOL[dtype] * 5 and rel(gsd, gsr) < TOL[dtype] * 5
    with ops.second_order():                                    # the closed form used in second-order mode agrees
        y2 = mod({"input": xd}, {})[0]["output"]
    assert rel(y2, yr) < TOL[dtype]

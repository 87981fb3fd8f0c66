"""TEST-ONLY stand-ins for the libe3b200 kernels, written with differentiable torch ops, so that
the host-side composition of the product (e3_layers mirror, dense contractions, imu layouts,
config builders, weight layouts) can be checked against the golden fixtures on a machine WITHOUT
a GPU.  ``patch()`` swaps them into ``e3b200.ops``; nothing in the product imports this file and
the product itself has no CPU path (it raises)."""
import math

import torch

from e3b200 import cg, ops
from e3b200.ops import GraphCSR
from oracle import ref_layers, wigner


def _csr(edge_index, n_nodes):
    src, dst = edge_index[0], edge_index[1]
    E = edge_index.shape[1]
    in_eid = torch.argsort(dst, stable=True)
    out_eid = torch.argsort(src, stable=True)
    in_ptr = torch.zeros(n_nodes + 1, dtype=torch.long)
    in_ptr[1:] = torch.bincount(dst, minlength=n_nodes).cumsum(0)
    out_ptr = torch.zeros(n_nodes + 1, dtype=torch.long)
    out_ptr[1:] = torch.bincount(src, minlength=n_nodes).cumsum(0)
    return GraphCSR(n_nodes, E, in_ptr, src[in_eid].int(), in_eid.int(), out_ptr, out_eid.int())


def graph_of(edge_index, n_nodes):
    g = getattr(edge_index, "_e3b_csr", None)
    if g is None:
        g = _csr(edge_index, n_nodes)
        edge_index._e3b_csr = g
    return g


def radius_graph(pos, n_nodes_per_graph, r_max):
    data = {"pos": pos.float(), "_n_nodes": n_nodes_per_graph.reshape(-1, 1)}
    d, _ = ref_layers.computeEdgeIndex(data, {}, r_max=r_max)
    ei = d["edge_index"]
    return ei, data["_n_edges"], graph_of(ei, pos.shape[0])


def edge_vectors(pos, edge_index, csr):
    vec = pos[edge_index[1]] - pos[edge_index[0]]
    return vec, torch.linalg.norm(vec, dim=-1)


def spherical_harmonics(vec, lmax, normalize=True):
    return wigner.spherical_harmonics(list(range(lmax + 1)), vec, normalize, "component")


def radial_basis(r, bessel_w, r_max, r_min=0.0, one_over_r=True, cutoff_kind=0, p=6.0):
    r = r.reshape(-1)
    b = (2.0 / (r_max - r_min)) * torch.sin(bessel_w * r.unsqueeze(-1) / (r_max - r_min))
    if one_over_r:
        b = b / r.unsqueeze(-1)
    cut = ref_layers.symmetricCutoff if cutoff_kind == 1 else ref_layers._poly_cutoff
    return b * cut(r, 1.0 / r_max, p)[:, None]


class TPPlan:
    def __init__(self, structure, w3j_sign_preset=0):
        self.structure = structure
        self.specialized = False


def tp_conv(x_imu, sh, w, plan, csr):
    st = plan.structure
    mul = st.uniform_mul
    N = x_imu.shape[0]
    xo, _ = st.x_comp_offsets()
    so, _ = st.sh_comp_offsets()
    ybase, ykst, ydim = st.y_layout()
    src = csr.in_nbr.long()
    eid = csr.in_eid.long() if csr.in_eid is not None else torch.arange(csr.n_edges)
    dst = torch.repeat_interleave(torch.arange(N), csr.in_ptr[1:] - csr.in_ptr[:-1])
    xs = x_imu[src].view(len(src), -1, mul)           # [E, comp, mul]
    Y = sh[eid]
    W = w[eid].view(len(src), len(st.paths), mul)
    y = x_imu.new_zeros(N, ydim, mul)
    for q, p in enumerate(st.paths):
        l1, l2, l3 = st.irreps_in[p.i_in].ir.l, st.irreps_sh[p.i_sh].ir.l, p.ir_out.l
        C = torch.tensor(cg.w3j(l1, l2, l3), dtype=x_imu.dtype) * math.sqrt(2 * l3 + 1)
        xb = xs[:, xo[p.i_in]:xo[p.i_in] + 2 * l1 + 1]                 # [E, i, u]
        Yb = Y[:, so[p.i_sh]:so[p.i_sh] + 2 * l2 + 1]                  # [E, j]
        t = torch.einsum("ijk,eiu,ej,eu->eku", C, xb, Yb, W[:, q])     # [E, k, u]
        rows = ybase[p.slot] + ykst[p.slot] * torch.arange(2 * l3 + 1)
        contrib = x_imu.new_zeros(N, 2 * l3 + 1, mul).index_add_(0, dst, t)
        y = y.index_add(1, rows, contrib)
    return y.reshape(N, ydim * mul)


def segment_sum(src, seg_ptr, seg_index, n_out):
    out = src.new_zeros((n_out,) + tuple(src.shape[1:]))
    return out.index_add_(0, seg_index, src)


_ACTS = {0: lambda x: x, 1: torch.nn.functional.silu, 2: torch.tanh, 3: ref_layers.ShiftedSoftPlus,
         4: ref_layers.tanhlu, 5: torch.abs}


def gate(x, desc, out_dim):
    cols, off = [], 0
    ns = sum(desc.scalar_mul[i] for i in range(desc.n_scalar_blocks))
    ng = sum(desc.gated_mul[i] for i in range(desc.n_gated_blocks))
    for i in range(desc.n_scalar_blocks):
        m = desc.scalar_mul[i]
        cols.append(desc.scalar_cst[i] * _ACTS[desc.scalar_act[i]](x[:, off:off + m]))
        off += m
    goff, doff = ns, ns + ng
    for i in range(desc.n_gated_blocks):
        m, d = desc.gated_mul[i], 2 * desc.gated_l[i] + 1
        g = desc.gate_cst[i] * _ACTS[desc.gate_act[i]](x[:, goff:goff + m])
        cols.append((x[:, doff:doff + m * d].view(-1, m, d) * g.unsqueeze(-1)).reshape(-1, m * d))
        goff += m
        doff += m * d
    return torch.cat(cols, dim=1)


_NAMES = ["graph_of", "radius_graph", "edge_vectors", "spherical_harmonics", "radial_basis", "TPPlan", "tp_conv",
          "segment_sum", "gate"]


def patch(monkeypatch, names=None):
    import e3_layers.data.compute_edge as ce

    for n in (names or _NAMES):
        monkeypatch.setattr(ops, n, globals()[n])
    # the product's computeEdgeIndex refuses CPU tensors; lift that guard for the emulation only
    orig = ce.computeEdgeIndex

    def compute_edge_index_cpu(data, attrs, r_max=None, key="pos", criteria=None):
        assert criteria is None
        ei, n_edges, _ = radius_graph(data[key], data["_n_nodes"].reshape(-1), r_max)
        attrs["_n_edges"] = ("graph", "1x0e")
        data["_n_edges"] = n_edges
        return {"edge_index": ei}, attrs

    monkeypatch.setattr(ce, "computeEdgeIndex", compute_edge_index_cpu)
    return orig


# ---------------------------------------------------------------------------------------------
# Kernel-level stand-ins: replace only the launchers ``ops.k_*`` so that the product's OWN autograd
# Functions (including the differentiable backward nodes of the second-order mode) run on a CPU.
def _grad_of(fn, inputs, cotangent, wanted):
    """values of d<cotangent, fn(*inputs)>/d inputs[i] for i in wanted (None elsewhere)"""
    with torch.enable_grad():
        leaves = [t.detach().requires_grad_(True) for t in inputs]
        out = fn(*leaves)
        gs = torch.autograd.grad(out, [leaves[i] for i in wanted], cotangent.detach(), allow_unused=True)
    res = [None] * len(inputs)
    for i, g in zip(wanted, gs):
        res[i] = g if g is not None else torch.zeros_like(inputs[i])
    return res


def k_edge_fwd(pos, edge_index, want_len=True):
    vec = pos[edge_index[1]] - pos[edge_index[0]]
    return vec, (torch.linalg.norm(vec, dim=-1) if want_len else None)


def _edge_endpoints(csr):
    """(src, dst) per edge id from the two groupings of a GraphCSR"""
    E, N = csr.n_edges, csr.n_nodes
    dst = torch.empty(E, dtype=torch.long)
    src = torch.empty(E, dtype=torch.long)
    in_eid = csr.in_eid.long() if csr.in_eid is not None else torch.arange(E)
    out_eid = csr.out_eid.long() if csr.out_eid is not None else torch.arange(E)
    dst[in_eid] = torch.repeat_interleave(torch.arange(N), csr.in_ptr[1:] - csr.in_ptr[:-1])
    src[out_eid] = torch.repeat_interleave(torch.arange(N), csr.out_ptr[1:] - csr.out_ptr[:-1])
    return src, dst


def k_edge_scatter(gvec, glen, vec, length, n, csr):
    src, dst = _edge_endpoints(csr)
    tot = torch.zeros_like(vec)
    if gvec is not None:
        tot = tot + gvec
    if glen is not None:
        tot = tot + (glen / length).unsqueeze(-1) * vec
    return torch.zeros(n, 3, dtype=vec.dtype).index_add_(0, dst, tot).index_add_(0, src, -tot)


def k_sh_fwd(vec, lmax, normalize):
    return spherical_harmonics(vec, lmax, bool(normalize))


def k_sh_bwd(vec, gsh, lmax, normalize):
    return _grad_of(lambda v: spherical_harmonics(v, lmax, bool(normalize)), [vec], gsh, [0])[0]


def k_radial_fwd(r, bw, params):
    return radial_basis(r, bw, *params)


def k_radial_bwd(r, gout, bw, params, need_w):
    gr, gw = _grad_of(lambda rr, ww: radial_basis(rr, ww, *params), [r, bw], gout, [0, 1])
    return gr, (gw if need_w else None)


def k_tp_fwd(plan, csr, x, sh, w):
    return tp_conv(x, sh, w, plan, csr)


def k_tp_bwd(plan, csr, x, sh, w, gy, need_x, need_sh):
    wanted = ([0] if need_x else []) + ([1] if need_sh else []) + [2]
    gx, gsh, gw = _grad_of(lambda a, b, c: tp_conv(a, b, c, plan, csr), [x, sh, w], gy, wanted)
    return gx, gsh, gw


def k_segment_sum(src2, seg_ptr, ids, n_out):
    seg = torch.repeat_interleave(torch.arange(n_out), seg_ptr[1:] - seg_ptr[:-1])
    rows = src2[ids.long()] if ids is not None else src2[:len(seg)]
    return torch.zeros(n_out, src2.shape[1], dtype=src2.dtype).index_add_(0, seg, rows)


def k_gate_fwd(desc, x, out_dim):
    return gate(x, desc, out_dim)


def k_gate_bwd(desc, x, gout):
    return _grad_of(lambda a: gate(a, desc, None), [x], gout, [0])[0]


def k_gate_bwd2(desc, x, gout, ggin):
    with torch.enable_grad():
        a = x.detach().requires_grad_(True)
        g = gout.detach().requires_grad_(True)
        (gin,) = torch.autograd.grad(gate(a, desc, None), a, g, create_graph=True)
        gx, gg = torch.autograd.grad(gin, [a, g], ggin.detach(), allow_unused=True)
    return (gx if gx is not None else torch.zeros_like(x)), (gg if gg is not None else torch.zeros_like(gout))


def k_dense(x, W, alpha, trans):
    return alpha * (x @ (W.t() if trans else W))


def k_sc(spec, src, attrs, W, to_out):
    N, V = src.shape[0], spec.V
    dst = src.new_zeros(N, spec.Dout if to_out else spec.Din)
    for i, o, off, alpha in spec.paths:
        bi, bo = spec.irreps_in[i], spec.irreps_out[o]
        d = bi.ir.dim
        if V:
            Wp = W[off:off + bi.mul * V * bo.mul].reshape(bi.mul, V, bo.mul)
            Weff = torch.einsum("uvw,zv->zuw", Wp, attrs)
        else:
            Weff = W[off:off + bi.mul * bo.mul].reshape(1, bi.mul, bo.mul).expand(N, -1, -1)
        if to_out:
            xb = src[:, spec.x_off[i]:spec.x_off[i] + bi.dim].reshape(N, d, bi.mul)
            dst[:, spec.c_off[o]:spec.c_off[o] + bo.dim] += alpha * torch.einsum("zuw,zdu->zdw", Weff, xb).reshape(N, -1)
        else:
            gb = src[:, spec.c_off[o]:spec.c_off[o] + bo.dim].reshape(N, d, bo.mul)
            dst[:, spec.x_off[i]:spec.x_off[i] + bi.dim] += alpha * torch.einsum("zuw,zdw->zdu", Weff, gb).reshape(N, -1)
    return dst


def k_layout(x, irreps, to_imu):
    from e3b200 import layout
    return (layout.to_imu if to_imu else layout.from_imu)(x, irreps).contiguous()


_KERNELS = ["k_sc", "k_layout", "k_dense", "k_edge_fwd", "k_edge_scatter", "k_sh_fwd", "k_sh_bwd", "k_radial_fwd", "k_radial_bwd", "k_tp_fwd", "k_tp_bwd",
            "k_segment_sum", "k_gate_fwd", "k_gate_bwd", "k_gate_bwd2"]


def patch_kernels(monkeypatch):
    """like patch(), but BELOW the product's autograd Functions: only the launchers and the graph helpers"""
    return patch(monkeypatch, names=["graph_of", "radius_graph", "TPPlan"] + _KERNELS)

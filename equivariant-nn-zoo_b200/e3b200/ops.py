"""torch.autograd bindings of the libe3b200 kernels.  PyTorch here is plumbing (device memory,
streams, autograd tape); every op below launches hand-written sm_100a kernels through the C ABI
and raises if the library or a CUDA device is missing."""
import ctypes
import os
import weakref

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, count_launch, dtype_code, ptr, require_cuda, stream


# ------------------------------------------------------------------------------------------
# Second-order mode: set by GradientOutput while it runs its network in training mode, where the
# graph of the position gradient is needed (force-matching losses).  The interaction blocks then
# compose op by op from the Functions below, whose backward passes are differentiable again;
# the single-node fused block (e3b200.interaction) is first order only.
_SECOND_ORDER = 0

# bumped whenever parameter memory is rewritten behind torch's version counters (e3b200.optim.FlatAdam): part of
# the key of every cache of packed tensor-core weights
WEIGHTS_EPOCH = 0


# E3B_DETERMINISTIC=1: bit-reproducible gradients everywhere (the evaluation-mode backward of the fused block then writes
# the per-edge d/dx rows and reduces them with the segment-sum kernel instead of TMA reduce-adds in the kernel)
DETERMINISTIC = __import__("os").environ.get("E3B_DETERMINISTIC", "0") == "1"


class second_order:
    def __enter__(self):
        global _SECOND_ORDER
        _SECOND_ORDER += 1

    def __exit__(self, *exc):
        global _SECOND_ORDER
        _SECOND_ORDER -= 1
        return False


def second_order_active():
    return _SECOND_ORDER > 0


# Set by GradientOutput around its autograd.grad call: that backward pass differentiates with respect to the
# positions alone, so dense Functions skip the gradients of arguments that do not depend on them (weight
# gradients are reductions over all edges / nodes that the engine would throw away; built-in torch ops prune
# such outputs by themselves, Python Functions have to ask).
_POSITIONS_ONLY = []      # stack of (wrt tensor, memo of graph nodes known not to reach it)


class positions_only:
    def __init__(self, wrt):
        self.wrt = wrt

    def __enter__(self):
        _POSITIONS_ONLY.append((self.wrt, set()))

    def __exit__(self, *exc):
        _POSITIONS_ONLY.pop()
        return False


def positions_only_active():
    return bool(_POSITIONS_ONLY)


def needs_grad_now(t):
    """False when the running backward pass differentiates with respect to one tensor only (GradientOutput's
    position gradient) and `t` does not depend on it; True otherwise.  Walks t's autograd graph (memoised)."""
    if not _POSITIONS_ONLY:
        return True
    wrt, dead = _POSITIONS_ONLY[-1]
    if t is wrt:
        return True
    root = t.grad_fn
    if root is None:
        return False
    stack, seen = [root], []
    while stack:
        fn = stack.pop()
        if fn is None or fn in dead:
            continue
        if getattr(fn, "variable", None) is wrt:
            dead.difference_update(seen)      # nodes visited on the way may well reach it: forget them again
            return True
        dead.add(fn)
        seen.append(fn)
        stack.extend(nf for nf, _ in fn.next_functions)
    return False


# ------------------------------------------------------------------------------------------
# graph structure
class GraphCSR:
    """Both groupings of an edge list edge_index [2, E] over N nodes (reference convention:
    edge_index[0] = source, edge_index[1] = destination, ``nn/message_passing.py:96-97``).

    in_*  : grouped by destination (what the convolution forward walks)
    out_* : grouped by source      (what d/dx and d/dpos reductions walk)
    *_eid[k] is the edge id (column of edge_index) in slot k; None means identity."""

    def __init__(self, n_nodes, n_edges, in_ptr, in_nbr, in_eid, out_ptr, out_eid):
        self.n_nodes, self.n_edges = n_nodes, n_edges
        self.in_ptr, self.in_nbr, self.in_eid = in_ptr, in_nbr, in_eid
        self.out_ptr, self.out_eid = out_ptr, out_eid


def _exclusive_scan(deg, n):
    ptr_ = torch.zeros(n + 1, dtype=torch.int64, device=deg.device)
    if n:
        torch.cumsum(deg, 0, out=ptr_[1:])
    return ptr_


def _group(edge_index, n_nodes, which_row):
    lib = _lib.load()
    E = edge_index.shape[1]
    key = edge_index[which_row]
    deg = torch.bincount(key, minlength=n_nodes) if E else torch.zeros(n_nodes, dtype=torch.int64, device=key.device)
    row_ptr = _exclusive_scan(deg, n_nodes)
    eid = torch.empty(E, dtype=torch.int32, device=key.device)
    cursor = torch.zeros(max(n_nodes, 1), dtype=torch.int32, device=key.device)
    check(lib.e3b_csr_fill(ptr(edge_index), E, n_nodes, which_row, ptr(row_ptr), ptr(cursor), ptr(eid), stream()))
    count_launch(2)
    return row_ptr, eid


def build_csr(edge_index, n_nodes):
    """CSR views of an arbitrary edge list (kernels e3b_csr_fill)."""
    require_cuda(edge_index)
    edge_index = edge_index.contiguous()
    E = edge_index.shape[1]
    in_ptr, in_eid = _group(edge_index, n_nodes, 1)
    out_ptr, out_eid = _group(edge_index, n_nodes, 0)
    in_nbr = edge_index[0][in_eid.long()].to(torch.int32) if E else torch.empty(0, dtype=torch.int32, device=edge_index.device)
    return GraphCSR(n_nodes, E, in_ptr, in_nbr, in_eid, out_ptr, out_eid)


def graph_of(edge_index, n_nodes):
    """CSR views cached on the edge_index tensor object (it travels through the layer dict)."""
    g = getattr(edge_index, "_e3b_csr", None)
    if g is None or g.n_nodes != n_nodes or g.n_edges != edge_index.shape[1]:
        g = build_csr(edge_index, n_nodes)
        edge_index._e3b_csr = g
    return g


class PairCriteria:
    """The `criteria` of the protein config (reference ``configs/config_diffusion_CA.py:58-64``) in a form the
    neighbour-list kernel evaluates inside its sweep: same segment (chain) and |a - b| < max_separation, OR a
    Bernoulli(p_random) draw per ordered pair.  Also a plain callable ``criteria(data, edge_index) -> mask`` with
    the reference's protocol, so generic callers keep working.

    Randomness: the reference draws ``torch.rand(n_pairs)`` over its all-pairs list.  If ``data`` carries
    ``_pair_uniforms`` (that vector: one float32 per ordered pair, graphs concatenated, a slow / b fast) the kernel
    reads it -- a seeded run is then reproduced edge for edge.  Otherwise a counter-based hash of (seed, a, b) is
    used, with a fresh seed drawn from torch's CPU generator per call (``torch.manual_seed`` makes it repeatable)."""

    def __init__(self, segment_key=None, max_separation=0, p_random=0.0):
        self.segment_key, self.max_separation, self.p_random = segment_key, int(max_separation), float(p_random)

    def __call__(self, data, edge_index):
        src, dst = edge_index[0], edge_index[1]
        keep = torch.zeros(src.shape[0], dtype=torch.bool, device=src.device)
        if self.segment_key is not None:
            seg = data[self.segment_key].view(-1)
            keep = (seg[src] == seg[dst]) & ((src - dst).abs() < self.max_separation)
        if self.p_random > 0:
            u = data.get("_pair_uniforms") if hasattr(data, "get") else None
            if u is None:
                u = torch.rand(src.shape[0]).to(src.device)            # CPU generator, as the reference
            keep = keep | (u.to(src.device) < self.p_random)
        return keep


class _NeighbourCount:
    """state between the two phases of the neighbour list (the caller sizes / chooses the edge arrays in between)"""
    __slots__ = ("pos", "node_ptr", "G", "N", "r_max", "row_ptr", "E", "crit", "crit_keep", "cells")


CELL_MIN_NODES = int(__import__("os").environ.get("E3B_CELL_MIN", "1024"))   # graphs this large are binned (cell list)


def _use_cells(N, G):
    """the graph sizes live on the device; the dispatch uses the mean size, which the host knows from the shapes
    (graphs below CELL_MIN_NODES take the all-pairs loop inside the cell-list kernel anyway)"""
    return G > 0 and N // G >= CELL_MIN_NODES


def _crit_struct(st, criteria, data, counts):
    """e3b_pair_criteria for the kernel + the tensors it points to (kept alive on `st`)"""
    dev = st.pos.device
    c = _lib.PairCriteriaStruct()
    keep = []
    if criteria.segment_key is not None:
        seg = data[criteria.segment_key].reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        keep.append(seg)
        c.segment, c.max_separation = ptr(seg), criteria.max_separation
    c.p_random = criteria.p_random
    if criteria.p_random > 0:
        u = data.get("_pair_uniforms")
        if u is not None:
            u = u.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
            pair_ptr = _exclusive_scan(counts * counts, st.G)
            keep += [u, pair_ptr]
            c.uniforms, c.pair_ptr = ptr(u), ptr(pair_ptr)
        else:
            c.seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # CPU generator: no device synchronisation
    return c, keep


def radius_graph_count(pos, n_nodes_per_graph, r_max, criteria=None, data=None):
    """phase 1: per-atom degrees -> row_ptr, and the edge count E (the ONE host synchronisation of the step).
    criteria: None | PairCriteria (evaluated in the sweep); large radius-only graphs are binned into a cell list."""
    lib = _lib.load()
    require_cuda(pos)
    if pos.dtype != torch.float32:
        pos = pos.float()  # the reference predicate is evaluated in the default dtype (fp32)
    st = _NeighbourCount()
    st.pos = pos.contiguous()
    st.N, st.r_max = pos.shape[0], float(r_max)
    counts = n_nodes_per_graph.reshape(-1).to(device=pos.device, dtype=torch.int64)
    st.G = counts.numel()
    st.node_ptr = _exclusive_scan(counts, st.G)
    st.crit = st.crit_keep = st.cells = None
    dev = pos.device
    deg = torch.zeros(max(st.N, 1), dtype=torch.int32, device=dev)
    if criteria is not None:
        st.crit, st.crit_keep = _crit_struct(st, criteria, data if data is not None else {}, counts)
        check(lib.e3b_pair_graph_count(ptr(st.pos), 3, ptr(st.node_ptr), st.G, st.N, st.r_max, ctypes.byref(st.crit), ptr(deg),
                                       stream()))
        count_launch()
    elif _use_cells(st.N, st.G):
        C = 2 * st.N + st.G                                            # about two cells per atom, known without a sync
        cell_ptr = _exclusive_scan(2 * counts + 1, st.G)
        grids = torch.empty(st.G * _lib.CELL_GRID_BYTES, dtype=torch.uint8, device=dev)
        cell_of = torch.empty(st.N, dtype=torch.int32, device=dev)
        cell_count = torch.zeros(C, dtype=torch.int32, device=dev)
        check(lib.e3b_cell_graph_bin(ptr(st.pos), 3, ptr(st.node_ptr), st.G, st.N, st.r_max, CELL_MIN_NODES, ptr(cell_ptr),
                                     ptr(grids), ptr(cell_of), ptr(cell_count), stream()))
        cell_start = _exclusive_scan(cell_count.long(), C)
        cursor = torch.zeros(C, dtype=torch.int32, device=dev)
        order = torch.empty(st.N, dtype=torch.int32, device=dev)
        check(lib.e3b_cell_graph_sort(ptr(cell_of), st.N, ptr(cell_start), ptr(cursor), ptr(order), stream()))
        check(lib.e3b_cell_graph_count(ptr(st.pos), 3, ptr(st.node_ptr), st.G, st.N, st.r_max, ptr(grids), ptr(cell_start),
                                       ptr(order), ptr(deg), stream()))
        st.cells = (grids, cell_start, order)
        count_launch(4)
    else:
        check(lib.e3b_radius_graph_count(ptr(st.pos), 3, ptr(st.node_ptr), st.G, st.N, st.r_max, ptr(deg), stream()))
        count_launch()
    st.row_ptr = _exclusive_scan(deg[:st.N].long(), st.N)
    st.E = int(st.row_ptr[-1].item()) if st.N else 0   # the one host sync: the caller must size edge_index
    count_launch()
    return st


def radius_graph_fill(st, edge_index=None, rev=None):
    """phase 2: the edges in the reference's order into `edge_index` [2,E] int64 and, per slot, the id of the
    reversed edge into `rev` [E] int32 (allocated here unless the caller passes its own buffers; no `rev` with
    criteria edges, which are not symmetric)"""
    lib = _lib.load()
    dev = st.pos.device
    if edge_index is None:
        edge_index = torch.empty(2, st.E, dtype=torch.int64, device=dev)
    assert edge_index.shape == (2, st.E) and edge_index.is_contiguous()
    if st.crit is not None:
        check(lib.e3b_pair_graph_fill(ptr(st.pos), 3, ptr(st.node_ptr), st.G, st.N, st.r_max, ctypes.byref(st.crit),
                                      ptr(st.row_ptr), st.E, ptr(edge_index), stream()))
        count_launch()
        return edge_index, None
    if rev is None:
        rev = torch.empty(st.E, dtype=torch.int32, device=dev)
    assert rev.shape == (st.E,) and rev.is_contiguous()
    if st.cells is not None:
        grids, cell_start, order = st.cells
        check(lib.e3b_cell_graph_fill(ptr(st.pos), 3, ptr(st.node_ptr), st.G, st.N, st.r_max, ptr(grids), ptr(cell_start),
                                      ptr(order), ptr(st.row_ptr), st.E, ptr(edge_index), ptr(rev), stream()))
    else:
        check(lib.e3b_radius_graph_fill(ptr(st.pos), 3, ptr(st.node_ptr), st.G, st.N, st.r_max, ptr(st.row_ptr), st.E,
                                        ptr(edge_index), ptr(rev), stream()))
    count_launch()
    return edge_index, rev


def padding_positions(n, r_max, device):
    """[n, 3] positions of n padding atoms for `radius_graph_fill_padded`: isolated pairs (2p, 2p + 1), the two atoms
    of a pair min(1, r_max / 2) apart, pairs 3 r_max apart (they form a graph of their own, so where they sit relative to
    the real atoms does not matter)."""
    j = torch.arange(n, device=device)
    pos = torch.zeros(n, 3, dtype=torch.float32, device=device)
    pos[:, 0] = (j // 2).float() * (3.0 * float(r_max))
    pos[:, 1] = (j % 2).float() * min(1.0, 0.5 * float(r_max))
    return pos


def radius_graph_fill_padded(st, n_pad, n_edges_total, edge_index, rev, nbr32=None):
    """phase 2 for a batch whose last `n_pad` atoms are padding pairs (see `padding_positions`): fills the caller's
    buffers with exactly `n_edges_total` edges (include/e3b200.h: e3b_radius_graph_fill_padded) and rewrites the padding
    rows of st.row_ptr; st.E becomes n_edges_total."""
    lib = _lib.load()
    assert st.crit is None and st.cells is None, "bucketed padding is for the plain radius graph"
    assert edge_index.shape == (2, n_edges_total) and edge_index.is_contiguous() and rev.shape == (n_edges_total,)
    check(lib.e3b_radius_graph_fill_padded(ptr(st.pos), 3, ptr(st.node_ptr), st.G, st.N, n_pad, st.r_max, ptr(st.row_ptr), st.E,
                                           n_edges_total, ptr(edge_index), ptr(rev), ptr(nbr32), stream()))
    count_launch(3)
    st.E = n_edges_total
    return edge_index, rev


def edges_per_graph(st):
    g = st.row_ptr[st.node_ptr]
    return (g[1:] - g[:-1]).view(-1, 1)


def radius_graph(pos, n_nodes_per_graph, r_max, criteria=None, data=None):
    """Neighbour list in the reference's order plus its CSR views.
    pos [N,3] float32 (cuda), n_nodes_per_graph int64 [G].  -> edge_index int64 [2,E], n_edges [G], GraphCSR"""
    st = radius_graph_count(pos, n_nodes_per_graph, r_max, criteria, data)
    return radius_graph_finish(st)


def radius_graph_finish(st):
    edge_index, rev = radius_graph_fill(st)
    if rev is None:                     # criteria edges: general grouping of an arbitrary (sorted) edge list
        csr = build_csr(edge_index, st.N)
    else:
        # symmetric graph: in-edges of n are the reversed out-edges; slot k of node n holds the
        # edge (nbr -> n) whose id is rev[k]; out-grouping is the identity (edges sorted by source)
        csr = GraphCSR(st.N, st.E, st.row_ptr, edge_index[1].to(torch.int32), rev, st.row_ptr, None)
    edge_index._e3b_csr = csr
    return edge_index, edges_per_graph(st), csr


# ------------------------------------------------------------------------------------------
# Kernel launchers (no autograd).  The autograd Functions below are written on top of these, and
# their backward passes are themselves Functions, so that the graph of a gradient can be built
# (reference nn/output.py:39-43: create_graph=self.training, force-matching losses).  When the
# backward runs without create_graph, autograd executes it in no-grad mode and the inner
# Functions record nothing.
def k_edge_fwd(pos, edge_index, want_len=True):
    lib = _lib.load()
    require_cuda(pos, edge_index)
    pos = pos.contiguous()
    E = edge_index.shape[1]
    vec = torch.empty(E, 3, dtype=pos.dtype, device=pos.device)
    length = torch.empty(E, dtype=pos.dtype, device=pos.device) if want_len else None
    check(lib.e3b_edge_vectors_fwd(dtype_code(pos), ptr(pos), ptr(edge_index), E, ptr(vec), ptr(length), stream()))
    count_launch()
    return vec, length


def k_edge_scatter(gvec, glen, vec, length, n, csr):
    """gpos[a] = sum over edges into a of (gvec + glen vec/len) - the same over edges out of a"""
    lib = _lib.load()
    gpos = torch.empty(n, 3, dtype=vec.dtype, device=vec.device)
    gvec = gvec.contiguous() if gvec is not None else None
    glen = glen.contiguous() if glen is not None else None
    check(lib.e3b_edge_vectors_bwd(dtype_code(vec), ptr(gvec), ptr(glen), ptr(vec), ptr(length), n,
                                   ptr(csr.in_ptr), ptr(csr.in_eid), ptr(csr.out_ptr), ptr(csr.out_eid), ptr(gpos), stream()))
    count_launch()
    return gpos


def k_sh_fwd(vec, lmax, normalize):
    lib = _lib.load()
    require_cuda(vec)
    n = vec.shape[0]
    sh = torch.empty(n, (lmax + 1) ** 2, dtype=vec.dtype, device=vec.device)
    check(lib.e3b_sh_fwd(dtype_code(vec), ptr(vec), n, lmax, int(normalize), ptr(sh), stream()))
    count_launch()
    return sh


def k_sh_bwd(vec, gsh, lmax, normalize):
    lib = _lib.load()
    gvec = torch.empty_like(vec)
    check(lib.e3b_sh_bwd(dtype_code(vec), ptr(vec), ptr(gsh.contiguous()), vec.shape[0], lmax, int(normalize),
                         ptr(gvec), stream()))
    count_launch()
    return gvec


def k_radial_fwd(r, bw, params):
    lib = _lib.load()
    require_cuda(r, bw)
    r_max, r_min, one_over_r, cutoff_kind, p = params
    n, nb = r.shape[0], bw.shape[0]
    out = torch.empty(n, nb, dtype=r.dtype, device=r.device)
    check(lib.e3b_radial_fwd(dtype_code(r), ptr(r), n, ptr(bw), nb, r_max, r_min, one_over_r, cutoff_kind, p,
                             ptr(out), stream()))
    count_launch()
    return out


def k_radial_bwd(r, gout, bw, params, need_w):
    """-> (d/dr [n], d/d bessel_w [n_basis] or None)"""
    lib = _lib.load()
    r_max, r_min, one_over_r, cutoff_kind, p = params
    n, nb = r.shape[0], bw.shape[0]
    gr = torch.empty_like(r)
    nblk = lib.e3b_radial_bwd_blocks(n)
    part = torch.empty(nblk, nb, dtype=r.dtype, device=r.device)
    check(lib.e3b_radial_bwd(dtype_code(r), ptr(r), ptr(gout.contiguous()), n, ptr(bw), nb, r_max, r_min, one_over_r,
                             cutoff_kind, p, ptr(gr), ptr(part), stream()))
    count_launch()
    return gr, (part.sum(0) if need_w else None)


# closed forms of the two per-edge embeddings, used ONLY to differentiate their backward kernels
# once more (Hessian-vector products of a [E,3] -> [E,9] and a [E] -> [E,n_basis] map); values and
# first derivatives always come from the kernels above
def _sh_closed_form(vec, lmax, normalize):
    x, y, z = vec.unbind(-1)
    if normalize:
        r = (x * x + y * y + z * z).sqrt().clamp_min(1e-12)
        x, y, z = x / r, y / r, z / r
    s3, s5 = 1.7320508075688772, 2.23606797749979
    cols = [torch.ones_like(x)]
    if lmax >= 1:
        cols += [s3 * x, s3 * y, s3 * z]
    if lmax >= 2:
        cols += [s5 * s3 * x * z, s5 * s3 * x * y, s5 * (y * y - 0.5 * (x * x + z * z)), s5 * s3 * y * z,
                 s5 * (s3 / 2) * (z * z - x * x)]
    return torch.stack(cols, dim=-1)


def _radial_closed_form(r, bw, params):
    r_max, r_min, one_over_r, cutoff_kind, p = params
    span = r_max - r_min
    b = (2.0 / span) * torch.sin(bw * r.unsqueeze(-1) / span)
    if one_over_r:
        b = b / r.unsqueeze(-1)
    x = r / r_max
    if cutoff_kind == 1:
        c = (x - 1) ** 2 * (x + 1) ** 2 * (x.abs() < 1.0).to(r.dtype)
    else:
        c = (1.0 - ((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(x, p) + p * (p + 2.0) * torch.pow(x, p + 1.0)
             - (p * (p + 1.0) / 2) * torch.pow(x, p + 2.0)) * (x < 1.0).to(r.dtype)
    return b * c.unsqueeze(-1)


def _vjp_of_vjp(closed_form, inputs, cotangent, hhs):
    """inputs: tensors (detached here); B_i = d<cotangent, closed_form(*inputs)>/d inputs[i];
    returns the gradients of sum_i <hhs[i], B_i> (None entries skipped) with respect to every
    input and to the cotangent."""
    with torch.enable_grad():
        leaves = [t.detach().requires_grad_(True) for t in inputs]
        cot = cotangent.detach().requires_grad_(True)
        out = closed_form(*leaves)
        idx = [i for i, h in enumerate(hhs) if h is not None]
        Bs = torch.autograd.grad(out, [leaves[i] for i in idx], cot, create_graph=True)
        grads = torch.autograd.grad(Bs, leaves + [cot], [hhs[i] for i in idx], allow_unused=True)
    return grads


# ------------------------------------------------------------------------------------------
class _EdgeVectors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, edge_index, csr):
        vec, length = k_edge_fwd(pos, edge_index)
        ctx.csr, ctx.n = csr, pos.shape[0]
        ctx.save_for_backward(vec, length, edge_index)
        ctx.set_materialize_grads(False)
        return vec, length

    @staticmethod
    def backward(ctx, gvec, glen):
        vec, length, edge_index = ctx.saved_tensors
        if gvec is None and glen is None:
            return None, None, None
        return _EdgeVectorsBwd.apply(gvec, glen, vec, length, edge_index, ctx.csr, ctx.n), None, None


class _EdgeVectorsBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gvec, glen, vec, length, edge_index, csr, n):
        ctx.save_for_backward(glen, vec, length, edge_index)
        return k_edge_scatter(gvec, glen, vec, length, n, csr)

    @staticmethod
    @once_differentiable
    def backward(ctx, hpos):
        # gpos = S(gvec + glen vec / len), S = adjoint of D: pos -> pos[dst] - pos[src]
        glen, vec, length, edge_index = ctx.saved_tensors
        d, _ = k_edge_fwd(hpos.contiguous(), edge_index, want_len=False)           # D hpos  [E,3]
        g_gvec = d if ctx.needs_input_grad[0] else None
        g_glen = g_vec = g_len = None
        if glen is not None:
            inv = torch.where(length > 0, 1.0 / length, torch.zeros_like(length))
            dv = (d * vec).sum(-1)
            g_glen = dv * inv
            g_vec = d * (glen * inv).unsqueeze(-1)
            g_len = -glen * dv * inv * inv
        return g_gvec, g_glen, g_vec, g_len, None, None, None


def edge_vectors(pos, edge_index, csr):
    return _EdgeVectors.apply(pos.contiguous(), edge_index.contiguous(), csr)


class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, lmax, normalize):
        ctx.save_for_backward(vec)
        ctx.lmax, ctx.normalize = lmax, int(normalize)
        return k_sh_fwd(vec, lmax, normalize)

    @staticmethod
    def backward(ctx, gsh):
        (vec,) = ctx.saved_tensors
        return _SphericalHarmonicsBwd.apply(vec, gsh, ctx.lmax, ctx.normalize), None, None


class _SphericalHarmonicsBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, gsh, lmax, normalize):
        ctx.save_for_backward(vec, gsh)
        ctx.lmax, ctx.normalize = lmax, normalize
        return k_sh_bwd(vec, gsh, lmax, normalize)

    @staticmethod
    @once_differentiable
    def backward(ctx, hvec):
        vec, gsh = ctx.saved_tensors
        g_vec, g_gsh = _vjp_of_vjp(lambda v: _sh_closed_form(v, ctx.lmax, ctx.normalize), [vec], gsh, [hvec])
        return g_vec, g_gsh, None, None


def spherical_harmonics(vec, lmax, normalize=True):
    """[n,3] -> [n,(lmax+1)^2], e3nn 'component' normalisation, l <= 2"""
    return _SphericalHarmonics.apply(vec.contiguous(), lmax, normalize)


class _Radial(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r, bessel_w, r_max, r_min, one_over_r, cutoff_kind, p):
        bw = bessel_w.to(r.dtype).contiguous()
        ctx.params = (r_max, r_min, int(one_over_r), cutoff_kind, float(p))
        ctx.save_for_backward(r, bessel_w)
        return k_radial_fwd(r, bw, ctx.params)

    @staticmethod
    def backward(ctx, gout):
        r, bessel_w = ctx.saved_tensors
        gr, gw = _RadialBwd.apply(r, gout, bessel_w, ctx.params, ctx.needs_input_grad[1])
        return gr, gw, None, None, None, None, None


class _RadialBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r, gout, bessel_w, params, need_w):
        bw = bessel_w.to(r.dtype).contiguous()
        ctx.save_for_backward(r, gout, bessel_w)
        ctx.params, ctx.need_w = params, need_w
        ctx.set_materialize_grads(False)
        gr, gw = k_radial_bwd(r, gout, bw, params, need_w)
        return gr, (gw.to(bessel_w.dtype) if gw is not None else None)

    @staticmethod
    @once_differentiable
    def backward(ctx, hr, hw):
        r, gout, bessel_w = ctx.saved_tensors
        if hr is None and hw is None:
            return None, None, None, None, None
        bw = bessel_w.to(r.dtype)
        g_r, g_bw, g_gout = _vjp_of_vjp(lambda rr, ww: _radial_closed_form(rr, ww, ctx.params), [r, bw], gout,
                                        [hr, hw.to(r.dtype) if hw is not None else None])
        return g_r, g_gout, (g_bw.to(bessel_w.dtype) if g_bw is not None else None), None, None


def radial_basis(r, bessel_w, r_max, r_min=0.0, one_over_r=True, cutoff_kind=0, p=6.0):
    """Bessel x cutoff embedding of distances r [n] -> [n, n_basis]"""
    return _Radial.apply(r.contiguous().reshape(-1), bessel_w, float(r_max), float(r_min), one_over_r, int(cutoff_kind), p)


# When a list, every fused TP-conv launch appends (tag, start_event, end_event) recorded on the
# launching stream; bench.py uses it to time the dominant kernel inside the timed region.
TIMING = None


class stage:
    """``with ops.stage("name"):`` brackets a group of launches with CUDA events when TIMING is a list
    (bench.py --breakdown); free otherwise."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.end = _timed(("stage", self.name))
        return self

    def __exit__(self, *exc):
        if self.end is not None:
            self.end.record()
        return False


def _timed(tag):
    if TIMING is None:
        return None
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    TIMING.append((tag, s, e))
    s.record()
    return e


# ------------------------------------------------------------------------------------------
class TPPlan:
    """Owns an e3b_tp_plan (immutable after creation, shareable across streams)."""

    def __init__(self, structure, w3j_sign_preset=0):
        lib = _lib.load()
        mul = structure.uniform_mul
        if mul is None:
            raise RuntimeError("libe3b200 tensor-product plans need one common multiplicity for all input blocks "
                               f"and mul-1 spherical harmonics, got {structure.irreps_in} x {structure.irreps_sh}")
        d = _lib.TpDesc()
        d.mul = mul
        d.n_in = len(structure.irreps_in)
        for b, blk in enumerate(structure.irreps_in):
            d.in_l[b], d.in_p[b] = blk.ir.l, blk.ir.p
        d.n_sh = len(structure.irreps_sh)
        for s, blk in enumerate(structure.irreps_sh):
            d.sh_l[s], d.sh_p[s] = blk.ir.l, blk.ir.p
        d.n_paths = len(structure.paths)
        for q, path in enumerate(structure.paths):
            d.path_in[q], d.path_sh[q], d.path_lout[q], d.path_slot[q] = path.i_in, path.i_sh, path.ir_out.l, path.slot
        d.w3j_sign_preset = w3j_sign_preset
        handle = ctypes.c_void_p()
        check(lib.e3b_tp_plan_create(ctypes.byref(d), ctypes.byref(handle)))
        self.handle = handle
        dims = [ctypes.c_int32() for _ in range(5)]
        check(lib.e3b_tp_plan_dims(handle, *[ctypes.byref(v) for v in dims]))
        self.x_dim, self.sh_dim, self.w_dim, self.y_dim, self.n_part_f32 = [v.value for v in dims]
        self.specialized = bool(lib.e3b_tp_plan_is_specialized(handle))
        self.structure = structure

    def __del__(self):
        try:
            if self.handle:
                _lib.load().e3b_tp_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def k_tp_fwd(plan, csr, x, sh, w):
    lib = _lib.load()
    require_cuda(x, sh, w)
    N, E = x.shape[0], sh.shape[0]
    assert x.shape[1] == plan.x_dim and sh.shape[1] == plan.sh_dim and w.shape == (E, plan.w_dim), \
        (x.shape, sh.shape, w.shape, plan.x_dim, plan.sh_dim, plan.w_dim)
    assert csr.n_nodes == N and csr.n_edges == E
    y = torch.empty(N, plan.y_dim, dtype=x.dtype, device=x.device)
    end = _timed(("fwd", len(plan.structure.paths), plan.structure.uniform_mul, plan.x_dim, plan.y_dim, N, E))
    check(lib.e3b_tpconv_fwd(plan.handle, dtype_code(x), N, E, ptr(x), ptr(sh), ptr(w), ptr(csr.in_ptr),
                             ptr(csr.in_nbr), ptr(csr.in_eid), ptr(y), stream()))
    if end is not None:
        end.record()
    count_launch()
    return y


def k_tp_bwd(plan, csr, x, sh, w, gy, need_x, need_sh):
    """-> (d/dx [N, x_dim] or None, d/dsh [E, sh_dim] or None, d/dw [E, w_dim])"""
    lib = _lib.load()
    N, E = x.shape[0], sh.shape[0]
    fast = plan.specialized and x.dtype == torch.float32
    alloc = torch.empty if fast else torch.zeros
    n_part = plan.n_part_f32 if fast else 1
    gx_edge = alloc(E, plan.x_dim, dtype=x.dtype, device=x.device) if need_x else None
    gsh_part = alloc(E, n_part, plan.sh_dim, dtype=x.dtype, device=x.device) if need_sh else None
    gw = torch.empty_like(w)
    if E:
        gy = gy.contiguous()
        end = _timed(("bwd", len(plan.structure.paths), plan.structure.uniform_mul, plan.x_dim, plan.y_dim, N, E))
        check(lib.e3b_tpconv_bwd(plan.handle, dtype_code(x), N, E, ptr(x), ptr(sh), ptr(w), ptr(gy),
                                 ptr(csr.in_ptr), ptr(csr.in_nbr), ptr(csr.in_eid), ptr(gx_edge), ptr(gsh_part),
                                 ptr(gw), stream()))
        if end is not None:
            end.record()
        count_launch()
    gx = None
    if need_x:
        gx = k_segment_sum(gx_edge, csr.out_ptr, csr.out_eid, N)
    gsh = None
    if need_sh:
        gsh = gsh_part.sum(1) if n_part > 1 else gsh_part.view(E, plan.sh_dim)
    return gx, gsh, gw


def k_segment_sum(src2, seg_ptr, ids, n_out):
    """out[s] = sum of the rows ids[k] (k itself when ids is None), k in [seg_ptr[s], seg_ptr[s+1])"""
    lib = _lib.load()
    require_cuda(src2)
    out = torch.empty(n_out, src2.shape[1], dtype=src2.dtype, device=src2.device)
    check(lib.e3b_segment_sum(dtype_code(src2), ptr(src2), src2.shape[1], ptr(seg_ptr), ptr(ids), n_out, ptr(out), stream()))
    count_launch()
    return out


def _add(a, b):
    return b if a is None else (a if b is None else a + b)


class _TPConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sh, w, plan, csr):
        ctx.plan, ctx.csr = plan, csr
        ctx.save_for_backward(x, sh, w)
        return k_tp_fwd(plan, csr, x, sh, w)

    @staticmethod
    def backward(ctx, gy):
        x, sh, w = ctx.saved_tensors
        gx, gsh, gw = _TPConvBwd.apply(x, sh, w, gy, ctx.plan, ctx.csr, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return gx, gsh, gw, None, None


class _TPConvBwd(torch.autograd.Function):
    """(x, sh, w, gy) -> (d/dx, d/dsh, d/dw) of <gy, T(x, sh, w)>.  T is trilinear, so the
    adjoint of this map is made of T and of this map with one argument replaced by a cotangent:
    with F(x, sh, w, g) = <g, T(x, sh, w)> and cotangents (hx, hsh, hw) of the three outputs,
    the pulled-back scalar is F(hx, sh, w, gy) + F(x, hsh, w, gy) + F(x, sh, hw, gy)."""

    @staticmethod
    def forward(ctx, x, sh, w, gy, plan, csr, need_x, need_sh):
        ctx.plan, ctx.csr = plan, csr
        ctx.save_for_backward(x, sh, w, gy)
        ctx.set_materialize_grads(False)
        return k_tp_bwd(plan, csr, x, sh, w, gy, need_x, need_sh)

    @staticmethod
    def backward(ctx, hx, hsh, hw):
        x, sh, w, gy = ctx.saved_tensors
        plan, csr = ctx.plan, ctx.csr
        nx, nsh, nw, ngy = ctx.needs_input_grad[:4]
        g_x = g_sh = g_w = g_gy = None
        if hx is not None:                                   # F(hx, sh, w, gy)
            hx = hx.contiguous()
            if ngy:
                g_gy = _add(g_gy, _TPConv.apply(hx, sh, w, plan, csr))
            if nsh or nw:
                _, a, b = _TPConvBwd.apply(hx, sh, w, gy, plan, csr, False, nsh)
                g_sh, g_w = _add(g_sh, a), _add(g_w, b if nw else None)
        if hsh is not None:                                  # F(x, hsh, w, gy)
            hsh = hsh.contiguous()
            if ngy:
                g_gy = _add(g_gy, _TPConv.apply(x, hsh, w, plan, csr))
            if nx or nw:
                a, _, b = _TPConvBwd.apply(x, hsh, w, gy, plan, csr, nx, False)
                g_x, g_w = _add(g_x, a), _add(g_w, b if nw else None)
        if hw is not None:                                   # F(x, sh, hw, gy)
            hw = hw.contiguous()
            if ngy:
                g_gy = _add(g_gy, _TPConv.apply(x, sh, hw, plan, csr))
            if nx or nsh:
                a, b, _ = _TPConvBwd.apply(x, sh, hw, gy, plan, csr, nx, nsh)
                g_x, g_sh = _add(g_x, a), _add(g_sh, b)
        return g_x, g_sh, g_w, g_gy, None, None, None, None


def tp_conv(x_imu, sh, w, plan, csr):
    """y[n] = sum over incoming edges of the weighted uvu tensor product (imu layouts)."""
    return _TPConv.apply(x_imu.contiguous(), sh.contiguous(), w.contiguous(), plan, csr)


# ------------------------------------------------------------------------------------------
class _SegmentSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, seg_ptr, seg_index, n_out):
        src2 = src.contiguous().reshape(src.shape[0], -1)
        out = k_segment_sum(src2, seg_ptr, None, n_out)
        ctx.save_for_backward(seg_ptr, seg_index)
        ctx.n_out = n_out
        return out.reshape(n_out, *src.shape[1:])

    @staticmethod
    def backward(ctx, g):
        seg_ptr, seg_index = ctx.saved_tensors
        return _SegmentBroadcast.apply(g, seg_ptr, seg_index, ctx.n_out), None, None, None


class _SegmentBroadcast(torch.autograd.Function):
    """rows of a segment all receive the segment's value (adjoint of the segmented sum)"""

    @staticmethod
    def forward(ctx, g, seg_ptr, seg_index, n_out):
        ctx.save_for_backward(seg_ptr, seg_index)
        ctx.n_out = n_out
        return g[seg_index]

    @staticmethod
    def backward(ctx, h):
        seg_ptr, seg_index = ctx.saved_tensors
        return _SegmentSum.apply(h, seg_ptr, seg_index, ctx.n_out), None, None, None


def segment_sum(src, seg_ptr, seg_index, n_out):
    """rows of src are grouped in consecutive segments (seg_ptr [n_out+1]); seg_index [rows]
    is the segment of each row (used by the backward gather)."""
    return _SegmentSum.apply(src, seg_ptr, seg_index, n_out)


# ------------------------------------------------------------------------------------------
ACT_CODES = {None: 0, "silu": 1, "tanh": 2, "ssp": 3, "tanhlu": 4, "abs": 5}


def k_gate_fwd(desc, x, out_dim):
    lib = _lib.load()
    require_cuda(x)
    out = torch.empty(x.shape[0], out_dim, dtype=x.dtype, device=x.device)
    check(lib.e3b_gate_fwd(ctypes.byref(desc), dtype_code(x), ptr(x), x.shape[0], ptr(out), stream()))
    count_launch()
    return out


def k_gate_bwd(desc, x, gout):
    lib = _lib.load()
    gin = torch.empty_like(x)
    check(lib.e3b_gate_bwd(ctypes.byref(desc), dtype_code(x), ptr(x), ptr(gout.contiguous()), x.shape[0], ptr(gin), stream()))
    count_launch()
    return gin


def k_gate_bwd2(desc, x, gout, ggin):
    """adjoint of k_gate_bwd: -> (d<ggin, gin>/dx, d<ggin, gin>/dgout)"""
    lib = _lib.load()
    g_x, g_gout = torch.empty_like(x), torch.empty_like(gout)
    check(lib.e3b_gate_bwd2(ctypes.byref(desc), dtype_code(x), ptr(x), ptr(gout.contiguous()), ptr(ggin.contiguous()),
                            x.shape[0], ptr(g_x), ptr(g_gout), stream()))
    count_launch()
    return g_x, g_gout


class _Gate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, desc, out_dim):
        ctx.desc = desc
        ctx.save_for_backward(x)
        return k_gate_fwd(desc, x, out_dim)

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        return _GateBwd.apply(x, gout.contiguous(), ctx.desc), None, None


class _GateBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gout, desc):
        ctx.desc = desc
        ctx.save_for_backward(x, gout)
        return k_gate_bwd(desc, x, gout)

    @staticmethod
    @once_differentiable
    def backward(ctx, ggin):
        x, gout = ctx.saved_tensors
        g_x, g_gout = k_gate_bwd2(ctx.desc, x, gout, ggin)
        return g_x, g_gout, None


def gate(x, desc, out_dim):
    return _Gate.apply(x.contiguous(), desc, out_dim)


class _LayerNorm(torch.autograd.Function):
    """per (node, irreps block): x / sqrt(sum x^2 / mul + eps) * std_block (first order; the second-order
    mode uses the closed form in e3_layers.nn.pointwise.LayerNormalization)"""

    @staticmethod
    def forward(ctx, x, std, muls, ls, eps):
        lib = _lib.load()
        require_cuda(x, std)
        nb = len(muls)
        c_mul, c_l = (ctypes.c_int32 * nb)(*muls), (ctypes.c_int32 * nb)(*ls)
        sw = std.to(x.dtype).contiguous()
        y = torch.empty_like(x)
        rinv = torch.empty(x.shape[0], nb, dtype=x.dtype, device=x.device)
        check(lib.e3b_layernorm_fwd(dtype_code(x), ptr(x), x.shape[0], nb, c_mul, c_l, ptr(sw), float(eps), ptr(y), ptr(rinv), stream()))
        count_launch()
        ctx.save_for_backward(x, rinv, sw)
        ctx.blocks, ctx.std_dtype = (c_mul, c_l, nb), std.dtype
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        lib = _lib.load()
        x, rinv, sw = ctx.saved_tensors
        c_mul, c_l, nb = ctx.blocks
        gx = torch.empty_like(x)
        part = torch.empty(lib.e3b_layernorm_bwd_blocks(x.shape[0]), nb, dtype=x.dtype, device=x.device) \
            if ctx.needs_input_grad[1] else None
        check(lib.e3b_layernorm_bwd(dtype_code(x), ptr(x), ptr(gy.contiguous()), ptr(rinv), x.shape[0], nb, c_mul, c_l,
                                    ptr(sw), ptr(gx), ptr(part), stream()))
        count_launch()
        return gx, (part.sum(0).to(ctx.std_dtype) if part is not None else None), None, None, None


def layer_norm(x, std, muls, ls, eps=1e-6):
    return _LayerNorm.apply(x.contiguous(), std, tuple(muls), tuple(ls), eps)


def layout_convert(x, irreps, to_imu):
    """kernel-side mul_ir <-> imu conversion (no autograd; used for buffers/tests)"""
    lib = _lib.load()
    require_cuda(x)
    x = x.contiguous()
    nb = len(irreps)
    mul = (ctypes.c_int32 * nb)(*[b.mul for b in irreps])
    ls = (ctypes.c_int32 * nb)(*[b.ir.l for b in irreps])
    out = torch.empty_like(x)
    check(lib.e3b_layout_convert(dtype_code(x), ptr(x), x.shape[0], nb, mul, ls, int(to_imu), ptr(out), stream()))
    count_launch()
    return out


# ------------------------------------------------------------------------------------------
class PackedWeight:
    """A weight in the form the tcgen05 GEMM consumes (TF32 hi/lo split, UMMA canonical tiles)."""

    def __init__(self, N, K, device, n_sets=1):
        lib = _lib.load()
        self.N, self.K, self.n_sets = int(N), int(K), int(n_sets)
        self.tile_n = lib.e3b_gemm_tile_n(self.N, self.K)
        self.set_floats = lib.e3b_gemm_packed_floats(self.N, self.K)      # floats of one weight set
        self.buf = torch.empty(self.n_sets * self.set_floats, dtype=torch.float32, device=device)


def gemm_pack(views):
    """views: list of (src fp32 CUDA tensor, element offset, s1, s2, sk, d, n2_valid, N, K): element (n, k),
    n = n1 * d + n2, is src.flat[offset + n1 * s1 + n2 * s2 + k * sk] (zero for n2 >= n2_valid > 0).  Two more entries
    (n_sets, set_stride) pack n_sets weights of that shape, set i read set_stride elements further (grouped rows).
    -> list of PackedWeight
    (one kernel launch per E3B_GEMM_MAX_GROUP views)."""
    lib = _lib.load()
    out = []
    for lo in range(0, len(views), _lib.E3B_GEMM_MAX_GROUP):
        chunk = views[lo:lo + _lib.E3B_GEMM_MAX_GROUP]
        descs = (_lib.GemmPackDesc * len(chunk))()
        for i, view in enumerate(chunk):
            src, off, s1, s2, sk, d, n2_valid, N, K = view[:9]
            n_sets, set_stride = view[9:] if len(view) > 9 else (1, 0)
            require_cuda(src)
            assert src.dtype == torch.float32
            pw = PackedWeight(N, K, src.device, n_sets)
            descs[i].n_sets, descs[i].set_stride = n_sets, set_stride
            out.append(pw)
            descs[i].src, descs[i].dst = src.data_ptr() + 4 * off, pw.buf.data_ptr()
            descs[i].s1, descs[i].s2, descs[i].sk, descs[i].d, descs[i].N, descs[i].K = s1, s2, sk, d, N, K
            descs[i].n2_valid = n2_valid
        check(lib.e3b_gemm_pack(descs, len(chunk), stream()))
        count_launch()
    return out


def gemm_problem(A, Bp, C, M, a_off=0, a_rows=None, c_off=0, c_rows=None, c_col_stride=1, alpha=1.0, epilogue=0,
                 accumulate=False, aux=None, aux_d=1, aux_group=None, H=None, act_cst=1.0, groups=None):
    """One problem of a grouped launch: C[r, n] = epilogue(sum_k A[r, k] B[n, k]).  A / C are fp32 CUDA
    tensors used as raw storage (element offsets a_off / c_off); ``a_rows`` / ``c_rows`` = (s1, s2, d)
    give the affine row addressing base + (r // d) * s1 + (r % d) * s2 (default: dense rows).  ``groups`` (a `RowGroups`):
    M counts VIRTUAL rows, row group r // d stands for the actual group groups.row_map[r // d] of A and C and uses the
    weight set groups.b_sel[(r // d) >> 7] of a PackedWeight with n_sets > 1."""
    N, K = Bp.N, Bp.K
    g = _lib.GemmProblem()
    a_s1, a_s2, a_d = a_rows if a_rows is not None else (K, 0, 1)
    V = (aux_group or aux.shape[1]) if aux is not None else 0
    n_out = N if epilogue != 1 else N // V
    c_s1, c_s2, c_d = c_rows if c_rows is not None else (n_out * c_col_stride, 0, 1)
    g.A, g.a_s1, g.a_s2, g.a_d = A.data_ptr() + 4 * a_off, a_s1, a_s2, a_d
    g.B_packed = Bp.buf.data_ptr()
    g.C, g.c_s1, g.c_s2, g.c_s3, g.c_d = C.data_ptr() + 4 * c_off, c_s1, c_s2, c_col_stride, c_d
    if aux is not None:
        assert aux.is_contiguous() and aux.dtype == torch.float32
        g.aux, g.aux_ld, g.aux_d, g.V, g.aux_cols = aux.data_ptr(), aux.stride(0), aux_d, V, aux.shape[1]
    else:
        g.aux, g.aux_ld, g.aux_d, g.V, g.aux_cols = None, 0, 1, 0, 0
    if H is not None:
        assert H.dtype == torch.float32 and H.stride(1) == 1
        g.H, g.h_ld = H.data_ptr(), H.stride(0)
    else:
        g.H, g.h_ld = None, 0
    g.M, g.N, g.K = M, N, K
    g.epilogue, g.accumulate, g.alpha, g.act_cst = epilogue, int(bool(accumulate)), float(alpha), float(act_cst)
    if groups is not None:
        assert Bp.n_sets == groups.n_sets and a_d == c_d
        g.row_map, g.b_sel, g.b_set_stride = groups.row_map.data_ptr(), groups.b_sel.data_ptr(), Bp.set_floats
        g.n_blocks = groups.n_blocks.data_ptr()
    else:
        g.row_map, g.b_sel, g.b_set_stride, g.n_blocks = None, None, 0, None
    return g


class RowGroups:
    """Rows (nodes) of a grouped GEMM in VIRTUAL order: sorted by weight set (species), every set padded to a multiple
    of 128 rows.  row_map [n_virtual] int32: virtual -> actual row (-1 = padding); b_sel [n_virtual / 128] int32: the
    weight set of every block; n_blocks [1] int32: blocks in use (the kernel skips the rest of the static bound n_virtual);
    table [n_sets, V]: the attribute row of every set; params: what `table` depends on."""
    __slots__ = ("row_map", "b_sel", "n_blocks", "n_virtual", "n_sets", "table", "params", "idx", "_onehot")

    def onehot(self):
        """[N, n_sets] float32 membership rows (the expansion operand of the weight-gradient kernel)"""
        if self._onehot is None:
            self._onehot = torch.nn.functional.one_hot(self.idx, self.n_sets).to(torch.float32).contiguous()
        return self._onehot


def species_row_groups(idx, n_sets):
    """idx [N] int64 in [0, n_sets) -> RowGroups (table / params left to the caller).  Static shapes, no host
    synchronisation: n_virtual = 128 * (ceil(N / 128) + n_sets) bounds every distribution of the rows over the sets."""
    N, dev = idx.shape[0], idx.device
    NB = (N + 127) // 128 + n_sets
    # [S, N] with the scan along the CONTIGUOUS dimension (torch's outer-dimension scan walks the rows one by one: 3 ms)
    member = (idx.view(1, -1) == torch.arange(n_sets, device=dev).view(-1, 1)).to(torch.int32)
    seen = torch.cumsum(member, 1)                                          # rows of set s among the first n + 1
    blocks = (seen[:, -1].long() + 127) // 128 if N else torch.zeros(n_sets, dtype=torch.int64, device=dev)
    bend = torch.cumsum(blocks, 0)
    rank = seen.gather(0, idx.view(1, -1)).view(-1).long() - 1              # position among the rows of the same set
    slot = (bend - blocks).index_select(0, idx) * 128 + rank
    g = RowGroups()
    g.row_map = torch.full((NB * 128,), -1, dtype=torch.int32, device=dev)
    g.row_map.scatter_(0, slot, torch.arange(N, dtype=torch.int32, device=dev))
    g.b_sel = torch.searchsorted(bend, torch.arange(NB, device=dev), right=True).clamp_(max=n_sets - 1).to(torch.int32)
    g.n_blocks = bend[-1:].to(torch.int32)
    g.idx, g._onehot = idx, None
    g.n_virtual, g.n_sets, g.table, g.params = NB * 128, n_sets, None, ()
    return g


SPECIES_SC = os.environ.get("E3B_SPECIES_SC", "1") != "0"


def species_groups_of(attrs):
    """`RowGroups` of the nodes by species when the node attributes are a function of the species alone (tagged
    `_e3b_species` by PointwiseLinear applied to the output of OneHotEncoding: the reference's embedCategorial,
    configs/layer_configs.py:8-30), else None.  Built once per forward pass, remembered on the tensor."""
    sp = getattr(attrs, "_e3b_species", None)
    if sp is None or not SPECIES_SC or not attrs.is_cuda or attrs.dtype != torch.float32:
        return None
    grp = getattr(attrs, "_e3b_groups", None)
    if grp is None:
        idx, S, lin = sp
        grp = species_row_groups(idx, S)
        with torch.no_grad():
            grp.table = lin(torch.eye(S, dtype=attrs.dtype, device=attrs.device)).contiguous()
        grp.params = tuple(lin.parameters())
        attrs._e3b_groups = grp
    return grp


def sc_weight_sets(cache, paths, V, W, grp):
    """Self-connection weights W[u, v, w] (reference nn/message_passing.py:81-87, e3nn FullyConnectedTensorProduct with
    scalar attributes) contracted with the attribute row of every species, W_eff[s][u, w] = sum_v table[s, v] W[u, v, w],
    packed as grp.n_sets weight sets per path for both directions of the map: {'fwd': [PackedWeight per path] (rows w,
    K = u), 'bwd': [...] (rows u, K = w)}.  `paths`: [(mul_in, mul_out, offset in W)]; `cache`: a dict of the caller,
    the entry is recomputed only when W or a parameter behind the table changed."""
    S = grp.n_sets
    key = (WEIGHTS_EPOCH, W.data_ptr(), W._version) + tuple((p.data_ptr(), p._version) for p in grp.params)
    hit = cache.get("sets")
    if hit is not None and hit[0] == key:
        return hit[1]
    with torch.no_grad():
        parts, fwd, bwd, o = [], [], [], 0
        for m1, mo, off in paths:
            Wq = W[off:off + m1 * V * mo].view(m1, V, mo)
            parts.append(torch.einsum("sv,uvw->suw", grp.table, Wq).reshape(-1))       # [s, u, w]
            fwd.append((o, 1, 0, mo, 1, 0, mo, m1, S, m1 * mo))
            bwd.append((o, mo, 0, 1, 1, 0, m1, mo, S, m1 * mo))
            o += S * m1 * mo
        flat = torch.cat(parts)
        out = {"fwd": gemm_pack([(flat,) + v for v in fwd]), "bwd": gemm_pack([(flat,) + v for v in bwd])}
    cache["sets"] = (key, out)
    return out


def gemm_run(problems):
    """launches the problems (built by gemm_problem) grouped by tile shape, <= 8 per kernel"""
    lib = _lib.load()
    classes = {}
    for g in problems:
        if g.M == 0 or g.N == 0:
            continue
        classes.setdefault((g.K <= 64, g.N > 64), []).append(g)      # the launch classes of e3b_gemm_tile_n
    for group in classes.values():
        for lo in range(0, len(group), _lib.E3B_GEMM_MAX_GROUP):
            chunk = group[lo:lo + _lib.E3B_GEMM_MAX_GROUP]
            arr = (_lib.GemmProblem * len(chunk))(*chunk)
            check(lib.e3b_gemm_run(arr, len(chunk), stream()))
            count_launch()


def gemm_tf32x3(A, B, C, M, N, K, a_rows=None, c_rows=None, c_col_stride=1, alpha=1.0, reduce_aux=None, aux_d=1,
                epilogue=None, H=None, act_cst=1.0, accumulate=False):
    """C[r, n] = alpha * sum_k A[r, k] B[n, k] on the tcgen05 tensor cores (3xTF32, fp32 accumulate);
    B is [N, K] (row stride B.stride(0)) and is packed here -- convenience form of
    gemm_pack + gemm_problem + gemm_run.  With ``reduce_aux`` ([*, V], V in {16, 32}) the epilogue
    contracts every group of V accumulator columns with aux[r // aux_d] (self-connection)."""
    require_cuda(A, B, C)
    assert A.dtype == B.dtype == C.dtype == torch.float32
    (Bp,) = gemm_pack([(B, 0, B.stride(0), 0, B.stride(1) if B.dim() == 2 else 1, 1, 0, N, K)])
    if epilogue is None:
        epilogue = 1 if reduce_aux is not None else 0
    gemm_run([gemm_problem(A, Bp, C, M, a_rows=a_rows, c_rows=c_rows, c_col_stride=c_col_stride, alpha=alpha,
                           epilogue=epilogue, aux=reduce_aux, aux_d=aux_d, H=H, act_cst=act_cst, accumulate=accumulate)])
    return C


# ------------------------------------------------------------------------------------------
# Weight gradients: reductions over all rows, C[m, n] = alpha sum_r A'[r, m] B[r, n] (csrc/wgrad_tf32x3.cu)
_WGRAD_WS = {}


def wgrad_problem(A, B, C, R, K1, K2, a_off=0, a_rows=None, b_off=0, b_rows=None, c_off=0, c_rows=None, c_col_stride=1,
                  alpha=1.0, accumulate=False, aux=None, aux_d=1):
    """One problem of a grouped weight-gradient launch.  A / B / C are fp32 CUDA tensors used as raw storage (element
    offsets *_off); ``a_rows`` / ``b_rows`` = (s1, s2, d): row r starts at base + (r // d) * s1 + (r % d) * s2 (default:
    dense rows of K1 / K2 floats); ``c_rows`` = (s1, s2, d) addresses output row m likewise (default: dense [M, K2]).
    With ``aux`` [*, V] the rows of A are expanded to (v, u): A'[r, v * K1 + u] = A[r, u] * aux[r // aux_d, v]."""
    g = _lib.WgradProblem()
    V = aux.shape[1] if aux is not None else 0
    a_s1, a_s2, a_d = a_rows if a_rows is not None else (K1, 0, 1)
    b_s1, b_s2, b_d = b_rows if b_rows is not None else (K2, 0, 1)
    c_s1, c_s2, c_d = c_rows if c_rows is not None else (K2 * c_col_stride, 0, 1)
    # The kernel's tile is 128 (rows of C, operand A) x 64 (columns, operand B).  A narrow A with a wide B -- dW of the
    # last radial layer: [E, 64]^T [E, 1920] -- would use half of every A tile and convert the narrow operand once per
    # column tile; with the operands exchanged (C written transposed through its strides) the same product takes half
    # the tiles, all of them full.
    tiles = lambda k1, k2: -(-k1 // 128) * -(-k2 // 64)
    if aux is None and c_d == 1 and tiles(K2, K1) < tiles(K1, K2):
        A, B, a_off, b_off, K1, K2 = B, A, b_off, a_off, K2, K1
        (a_s1, a_s2, a_d), (b_s1, b_s2, b_d) = (b_s1, b_s2, b_d), (a_s1, a_s2, a_d)
        c_s1, c_s2, c_col_stride = c_col_stride, 0, c_s1            # C^T[m', n'] = C[n', m']
    g.A, g.a_s1, g.a_s2, g.a_d = A.data_ptr() + 4 * a_off, a_s1, a_s2, a_d
    g.B, g.b_s1, g.b_s2, g.b_d = B.data_ptr() + 4 * b_off, b_s1, b_s2, b_d
    if aux is not None:
        assert aux.dtype == torch.float32 and aux.stride(1) == 1
        g.aux, g.aux_ld, g.aux_d, g.V = aux.data_ptr(), aux.stride(0), aux_d, V
    else:
        g.aux, g.aux_ld, g.aux_d, g.V = None, 0, 1, 0
    g.C, g.c_s1, g.c_s2, g.c_s3, g.c_d = C.data_ptr() + 4 * c_off, c_s1, c_s2, c_col_stride, c_d
    g.R, g.K1, g.K2, g.alpha, g.accumulate = R, K1, K2, float(alpha), int(bool(accumulate))
    return g


def wgrad_supported(*tensors, widths=()):
    return (all(t is not None and t.is_cuda and t.dtype == torch.float32 for t in tensors)
            and all(w % 4 == 0 and w > 0 for w in widths))


def wgrad_run(problems, device):
    """launches the problems (built by wgrad_problem), <= 8 per kernel pair"""
    lib = _lib.load()
    group = [g for g in problems if g.R and g.K1 and g.K2]
    if group:
        for lo in range(0, len(group), _lib.E3B_WGRAD_MAX_GROUP):
            chunk = group[lo:lo + _lib.E3B_WGRAD_MAX_GROUP]
            arr = (_lib.WgradProblem * len(chunk))(*chunk)
            need = lib.e3b_wgrad_workspace_floats(arr, len(chunk))
            if need < 0:
                raise RuntimeError("libe3b200: unsupported weight-gradient problem (alignment / widths)")
            key = (device.index, torch.cuda.current_stream(device).cuda_stream)
            ws = _WGRAD_WS.get(key)
            if ws is None or ws.numel() < need:
                ws = torch.empty(max(need, 1 << 20), dtype=torch.float32, device=device)
                _WGRAD_WS[key] = ws
            check(lib.e3b_wgrad_run(arr, len(chunk), ptr(ws), stream()))
            count_launch(2)


def k_wgrad(x, g, alpha=1.0):
    """alpha * x^T @ g for x [R, K1], g [R, K2] (row-major, contiguous) on the tensor cores -> [K1, K2]"""
    require_cuda(x, g)
    R, K1 = x.shape
    K2 = g.shape[1]
    out = torch.empty(K1, K2, dtype=torch.float32, device=x.device)
    wgrad_run([wgrad_problem(x, g, out, R, K1, K2, alpha=alpha)], x.device)
    return out


# ------------------------------------------------------------------------------------------
# Dense map with a small weight matrix as a differentiable node on the tensor cores (second-order mode):
# products of the form [many rows, K] x [K, N] run on the tcgen05 3xTF32 kernel, whichever argument the
# weight is; the weight gradient (a reduction over the rows) is a library GEMM.  The backward is made of the
# same node with the weight transposed, so the graph of a gradient can be differentiated again.
FORCE_DENSE_FUNCTION = False       # tests: take this path on CPU tensors too (with the launcher replaced)


_DENSE_PACKS = {}


def k_dense(x, W, alpha, trans):
    """alpha * x @ (W^T if trans else W); x [M, K] contiguous fp32, W a 2-D fp32 view (any strides)"""
    require_cuda(x, W)
    M, K = x.shape
    N = W.shape[0] if trans else W.shape[1]
    s_n, s_k = (W.stride(0), W.stride(1)) if trans else (W.stride(1), W.stride(0))
    Bp = None
    if isinstance(W, torch.nn.Parameter):      # packed once per parameter version and orientation (see k_sc)
        key = (id(W), trans)
        hit = _DENSE_PACKS.get(key)
        ver = (WEIGHTS_EPOCH, W._version, W.data_ptr())
        if hit is not None and hit[0] == ver and hit[2]() is W:      # the same live parameter object, unchanged
            Bp = hit[1]
    if Bp is None:
        (Bp,) = gemm_pack([(W, 0, s_n, 0, s_k, 1, 0, N, K)])
        if isinstance(W, torch.nn.Parameter):
            if len(_DENSE_PACKS) > 256:
                _DENSE_PACKS.clear()
            _DENSE_PACKS[key] = (ver, Bp, weakref.ref(W))
    out = torch.empty(M, N, dtype=torch.float32, device=x.device)
    gemm_run([gemm_problem(x, Bp, out, M, alpha=alpha)])
    return out


def dense_supported(x, W):
    K = x.shape[-1]
    ok = FORCE_DENSE_FUNCTION or (x.is_cuda and x.dtype == torch.float32 and W.dtype == torch.float32)
    return (ok and x.dtype == W.dtype and K % 4 == 0
            and x.numel() > 0 and W.numel() > 0 and x.numel() // K < 2 ** 31)


class _Dense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, alpha, trans):
        ctx.alpha, ctx.trans = alpha, trans
        ctx.save_for_backward(x, W)
        return k_dense(x, W, alpha, trans)

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        gx = gW = None
        gy = gy.contiguous()
        if ctx.needs_input_grad[0]:
            N = gy.shape[1]
            gx = _Dense.apply(gy, W, ctx.alpha, not ctx.trans) if N % 4 == 0 else \
                ctx.alpha * (gy @ (W if ctx.trans else W.t()))
        if ctx.needs_input_grad[1] and needs_grad_now(W):
            a, b = (gy, x) if ctx.trans else (x, gy)
            if not torch.is_grad_enabled() and wgrad_supported(a, b, widths=(a.shape[1], b.shape[1])) and a.is_contiguous():
                gW = k_wgrad(a, b, ctx.alpha)          # split-K tcgen05 reduction over all rows
            else:                                      # a graph of this gradient is being recorded: plain torch
                gW = ctx.alpha * (a.t() @ b)
        return gx, gW, None, None


def dense(x, W, alpha):
    """alpha * x @ W for x [..., K] and W [K, N] (a view of a parameter)"""
    K, N = W.shape
    x2 = x.reshape(-1, K).contiguous()
    return _Dense.apply(x2, W, float(alpha), False).reshape(*x.shape[:-1], N)


# ------------------------------------------------------------------------------------------
def gemm_waves(problems_with_target):
    """[(problem, target block id, block already written)] -> list of launches such that a problem accumulating into
    a block runs after the first writer of that block; sets .accumulate accordingly."""
    waves, depth = [], {}
    for g, tgt, pre_written in problems_with_target:
        d = depth.get(tgt, 1 if pre_written else 0)
        g.accumulate = 1 if d > 0 else 0
        while len(waves) <= d:
            waves.append([])
        waves[d].append(g)
        depth[tgt] = d + 1
    return [w for w in waves if w]


class SCSpec:
    """Static description of a block-wise dense map between irreps for the multilinear nodes below; rows in the
    imu layout.  V > 0: self-connection (e3nn FullyConnectedTensorProduct with one block of V scalar attributes,
    nn/message_passing.py:81-87), weight blocks [u, v, w].  V == 0: per-irrep linear map (e3nn o3.Linear,
    nn/message_passing.py:58-63, nn/pointwise.py:87-92), weight blocks [u, w], no attributes."""

    def __init__(self, irreps_in, irreps_out, V, paths):
        self.irreps_in, self.irreps_out, self.V = irreps_in, irreps_out, V
        self.Vg = 0 if V == 0 else (16 if V <= 16 else 32)
        self.paths = paths                      # (i_in, i_out, weight offset, alpha)
        self.x_off, self.Din = self._offsets(irreps_in)
        self.c_off, self.Dout = self._offsets(irreps_out)
        self.ok = (V <= 32 and all(b.mul % 4 == 0 for b in irreps_in) and all(b.mul % 4 == 0 for b in irreps_out))

    @staticmethod
    def _offsets(irreps):
        out, o = [], 0
        for b in irreps:
            out.append(o)
            o += b.dim
        return out, o


def k_sc(spec, src, attrs, W, to_out):
    """to_out: y[z,d,w] = alpha sum_{u,v} W[u,v,w] src[z,d,u] a[z,v]   (src = features, rows of irreps_in)
    else:     gx[z,d,u] = alpha sum_{v,w} W[u,v,w] src[z,d,w] a[z,v]  (src = output gradient, rows of irreps_out)
    (without the v index and `a` when spec.V == 0).  Grouped tcgen05 GEMMs, one problem per irreps block pair,
    the attribute contraction in the epilogue; imu layouts."""
    require_cuda(src, attrs, W)
    N, V, Vg = src.shape[0], spec.V, spec.Vg
    grp = species_groups_of(attrs) if V else None
    if grp is not None:
        # attributes = table[species]: one K = mul GEMM per path with the species' contracted weight, rows in species order
        if not hasattr(spec, "_set_cache"):
            spec._set_cache = {}
        packs = sc_weight_sets(spec._set_cache, [(spec.irreps_in[i].mul, spec.irreps_out[o].mul, off) for i, o, off, _ in spec.paths],
                               V, W, grp)["fwd" if to_out else "bwd"]
        extra, rows = dict(groups=grp), grp.n_virtual
    else:
        views = []
        for i, o, off, alpha in spec.paths:
            m1, mo = spec.irreps_in[i].mul, spec.irreps_out[o].mul
            if V:
                views.append((W, off, 1, mo, V * mo, Vg, V, mo * Vg, m1) if to_out else (W, off, V * mo, mo, 1, Vg, V, m1 * Vg, mo))
            else:
                views.append((W, off, 1, 0, mo, 1, 0, mo, m1) if to_out else (W, off, mo, 0, 1, 1, 0, m1, mo))
        # a parameter is packed once per version and direction (forward, backward and the second-order passes of a training
        # step all see the same weights); anything else (a cotangent standing in for W) is packed on the spot
        packs = None
        if isinstance(W, torch.nn.Parameter):
            cache = spec.__dict__.setdefault("_pack_cache", {})
            key = (WEIGHTS_EPOCH, W.data_ptr(), W._version)
            hit = cache.get(to_out)
            if hit is not None and hit[0] == key:
                packs = hit[1]
        if packs is None:
            packs = gemm_pack(views)
            if isinstance(W, torch.nn.Parameter):
                cache[to_out] = (key, packs)
        extra, rows = (dict(epilogue=1, aux=attrs, aux_group=Vg) if V else {}), N
    D_src, D_dst = (spec.Din, spec.Dout) if to_out else (spec.Dout, spec.Din)
    dst = torch.empty(N, D_dst, dtype=torch.float32, device=src.device)
    # The problem descriptors of a (spec, direction, grouped?) combination are built once; a call then only patches the
    # pointers and the row count (filling ~30 ctypes fields per path was most of the host time of a training step).
    tkey = (to_out, grp is not None)
    tcache = spec.__dict__.setdefault("_prob_cache", {})
    tmpl = tcache.get(tkey)
    if tmpl is None:
        probs, written = [], set()
        meta = []
        for q, (i, o, off, alpha) in enumerate(spec.paths):
            bi, bo = spec.irreps_in[i], spec.irreps_out[o]
            d = bi.ir.dim
            if V and grp is None:
                extra["aux_d"] = d
            if to_out:
                a_off, c_off, tgt = spec.x_off[i], spec.c_off[o], o
                g = gemm_problem(src, packs[q], dst, rows * d, a_off=a_off, a_rows=(D_src, bi.mul, d), c_off=c_off,
                                 c_rows=(D_dst, bo.mul, d), alpha=alpha, **extra)
            else:
                a_off, c_off, tgt = spec.c_off[o], spec.x_off[i], i
                g = gemm_problem(src, packs[q], dst, rows * d, a_off=a_off, a_rows=(D_src, bo.mul, d), c_off=c_off,
                                 c_rows=(D_dst, bi.mul, d), alpha=alpha, **extra)
            probs.append((g, tgt, False))
            meta.append((g, q, 4 * a_off, 4 * c_off, d))
            written.add(tgt)
        zero = len(written) < len(spec.irreps_out if to_out else spec.irreps_in)
        tmpl = (meta, gemm_waves(probs), zero)
        tcache[tkey] = tmpl
    else:
        meta = tmpl[0]
        sp, dp = src.data_ptr(), dst.data_ptr()
        ap = attrs.data_ptr() if (V and grp is None) else None
        if ap is not None:
            assert attrs.is_contiguous() and attrs.dtype == torch.float32
        for g, q, a_off, c_off, d in meta:
            g.A, g.C, g.M, g.B_packed = sp + a_off, dp + c_off, rows * d, packs[q].buf.data_ptr()
            if ap is not None:
                g.aux = ap
            if grp is not None:
                g.row_map, g.b_sel, g.n_blocks = grp.row_map.data_ptr(), grp.b_sel.data_ptr(), grp.n_blocks.data_ptr()
    if tmpl[2]:
        dst.zero_()
    for wave in tmpl[1]:
        gemm_run(wave)
    return dst


def _sc_weight_grad_kernel(spec, x, a, W, g):
    """dW of a block-wise linear map (V == 0: dW[u, w]) or of the self-connection (dW[u, v, w] = alpha sum_{z, d}
    x[z, d, u] a[z, v] g[z, d, w]) on the split-K tcgen05 kernel: one problem per irreps path, written straight into the
    flat weight layout; the (x, a) outer product is formed by the kernel's loader."""
    N, V = x.shape[0], spec.V
    covered = sum(spec.irreps_in[i].mul * max(V, 1) * spec.irreps_out[o].mul for i, o, _, _ in spec.paths)
    disjoint = len({off for _, _, off, _ in spec.paths}) == len(spec.paths)
    gW = torch.empty_like(W) if covered == W.numel() and disjoint else torch.zeros_like(W)
    probs, seen = [], set()
    for i, o, off, alpha in spec.paths:
        bi, bo = spec.irreps_in[i], spec.irreps_out[o]
        d = bi.ir.dim
        kw = dict(a_off=spec.x_off[i], a_rows=(spec.Din, bi.mul, d), b_off=spec.c_off[o], b_rows=(spec.Dout, bo.mul, d),
                  c_off=off, alpha=alpha, accumulate=off in seen)
        if V:
            kw.update(aux=a, aux_d=d, c_rows=(bo.mul, V * bo.mul, bi.mul))       # row m = v * m1 + u of the kernel -> W[u, v, :]
        probs.append(wgrad_problem(x, g, gW, N * d, bi.mul, bo.mul, **kw))
        seen.add(off)
    # problems accumulating into a block written by an earlier one must run after it: one launch per "wave"
    first = [p for p in probs if not p.accumulate]
    later = [p for p in probs if p.accumulate]
    wgrad_run(first, x.device)
    for p in later:
        wgrad_run([p], x.device)
    return gW


def _sc_reductions(spec, x, a, W, g, want_a, want_W):
    """d/da and d/dW of <g, S(x, a, W)>:  t[z,u,w] = sum_d x[z,d,u] g[z,d,w] per path first (4x fewer flops than
    going through the (u,v) outer product), then ONE skinny GEMM over all paths for each of the two results
    (a^T T and T Wcat): they are bound by reading T once instead of once per path.  Plain torch contractions,
    so the graph of these gradients can be differentiated again by torch."""
    N, V = x.shape[0], spec.V
    if (want_W and not torch.is_grad_enabled() and spec.ok and wgrad_supported(x, g, W)
            and x.is_contiguous() and g.is_contiguous() and (V == 0 or (a.is_contiguous() and a.shape[1] == V))):
        gW = _sc_weight_grad_kernel(spec, x, a, W, g)
        if not want_a:
            return None, gW
        return _sc_reductions(spec, x, a, W, g, True, False)[0], gW
    if V == 0:                      # plain linear map: dW[u, w] = alpha sum_{z,d} x[z,d,u] g[z,d,w] per block pair
        pieces = []
        for i, o, off, alpha in spec.paths:
            bi, bo = spec.irreps_in[i], spec.irreps_out[o]
            xb = x[:, spec.x_off[i]:spec.x_off[i] + bi.dim].reshape(-1, bi.mul)
            gb = g[:, spec.c_off[o]:spec.c_off[o] + bo.dim].reshape(-1, bo.mul)
            pieces.append((off, (alpha * (xb.t() @ gb)).reshape(-1)))
        pieces.sort(key=lambda p: p[0])
        if sum(p[1].numel() for p in pieces) == W.numel():
            return None, torch.cat([p[1] for p in pieces])
        gW = torch.zeros_like(W)
        for off, piece in pieces:
            gW = gW + torch.nn.functional.pad(piece, (off, W.numel() - off - piece.numel()))
        return None, gW
    ts, wcols, metas = [], [], []
    for i, o, off, alpha in spec.paths:
        bi, bo = spec.irreps_in[i], spec.irreps_out[o]
        xb = x[:, spec.x_off[i]:spec.x_off[i] + bi.dim].reshape(N, bi.ir.dim, bi.mul)
        gb = g[:, spec.c_off[o]:spec.c_off[o] + bo.dim].reshape(N, bi.ir.dim, bo.mul)
        ts.append(torch.bmm(xb.transpose(1, 2), gb).reshape(N, bi.mul * bo.mul))     # [z, (u, w)]
        metas.append((off, alpha, bi.mul, bo.mul))
        if want_a:
            Wp = W[off:off + bi.mul * V * bo.mul].reshape(bi.mul, V, bo.mul)
            wcols.append(alpha * Wp.transpose(0, 1).reshape(V, -1))                  # [v, (u, w)]
    # when no graph is being recorded (the final backward) the attribute gradient takes the K-long tcgen05 GEMM, path by
    # path (no [z, sum_p m1 mo] concatenation: 1.4 GB at W2)
    fast = ts[0].is_cuda and ts[0].dtype == torch.float32 and not torch.is_grad_enabled()
    ga = None
    if want_a and fast and all(t.shape[1] % 4 == 0 for t in ts) and not want_W:
        for t, wc in zip(ts, wcols):
            part = k_dense(t, wc, 1.0, True)
            ga = part if ga is None else ga + part
        return ga, None
    T = torch.cat(ts, dim=1)                                                          # [z, sum_p m1 mo]
    if want_a:
        Wcat = torch.cat(wcols, dim=1)
        ga = k_dense(T, Wcat, 1.0, True) if fast and T.shape[1] % 4 == 0 else T @ Wcat.t()
    gW = None
    if want_W:
        full = (T.t() @ a).t()      # [v, sum_p m1 mo]; this operand order is the faster cuBLAS shape (0.55 vs 0.73 ms)
        pieces, c0 = [], 0
        for off, alpha, m1, mo in metas:
            blk = full[:, c0:c0 + m1 * mo].reshape(V, m1, mo).transpose(0, 1).reshape(-1)   # -> [u, v, w]
            pieces.append((off, alpha * blk))
            c0 += m1 * mo
        pieces.sort(key=lambda p: p[0])
        if sum(p[1].numel() for p in pieces) == W.numel():       # e3nn layout: the paths tile the flat weight
            gW = torch.cat([p[1] for p in pieces])
        else:
            gW = torch.zeros_like(W)
            for off, piece in pieces:
                gW = gW + torch.nn.functional.pad(piece, (off, W.numel() - off - piece.numel()))
    return ga, gW


class _SC(torch.autograd.Function):
    """y = S(x, a, W) (to_out) or gx = Sx(g, a, W) (not to_out): the two row-parallel partial derivatives of the
    quadrilinear form F(x, a, W, g) = <g, S(x, a, W)>; each one's backward is the other plus two reductions."""

    @staticmethod
    def forward(ctx, src, a, W, spec, to_out):
        ctx.spec, ctx.to_out = spec, to_out
        ctx.save_for_backward(src, a, W)
        ctx.species = (getattr(a, "_e3b_species", None), getattr(a, "_e3b_groups", None)) if a is not None else (None, None)
        return k_sc(spec, src, a, W, to_out)

    @staticmethod
    def backward(ctx, h):
        src, a, W = ctx.saved_tensors
        spec = ctx.spec
        h = h.contiguous()
        if ctx.species[0] is not None and getattr(a, "_e3b_species", None) is None:
            a._e3b_species, a._e3b_groups = ctx.species           # provenance of the attributes (see species_groups_of)
        g_src = _SC.apply(h, a, W, spec, not ctx.to_out) if ctx.needs_input_grad[0] else None
        want_a = a is not None and ctx.needs_input_grad[1] and needs_grad_now(a)
        want_W = ctx.needs_input_grad[2] and needs_grad_now(W)
        ga = gW = None
        if want_a or want_W:
            x, g = (src, h) if ctx.to_out else (h, src)
            ga, gW = _sc_reductions(spec, x, a, W, g, want_a, want_W)
        return g_src, ga, gW, None, None


def self_connection(x_imu, attrs, W, spec):
    grp = species_groups_of(attrs) if (SPECIES_SC and spec.V) else None
    if grp is not None:
        return self_connection_species(x_imu, attrs, W, spec, grp)
    return _SC.apply(x_imu.contiguous(), attrs.contiguous(), W.contiguous(), spec, True)


def k_scg(spec, src, packs, grp, to_out):
    """y[z,d,w] = alpha sum_u Weff[s(z)][u,w] src[z,d,u] (to_out) or gx[z,d,u] = alpha sum_w Weff[s(z)][u,w] src[z,d,w]:
    one grouped-row tcgen05 GEMM per irreps path, `packs` = the species weight sets of that direction"""
    require_cuda(src)
    N = src.shape[0]
    D_src, D_dst = (spec.Din, spec.Dout) if to_out else (spec.Dout, spec.Din)
    dst = torch.empty(N, D_dst, dtype=torch.float32, device=src.device)
    tcache = spec.__dict__.setdefault("_prob_cache", {})
    tmpl = tcache.get(("scg", to_out))
    if tmpl is None:
        probs, written, meta = [], set(), []
        for q, (i, o, off, alpha) in enumerate(spec.paths):
            bi, bo = spec.irreps_in[i], spec.irreps_out[o]
            d = bi.ir.dim
            if to_out:
                a_off, c_off, tgt, ma, mc = spec.x_off[i], spec.c_off[o], o, bi.mul, bo.mul
            else:
                a_off, c_off, tgt, ma, mc = spec.c_off[o], spec.x_off[i], i, bo.mul, bi.mul
            g = gemm_problem(src, packs[q], dst, grp.n_virtual * d, a_off=a_off, a_rows=(D_src, ma, d), c_off=c_off,
                             c_rows=(D_dst, mc, d), alpha=alpha, groups=grp)
            probs.append((g, tgt, False))
            meta.append((g, q, 4 * a_off, 4 * c_off, d))
            written.add(tgt)
        tmpl = (meta, gemm_waves(probs), len(written) < len(spec.irreps_out if to_out else spec.irreps_in))
        tcache[("scg", to_out)] = tmpl
    else:
        sp, dp = src.data_ptr(), dst.data_ptr()
        rm, bs, nb = grp.row_map.data_ptr(), grp.b_sel.data_ptr(), grp.n_blocks.data_ptr()
        for g, q, a_off, c_off, d in tmpl[0]:
            g.A, g.C, g.M, g.B_packed = sp + a_off, dp + c_off, grp.n_virtual * d, packs[q].buf.data_ptr()
            g.row_map, g.b_sel, g.n_blocks = rm, bs, nb
    if tmpl[2]:
        dst.zero_()
    for wave in tmpl[1]:
        gemm_run(wave)
    return dst


def _scg_layout(spec, S):
    """offsets of the per-path blocks [S, mul_in, mul_out] inside the flat species weight"""
    offs, o = [], 0
    for i, oo, _, _ in spec.paths:
        offs.append(o)
        o += S * spec.irreps_in[i].mul * spec.irreps_out[oo].mul
    return offs, o


def _scg_weight_grad(spec, x, g, grp):
    """dWeff[s][u,w] = alpha sum_{z of species s, d} x[z,d,u] g[z,d,w] per path.  No graph being recorded: the split-K
    tcgen05 kernel with the one-hot species rows as the expansion operand, written straight into the flat layout;
    otherwise plain torch contractions (differentiable again)."""
    S = grp.n_sets
    offs, total = _scg_layout(spec, S)
    N = x.shape[0]
    if not torch.is_grad_enabled() and wgrad_supported(x, g) and x.is_contiguous() and g.is_contiguous():
        out = torch.empty(total, dtype=torch.float32, device=x.device)
        onehot = grp.onehot()
        probs = []
        for (i, o, _, alpha), off in zip(spec.paths, offs):
            bi, bo = spec.irreps_in[i], spec.irreps_out[o]
            d = bi.ir.dim
            probs.append(wgrad_problem(x, g, out, N * d, bi.mul, bo.mul, a_off=spec.x_off[i], a_rows=(spec.Din, bi.mul, d),
                                       b_off=spec.c_off[o], b_rows=(spec.Dout, bo.mul, d), aux=onehot, aux_d=d, c_off=off,
                                       c_rows=(bi.mul * bo.mul, bo.mul, bi.mul), alpha=alpha))
        wgrad_run(probs, x.device)
        return out
    onehot = grp.onehot().to(x.dtype)
    pieces = []
    for i, o, _, alpha in spec.paths:
        bi, bo = spec.irreps_in[i], spec.irreps_out[o]
        xb = x[:, spec.x_off[i]:spec.x_off[i] + bi.dim].reshape(N, bi.ir.dim, bi.mul)
        gb = g[:, spec.c_off[o]:spec.c_off[o] + bo.dim].reshape(N, bi.ir.dim, bo.mul)
        t = torch.bmm(xb.transpose(1, 2), gb).reshape(N, bi.mul * bo.mul)               # [z, (u, w)]
        pieces.append((alpha * (onehot.t() @ t)).reshape(-1))                            # [s, u, w]
    return torch.cat(pieces)


class _SCG(torch.autograd.Function):
    """The self-connection when the node attributes are a function of the species (reference embedCategorial,
    configs/layer_configs.py:8-30): y = S(x, Weff) with Weff[s] = sum_v table[s, v] W[:, v, :] contracted OUTSIDE this node
    by differentiable torch ops on [S, ...] tensors, so what is left here is bilinear in (x, Weff) and per-node work never
    sees the attribute index: no [z, mul_in * mul_out] intermediates, no attribute-gradient GEMMs.  to_out False: the
    adjoint map in x (gx = Sx(g, Weff)); each is the other's backward plus the reduction `_scg_weight_grad`."""

    @staticmethod
    def forward(ctx, src, Weff, spec, grp, to_out):
        ctx.spec, ctx.grp, ctx.to_out = spec, grp, to_out
        packs = getattr(Weff, "_e3b_sets", None)
        if packs is None:
            S = grp.n_sets
            offs, _ = _scg_layout(spec, S)
            fwd, bwd = [], []
            for (i, o, _, _), off in zip(spec.paths, offs):
                m1, mo = spec.irreps_in[i].mul, spec.irreps_out[o].mul
                fwd.append((Weff, off, 1, 0, mo, 1, 0, mo, m1, S, m1 * mo))
                bwd.append((Weff, off, mo, 0, 1, 1, 0, m1, mo, S, m1 * mo))
            packs = {True: gemm_pack(fwd), False: gemm_pack(bwd)}
            Weff._e3b_sets = packs
        ctx.packs = packs
        ctx.save_for_backward(src, Weff)
        return k_scg(spec, src, packs[to_out], grp, to_out)

    @staticmethod
    def backward(ctx, h):
        src, Weff = ctx.saved_tensors
        spec, grp = ctx.spec, ctx.grp
        h = h.contiguous()
        if getattr(Weff, "_e3b_sets", None) is None:
            Weff._e3b_sets = ctx.packs
        g_src = _SCG.apply(h, Weff, spec, grp, not ctx.to_out) if ctx.needs_input_grad[0] else None
        gW = None
        if ctx.needs_input_grad[1] and needs_grad_now(Weff):
            x, g = (src, h) if ctx.to_out else (h, src)
            gW = _scg_weight_grad(spec, x, g, grp)
        return g_src, gW, None, None, None


def self_connection_species(x_imu, attrs, W, spec, grp):
    """`self_connection` for attributes = table[species]: the weights are contracted with the table per species by torch
    (tiny, differentiable: gradients reach W and the embedding behind the table), the per-node work is `_SCG`"""
    idx, S, lin = attrs._e3b_species
    table = lin(torch.eye(S, dtype=x_imu.dtype, device=x_imu.device))                  # [S, V], tracked
    V = spec.V
    blocks = []
    for i, o, off, _ in spec.paths:
        m1, mo = spec.irreps_in[i].mul, spec.irreps_out[o].mul
        blocks.append(torch.einsum("sv,uvw->suw", table, W[off:off + m1 * V * mo].view(m1, V, mo)).reshape(-1))
    Weff = torch.cat(blocks)
    return _SCG.apply(x_imu.contiguous(), Weff, spec, grp, True)


class _LiveBlocks(torch.autograd.Function):
    """identity whose backward tags the gradient: "columns outside these irreps blocks are exactly zero" """

    @staticmethod
    def forward(ctx, x, tag):
        ctx.tag = tag
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g._e3b_live_blocks = ctx.tag
        return g, None


def tag_live_blocks(x, live, irreps):
    """x unchanged; in the backward pass the gradient that reaches x's producer carries `_e3b_live_blocks = (live block
    indices, irreps)`.  Used by a per-irrep linear map that reads only some blocks of its input (an energy read-out reads
    the 0e scalars, reference addEnergyOutput, configs/layer_configs.py:101-113): the last interaction block then skips
    every path that cannot reach them (interaction.FusedInteraction.scalar_only).  The tag survives only if this map is
    the sole consumer of x (autograd sums gradients of several consumers into a fresh tensor)."""
    return _LiveBlocks.apply(x, (tuple(live), str(irreps)))


def block_linear(x_imu, W, spec):
    """per-irrep linear map (spec.V == 0) as one bilinear node: imu rows in, imu rows out"""
    return _SC.apply(x_imu.contiguous(), None, W.contiguous(), spec, True)


class _Layout(torch.autograd.Function):
    """mul_ir <-> imu conversion of feature rows as a node (a permutation: the adjoint is the inverse)"""

    @staticmethod
    def forward(ctx, x, irreps, to_imu):
        ctx.irreps, ctx.to_imu = irreps, to_imu
        return k_layout(x, irreps, to_imu)

    @staticmethod
    def backward(ctx, g):
        return _Layout.apply(g.contiguous(), ctx.irreps, not ctx.to_imu), None, None


def k_layout(x, irreps, to_imu):
    return layout_convert(x, irreps, to_imu)


def layout(x, irreps, to_imu):
    return _Layout.apply(x.contiguous(), irreps, to_imu)

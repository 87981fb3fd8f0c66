"""One interaction block (reference ``MessagePassing.forward``, ``nn/message_passing.py:242-262`` around
``FactorizedConvolution.forward`` ``:91-124``) as ONE autograd node on the fp32 B200 path.

Forward, per layer (every step is a libe3b200 kernel; feature rows stay in the channel-fastest
"imu" layout between the kernels, so no transposes or concatenations are materialised):

    x_l   = linear_1(x)                               grouped tcgen05 GEMM (irreps blocks = problems)
    h     = radial MLP hidden layers(edge_radial)     ONE node for all blocks of the network (`_RadialHidden`):
                                                      layer i of every block in one grouped tcgen05 launch
    w     = h @ W_last                                tcgen05 GEMM
    mid   = sum_{e -> n} w_e * (x_l[src] (x) Y_e)     fused gather / CG product / segmented sum
    conv  = linear(mid) / sqrt(avg_num_neighbors)     grouped tcgen05 GEMM (after the reduction)
    conv += sc(x, node_attrs)                         grouped tcgen05 GEMM, attribute contraction in the epilogue
    out   = gate(conv)                                written in both layouts (mul_ir for the caller,
                                                      imu for the next block)

Backward (first order) runs the transposed GEMMs from the same parameter tensors (packed per
role), the backward tensor-product kernel and the activation-derivative epilogues.  Parameter
gradients (skipped in the position-gradient pass of an energy+force evaluation) are library GEMMs on the saved
activations."""
import ctypes
import math

import torch
from torch.autograd.function import once_differentiable

from . import _lib, dense, ops
from ._lib import check, count_launch, ptr, stream
from .irreps import Irreps


def _offsets(irreps):
    out, o = [], 0
    for b in irreps:
        out.append(o)
        o += b.dim
    return out, o


class FusedInteraction:
    """Static description of one MessagePassing layer + its packed weights (cached per parameter version)."""

    def __init__(self, mp):
        conv = mp.conv
        self.mp, self.conv = mp, conv
        self.feat_in = Irreps(conv.irreps_in["input_features"])
        self.conv_out = Irreps(conv.irreps_out["output_features"])
        self.attrs = Irreps(conv.irreps_in["node_attrs"])
        self.structure = conv.tp.tp.structure
        self.mid = self.structure.irreps_mid.simplify()
        self.x_off, self.Din = _offsets(self.feat_in)
        self.c_off, self.Dconv = _offsets(self.conv_out)
        self.m_off, self.Dmid = _offsets(self.mid)
        self.gate = mp.equivariant_nonlin
        self.Dout = self.gate.irreps_out.dim
        self.inv_sqrt_avg = 1.0 / math.sqrt(conv.avg_num_neighbors) if conv.avg_num_neighbors is not None else 1.0
        self.all_scalar_in = all(b.ir.l == 0 for b in self.feat_in)
        self.fc = conv.fc
        self.hs = list(conv.fc.hs)
        self.V = self.attrs.dim
        self.Vg = 16 if self.V <= 16 else 32           # epilogue group width (attributes zero-extended)
        self._cache = {}
        self.reason = self._unsupported()

    # -- which layers take the fused path -------------------------------------------------------
    def _unsupported(self):
        c = self.conv
        if self.structure.uniform_mul is None:
            return "input blocks of different multiplicity"
        if c.sc is None:
            return "no self-connection"
        if len(self.attrs) != 1 or not self.attrs[0].ir.is_scalar() or self.V > 32:
            return "node_attrs must be one block of <= 32 scalars"
        if any(h % 4 for h in self.hs[:-1]) or any(b.mul % 4 for b in self.feat_in) or any(b.mul % 4 for b in self.conv_out):
            return "widths must be multiples of 4"
        return None

    # -- packed weights -------------------------------------------------------------------------
    def _params(self):
        """every parameter the packed weights depend on"""
        c = self.conv
        return [getattr(c.fc, f"layer{i}").weight for i in range(c.fc.n_layers)] + [c.linear_1.weight, c.tp.linear.weight,
                                                                                   c.sc.weight]

    def _hidden_params(self):
        c = self.conv
        return [getattr(c.fc, f"layer{i}").weight for i in range(c.fc.n_layers - 1)]

    def _block_params(self):
        """parameters whose gradients the block's own autograd node produces: last radial layer, linear_1, post linear, sc"""
        c = self.conv
        return [getattr(c.fc, f"layer{c.fc.n_layers - 1}").weight, c.linear_1.weight, c.tp.linear.weight, c.sc.weight]

    def hidden_signature(self):
        return (tuple(self.hs[:-1]), float(self.fc.cst))

    def packs(self, role):
        """role 'fwd' | 'bwd' -> dict of PackedWeight lists, repacked only when a parameter changed"""
        key = (ops.WEIGHTS_EPOCH,) + tuple((p.data_ptr(), p._version) for p in self._params())
        hit = self._cache.get(role)
        if hit is not None and hit[0] == key:
            return hit[1]
        c = self.conv
        views, names = [], []
        fwd = role == "fwd"
        for i in range(c.fc.n_layers):                                   # radial MLP: W[h_in, h_out]
            W = getattr(c.fc, f"layer{i}").weight
            hi, ho = W.shape
            views.append((W, 0, 1, 0, ho, 1, 0, ho, hi) if fwd else (W, 0, ho, 0, 1, 1, 0, hi, ho))
            names.append(("fc", i))
        lin1 = c.linear_1
        for q, (i, o, off, _) in enumerate(lin1.paths):                   # W[u, w]
            mi, mo = lin1.irreps_in[i].mul, lin1.irreps_out[o].mul
            views.append((lin1.weight, off, 1, 0, mo, 1, 0, mo, mi) if fwd else (lin1.weight, off, mo, 0, 1, 1, 0, mi, mo))
            names.append(("lin1", q))
        post = c.tp.linear
        for q, (i, o, off, _) in enumerate(post.paths):                   # W[kk, w]
            mi, mo = post.irreps_in[i].mul, post.irreps_out[o].mul
            views.append((post.weight, off, 1, 0, mo, 1, 0, mo, mi) if fwd else (post.weight, off, mo, 0, 1, 1, 0, mi, mo))
            names.append(("post", q))
        sc, V, Vg = c.sc, self.V, self.Vg
        for q, (i1, i2, o, off, _) in enumerate(sc.paths):                # W[u, v, w]
            m1, mo = sc.irreps_in1[i1].mul, sc.irreps_out[o].mul
            if fwd:   # rows (w, v), v fastest, K = u
                views.append((sc.weight, off, 1, mo, V * mo, Vg, V, mo * Vg, m1))
            else:     # rows (u, v), v fastest, K = w
                views.append((sc.weight, off, V * mo, mo, 1, Vg, V, m1 * Vg, mo))
            names.append(("sc", q))
        packed = ops.gemm_pack(views)
        out = {"fc": [], "lin1": [], "post": [], "sc": []}
        for (kind, _), pw in zip(names, packed):
            out[kind].append(pw)
        self._cache[role] = (key, out)
        return out


    def restricted(self, live):
        """What the backward pass needs when only the output blocks `live` (indices into the gate's output irreps) have a
        non-zero gradient -- the last block under an energy read-out (live = the 0e scalars: 3 of 30 paths) and the block
        before it (live = the inputs of those paths: 0e, 1o, 2e, 15 of 30 paths): the tensor-product plan restricted to
        the paths whose output irrep is still needed, their columns in the full weight row, the problems of the
        post-reduction linear map (with their offsets in the restricted intermediate) and of the self-connection that reach
        live blocks, and the input blocks that receive a gradient (the tag handed to the block before).  None when the
        layer does not have that shape or the restriction saves less than 40 % of the paths."""
        key = ("restricted", tuple(live))
        if key not in self._cache:
            out = None
            st, mul = self.structure, self.structure.uniform_mul
            gate = self.gate
            n_s, n_g = len(gate.irreps_scalars), len(gate.irreps_gated)
            if (mul in (32, 64) and self.fc.n_layers > 1 and len(self.conv_out) == n_s + (1 if n_g else 0) + n_g
                    and all(0 <= b < n_s + n_g for b in live)):
                from .plan import output_restriction
                C = {b for b in live if b < n_s} | {n_s + 1 + (b - n_s) for b in live if b >= n_s}
                if any(b >= n_s for b in live):
                    C.add(n_s)                                    # the gates of the live gated blocks
                need = []
                for c in sorted(C):
                    if all(str(self.conv_out[c].ir) != str(ir) for ir in need):
                        need.append(self.conv_out[c].ir)
                pr = output_restriction(st, need)
                if pr.paths and len(pr.paths) <= 0.6 * len(st.paths):
                    plan = ops.TPPlan(pr)
                    idx = {(p.i_in, p.i_sh, str(p.ir_out)): q for q, p in enumerate(st.paths)}
                    cols = [idx[(p.i_in, p.i_sh, str(p.ir_out))] for p in pr.paths]
                    pmid = pr.irreps_mid.simplify()
                    p_off, _ = _offsets(pmid)
                    where = {str(b.ir): (k, b.mul) for k, b in enumerate(pmid)}
                    post, ok = {}, plan.specialized and plan.x_dim == self.conv.tp.plan.x_dim
                    for q, (i, o, _, _) in enumerate(self.conv.tp.linear.paths):
                        hit = where.get(str(self.mid[i].ir))
                        if o in C and hit is not None:
                            ok = ok and hit[1] == self.mid[i].mul  # every path into a needed irrep is kept
                            post[q] = (hit[0], p_off[hit[0]])
                    sc_q = [q for q, (_, _, o, _, _) in enumerate(self.conv.sc.paths) if o in C]
                    sup = sorted({p.i_in for p in pr.paths} | {self.conv.sc.paths[q][0] for q in sc_q})
                    if ok and post:
                        out = {"plan": plan, "cols": cols, "post": post, "n_mid": len(pmid), "sc_q": sc_q,
                               "sup": tuple(sup) if len(sup) < len(self.feat_in) else None}
            self._cache[key] = out
        return self._cache[key]

    def restricted_fc(self, so, role="bwd"):
        """packed last radial layer restricted to the live weight columns, per parameter version: role 'bwd' for the gradient
        of the hidden activations, role 'fwd' to RECOMPUTE the live columns of the per-edge weights from the saved hidden
        activations ([E, 64] x [64, live]: cheaper than gathering them out of the [E, 1920] rows of the forward pass)"""
        W = getattr(self.conv.fc, f"layer{self.conv.fc.n_layers - 1}").weight
        key = (ops.WEIGHTS_EPOCH, W.data_ptr(), W._version)
        ckey = ("restricted_fc", tuple(so["cols"]))
        hit = self._cache.get(ckey)
        if hit is None or hit[0] != key:
            mul = self.structure.uniform_mul
            with torch.no_grad():
                Wl = torch.cat([W[:, c * mul:(c + 1) * mul] for c in so["cols"]], 1).contiguous()      # [h, live]
            hi, ho = Wl.shape
            packed = ops.gemm_pack([(Wl, 0, ho, 0, 1, 1, 0, hi, ho), (Wl, 0, 1, 0, ho, 1, 0, ho, hi)])
            hit = (key, {"bwd": packed[0], "fwd": packed[1]})
            self._cache[ckey] = hit
        return hit[1][role]

    def sc_sets(self, grp):
        """self-connection weights contracted with the attribute row of every species (`ops.sc_weight_sets`)"""
        sc = self.conv.sc
        return ops.sc_weight_sets(self._cache, [(sc.irreps_in1[i1].mul, sc.irreps_out[o].mul, off) for i1, _, o, off, _ in sc.paths],
                                  self.V, sc.weight, grp)


SPECIES_SC_CALLS = 0      # forward passes of a block whose self-connection took the per-species path (tests)
SCALAR_ONLY = __import__("os").environ.get("E3B_SCALAR_ONLY", "1") != "0"
SCALAR_ONLY_CALLS = 0     # backward passes of a block restricted to the paths that reach its 0e scalars (tests)


_waves = ops.gemm_waves


class _RadialHidden(torch.autograd.Function):
    """The hidden layers of the radial MLPs of ALL interaction blocks of a network as one autograd node.  They depend on
    the edge lengths only, not on the node features, so layer i of every block runs in ONE grouped tcgen05 launch
    (forward: 3 launches instead of 15 for a 5-block network; backward likewise -- autograd calls this node's backward
    once, after every block has delivered its gradient).  Contract with `_Interaction` (both private): output b is the
    last hidden activation of block b; the gradient that comes back for it is already multiplied by the activation
    derivative at that output (fused in the epilogue of the block's last-layer backward GEMM)."""

    @staticmethod
    def forward(ctx, er, fis, *weights):
        ctx.set_materialize_grads(False)
        er = er.contiguous()
        E, dev = er.shape[0], er.device
        hs, cst = fis[0].hs, fis[0].fc.cst
        n_hid = len(hs) - 2
        h = [[er] for _ in fis]
        with ops.stage("f.mlp_hidden"):
            for i in range(n_hid):
                probs = []
                for b, fi in enumerate(fis):
                    out = torch.empty(E, hs[i + 1], dtype=torch.float32, device=dev)
                    src = h[b][-1]
                    probs.append(ops.gemm_problem(src, fi.packs("fwd")["fc"][i], out, E, a_rows=(src.stride(0), 0, 1),
                                                  alpha=1.0 / math.sqrt(hs[i]), epilogue=2, act_cst=cst))
                    h[b].append(out)
                ops.gemm_run(probs)
        ctx.fis, ctx.n_hid = fis, n_hid
        ctx.save_for_backward(*[t for hb in h for t in hb])
        return tuple(hb[-1] for hb in h)

    @staticmethod
    @once_differentiable
    def backward(ctx, *g_out):
        fis, n_hid = ctx.fis, ctx.n_hid
        saved = ctx.saved_tensors
        L = len(fis)
        h = [list(saved[b * (n_hid + 1):(b + 1) * (n_hid + 1)]) for b in range(L)]
        hs, cst = fis[0].hs, fis[0].fc.cst
        E, dev = h[0][0].shape[0], h[0][0].device
        need_er = ctx.needs_input_grad[0]
        need_w = [ctx.needs_input_grad[2 + k] for k in range(L * n_hid)]
        need_params = any(need_w) and not ops.positions_only_active()
        live = [b for b in range(L) if g_out[b] is not None]
        if not live or not (need_er or need_params):
            return (None, None) + (None,) * (L * n_hid)
        gz = [[None] * (n_hid + 1) for _ in range(L)]
        for b in live:
            gz[b][n_hid] = g_out[b].contiguous()
        lo = 0 if need_er else 1
        with ops.stage("b.mlp_hidden"):
            for i in range(n_hid - 1, lo - 1, -1):
                probs = []
                for b in live:
                    out = torch.empty(E, hs[i], dtype=torch.float32, device=dev)
                    probs.append(ops.gemm_problem(gz[b][i + 1], fis[b].packs("bwd")["fc"][i], out, E, alpha=1.0 / math.sqrt(hs[i]),
                                                  epilogue=3 if i > 0 else 0, H=h[b][i] if i > 0 else None, act_cst=cst))
                    gz[b][i] = out
                ops.gemm_run(probs)
        g_er = None
        if need_er:
            g_er = gz[live[0]][0]
            for b in live[1:]:
                g_er = g_er + gz[b][0]
        g_w = [None] * (L * n_hid)
        if need_params:
            probs = []
            for b in live:
                for i in range(n_hid):
                    if not need_w[b * n_hid + i]:
                        continue
                    K1, K2 = hs[i], hs[i + 1]
                    a, g = h[b][i], gz[b][i + 1]
                    if K1 % 4 == 0 and K2 % 4 == 0 and a.is_contiguous():
                        out = torch.empty(K1, K2, dtype=torch.float32, device=dev)
                        probs.append(ops.wgrad_problem(a, g, out, E, K1, K2, alpha=1.0 / math.sqrt(K1)))
                    else:
                        out = (a.t() @ g) * (1.0 / math.sqrt(K1))
                    g_w[b * n_hid + i] = out
            ops.wgrad_run(probs, dev)
        return (g_er, None, *g_w)


SHARED_W = __import__("os").environ.get("E3B_SHARED_W", "1") != "0"
SHARED_CALLS = 0          # forward passes of a block that took the shared-weight-row path (tests)


class _Undirected:
    """undirected-edge view of a symmetric radius graph: uid [E] int32 (row of the shared tensors for edge e), canon [E/2]
    int64 (the direction src < dst of every undirected edge), rev [E] int32, er_u = edge_radial[canon]"""
    __slots__ = ("uid", "canon", "rev", "er_u")


def undirected(er, edge_index, csr):
    """The radial embedding (tagged `_e3b_length_only` by RadialBasisEncoding when its input is the edge length computed
    by computeEdgeVector) is a function of the edge LENGTH, so the two directions of an edge of the radius graph have
    bit-identical rows: everything downstream of it that does not see the direction (the whole radial MLP, i.e. the
    per-edge weights) is evaluated once per undirected edge.  Built once per forward pass, remembered on `er`."""
    u = getattr(er, "_e3b_und", None)
    if u is None:
        E = er.shape[0]
        src, dst = edge_index[0], edge_index[1]
        flag = src < dst
        rank = torch.cumsum(flag, 0, dtype=torch.int32) - 1                 # index among the canonical directions
        rev = csr.in_eid
        uid = torch.where(flag, rank, rank[rev.long()])
        idx = torch.arange(E, device=er.device)
        canon = torch.zeros(E // 2 + 1, dtype=torch.int64, device=er.device)
        canon.scatter_(0, torch.where(flag, rank, torch.full_like(rank, E // 2)).long(), idx)
        u = _Undirected()
        u.uid, u.canon, u.rev = uid.contiguous(), canon[:E // 2].contiguous(), rev
        u.er_u = er.index_select(0, u.canon)
        er._e3b_und = u
    return u


def radial_hidden(er, fi, group):
    """last hidden activation of block `fi`'s radial MLP; computed for every compatible block of `group` at the first
    request and remembered on the `edge_radial` tensor object (a new one every forward pass)"""
    cache = getattr(er, "_e3b_rh", None)
    if cache is None or id(fi) not in cache:
        sig = fi.hidden_signature()
        fis = [fi]
        for m in (group or []):
            f2 = m.fused
            if f2 is not fi and f2.reason is None and f2.hidden_signature() == sig and (cache is None or id(f2) not in cache):
                fis.append(f2)
        weights = [w for f2 in fis for w in f2._hidden_params()]
        outs = _RadialHidden.apply(er, tuple(fis), *weights)
        cache = dict(cache or {})
        for f2, o in zip(fis, outs):
            cache[id(f2)] = o
        er._e3b_rh = cache
    return cache[id(fi)]


class _Interaction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_mi, x_imu, attrs, h_last, Y, fi, csr, und, grp, *params):
        lib = _lib.load()
        ctx.set_materialize_grads(False)
        conv = fi.conv
        src_is_imu = x_imu is not None
        if x_imu is None:
            x_imu = x_mi.contiguous() if fi.all_scalar_in else ops.layout_convert(x_mi, fi.feat_in, True)
        x_imu, attrs, h_last, Y = x_imu.contiguous(), attrs.contiguous(), h_last.contiguous(), Y.contiguous()
        N, E, Ew = x_imu.shape[0], Y.shape[0], h_last.shape[0]       # directed edges, rows of the weight tensor
        dev = x_imu.device
        P = fi.packs("fwd")
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        # ---- linear_1 (imu -> imu)
        lin1 = conv.linear_1
        xl = new(N, fi.Din)
        probs, written = [], set()
        for q, (i, o, off, alpha) in enumerate(lin1.paths):
            bi, bo = fi.feat_in[i], fi.feat_in[o]
            probs.append((ops.gemm_problem(x_imu, P["lin1"][q], xl, N * bi.ir.dim, a_off=fi.x_off[i],
                                           a_rows=(fi.Din, bi.mul, bi.ir.dim), c_off=fi.x_off[o],
                                           c_rows=(fi.Din, bo.mul, bo.ir.dim), alpha=alpha), o, False))
            written.add(o)
        if len(written) < len(fi.feat_in):
            xl.zero_()
        lin1_waves = _waves(probs)
        # ---- self-connection into conv (imu); the linear map after the reduction accumulates into it below
        post, sc = conv.tp.linear, conv.sc
        cv = new(N, fi.Dconv)
        # (the self-connection writes first: its reducing epilogue stores whole sectors without reading;
        #  the dense epilogue of the linear map then accumulates with coalesced read-modify-writes)
        sc_probs, written = [], set()
        sets = fi.sc_sets(grp) if grp is not None else None
        for q, (i1, i2, o, off, alpha) in enumerate(sc.paths):
            bi, bo = fi.feat_in[i1], fi.conv_out[o]
            if grp is not None:
                # attributes = a function of the species: one K = mul GEMM per path with the species' contracted weight
                # (rows in species order through grp.row_map), 1/V of the flops of the attribute-contraction epilogue
                g = ops.gemm_problem(x_imu, sets["fwd"][q], cv, grp.n_virtual * bi.ir.dim, a_off=fi.x_off[i1],
                                     a_rows=(fi.Din, bi.mul, bi.ir.dim), c_off=fi.c_off[o],
                                     c_rows=(fi.Dconv, bo.mul, bo.ir.dim), alpha=alpha, groups=grp)
            else:
                g = ops.gemm_problem(x_imu, P["sc"][q], cv, N * bi.ir.dim, a_off=fi.x_off[i1],
                                     a_rows=(fi.Din, bi.mul, bi.ir.dim), c_off=fi.c_off[o],
                                     c_rows=(fi.Dconv, bo.mul, bo.ir.dim), alpha=alpha, epilogue=1, aux=attrs,
                                     aux_d=bi.ir.dim, aux_group=fi.Vg)
            sc_probs.append((g, o, False))
            written.add(o)
        if len(written | {o for _, o, _, _ in post.paths}) < len(fi.conv_out):
            cv.zero_()
        # linear_1 and the self-connection read the same rows and (per species) have the same K <= 64, N <= 64 shape: their
        # first waves share one grouped launch
        sc_waves = _waves(sc_probs)
        with ops.stage("f.linear_1+self_connection"):
            ops.gemm_run(lin1_waves[0] + sc_waves[0])
            for wave in lin1_waves[1:] + sc_waves[1:]:
                ops.gemm_run(wave)
        # ---- last layer of the radial MLP (the hidden layers are the shared `_RadialHidden` node)
        hs = fi.hs
        n_fc = conv.fc.n_layers
        w = new(Ew, hs[-1])
        with ops.stage("f.mlp_last"):
            ops.gemm_run([ops.gemm_problem(h_last, P["fc"][n_fc - 1], w, Ew, a_rows=(h_last.stride(0), 0, 1),
                                           alpha=1.0 / math.sqrt(hs[-2]), epilogue=0, act_cst=conv.fc.cst)])
        # ---- fused gather + CG tensor product + segmented sum
        plan = conv.tp.plan
        mid = new(N, plan.y_dim)
        end = ops._timed(("fwd", len(plan.structure.paths), plan.structure.uniform_mul, plan.x_dim, plan.y_dim, N, E))
        if und is not None:
            check(lib.e3b_tpconv_fwd_shared(plan.handle, N, E, ptr(xl), ptr(Y), ptr(w), ptr(und.uid), ptr(csr.in_ptr),
                                            ptr(csr.in_nbr), ptr(csr.in_eid), ptr(mid), stream()))
        else:
            check(lib.e3b_tpconv_fwd(plan.handle, 0, N, E, ptr(xl), ptr(Y), ptr(w), ptr(csr.in_ptr), ptr(csr.in_nbr),
                                     ptr(csr.in_eid), ptr(mid), stream()))
        if end is not None:
            end.record()
        count_launch()
        # ---- post-reduction linear (scaled by 1/sqrt(avg_num_neighbors)) + self-connection, both into conv (imu)
        post, sc = conv.tp.linear, conv.sc
        probs = []
        for q, (i, o, off, alpha) in enumerate(post.paths):
            bi, bo = fi.mid[i], fi.conv_out[o]
            probs.append((ops.gemm_problem(mid, P["post"][q], cv, N * bi.ir.dim, a_off=fi.m_off[i],
                                           a_rows=(fi.Dmid, bi.mul, bi.ir.dim), c_off=fi.c_off[o],
                                           c_rows=(fi.Dconv, bo.mul, bo.ir.dim), alpha=alpha * fi.inv_sqrt_avg),
                          o, o in written))
        with ops.stage("f.post_linear"):
            for wave in _waves(probs):
                ops.gemm_run(wave)
        # ---- gate, in both layouts
        out_mi, out_imu = new(N, fi.Dout), new(N, fi.Dout)
        with ops.stage("f.gate"):
            check(lib.e3b_gate_imu_fwd(ctypes.byref(fi.gate.desc), 0, ptr(cv), N, ptr(out_mi), ptr(out_imu), stream()))
        count_launch()
        ctx.fi, ctx.csr, ctx.src_is_imu, ctx.und, ctx.grp = fi, csr, src_is_imu, und, grp
        # parameter gradients are produced whenever a parameter requires them, in training AND in evaluation mode
        # (fine-tuning under model.eval(), gradient diagnostics); the one pass that must not pay for them -- the
        # position gradient of an energy+force evaluation -- is marked by GradientOutput with ops.positions_only
        ctx.want_params = bool(any(p.requires_grad for p in params))
        ctx.save_for_backward(x_imu, attrs, Y, xl, cv, h_last, w, mid if ctx.want_params else None)
        return out_mi, out_imu

    @staticmethod
    @once_differentiable
    def backward(ctx, g_mi, g_imu):
        lib = _lib.load()
        fi, csr = ctx.fi, ctx.csr
        conv = fi.conv
        saved = ctx.saved_tensors
        x_imu, attrs, Y, xl, cv, h_last, w, mid = saved
        und = ctx.und
        N, E, Ew = x_imu.shape[0], Y.shape[0], h_last.shape[0]
        dev = x_imu.device
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        need_x = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_attrs, need_h, need_Y = ctx.needs_input_grad[2], ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        pos_only = ops.positions_only_active()          # parameters never depend on the positions
        need_params = ctx.want_params and any(ctx.needs_input_grad[9:]) and not pos_only
        need_attrs = need_attrs and ops.needs_grad_now(attrs)
        if g_mi is None and g_imu is None:
            return (None,) * len(ctx.needs_input_grad)
        P = fi.packs("bwd")
        # Only the 0e scalars of the output have a non-zero gradient (tag set by the read-out's linear map, see
        # ops.tag_live_blocks): every path of the block that does not reach them contributes exactly zero.
        so = None
        g_only = g_mi if g_imu is None else (g_imu if g_mi is None else None)      # a tag counts only on the sole gradient
        live = getattr(g_only, "_e3b_live_blocks", None) if g_only is not None else None
        if (SCALAR_ONLY and live is not None and not need_params and not need_attrs and E > 0
                and live[1] == str(fi.gate.irreps_out) and conv.tp.plan.specialized):
            so = fi.restricted(live[0])
        if so is not None:
            global SCALAR_ONLY_CALLS
            SCALAR_ONLY_CALLS += 1
        # ---- gate
        g_cv = new(N, fi.Dconv)
        with ops.stage("b.gate"):
            check(lib.e3b_gate_imu_bwd(ctypes.byref(fi.gate.desc), 0, ptr(cv), ptr(g_mi.contiguous()) if g_mi is not None else None,
                                       ptr(g_imu.contiguous()) if g_imu is not None else None, N, ptr(g_cv), stream()))
        count_launch()
        # ---- post linear, transposed: g_mid[z, k, kk] = alpha sum_w W[kk, w] g_cv[z, k, w]
        post, sc, lin1 = conv.tp.linear, conv.sc, conv.linear_1
        plan = conv.tp.plan if so is None else so["plan"]
        g_mid = new(N, plan.y_dim)
        probs, written = [], set()
        for q, (i, o, off, alpha) in enumerate(post.paths):
            bi, bo = fi.mid[i], fi.conv_out[o]
            if so is not None:
                hit = so["post"].get(q)    # a block of the restricted intermediate has the layout of the same irrep's full block
                if hit is not None:
                    probs.append((ops.gemm_problem(g_cv, P["post"][q], g_mid, N * bi.ir.dim, a_off=fi.c_off[o],
                                                   a_rows=(fi.Dconv, bo.mul, bo.ir.dim), c_off=hit[1],
                                                   c_rows=(plan.y_dim, bi.mul, bi.ir.dim), alpha=alpha * fi.inv_sqrt_avg),
                                  hit[0], False))
                    written.add(hit[0])
                continue
            probs.append((ops.gemm_problem(g_cv, P["post"][q], g_mid, N * bi.ir.dim, a_off=fi.c_off[o],
                                           a_rows=(fi.Dconv, bo.mul, bo.ir.dim), c_off=fi.m_off[i],
                                           c_rows=(fi.Dmid, bi.mul, bi.ir.dim), alpha=alpha * fi.inv_sqrt_avg), i, False))
            written.add(i)
        if len(written) < (len(fi.mid) if so is None else so["n_mid"]):
            g_mid.zero_()
        if so is not None:
            # weight columns of the live paths, recomputed from the hidden activations with the column slice of the last
            # radial layer (same arithmetic as the forward pass; a gather out of the full rows cost 0.20 + 0.05 ms per step)
            hs_, n_fc_ = fi.hs, conv.fc.n_layers
            w_live = new(w.shape[0], len(so["cols"]) * fi.structure.uniform_mul)
            with ops.stage("b.w_live"):
                ops.gemm_run([ops.gemm_problem(h_last, fi.restricted_fc(so, "fwd"), w_live, w.shape[0],
                                               a_rows=(h_last.stride(0), 0, 1), alpha=1.0 / math.sqrt(hs_[-2]), epilogue=0,
                                               act_cst=conv.fc.cst)])
            w = w_live
        with ops.stage("b.post_linear"):
            for wave in _waves(probs):
                ops.gemm_run(wave)
        # ---- tensor-product convolution backward
        fast = plan.specialized
        alloc = torch.empty if fast else torch.zeros
        n_part = plan.n_part_f32 if fast else 1
        # d/dx per source node: reduced inside the kernel (TMA reduce-add of each edge's row into its source's row, no
        # per-edge buffer, no segment sum) in the evaluation pass; passes that produce parameter gradients (training) and
        # E3B_DETERMINISTIC=1 keep the bit-reproducible per-edge rows + segment sum
        in_kernel = bool(need_x and fast and E and plan.structure.uniform_mul in (32, 64)
                         and ((not need_params and not ops.DETERMINISTIC) or und is not None))
        gx_edge = alloc(E, plan.x_dim, dtype=torch.float32, device=dev) if need_x and not in_kernel else None
        g_xl = torch.zeros(N, fi.Din, dtype=torch.float32, device=dev) if in_kernel else None
        gsh_part = alloc(E, n_part, plan.sh_dim, dtype=torch.float32, device=dev) if need_Y else None
        gw = new(E, plan.w_dim)
        if E:
            end = ops._timed(("bwd", len(plan.structure.paths), plan.structure.uniform_mul, plan.x_dim, plan.y_dim, N, E))
            if in_kernel:
                check(lib.e3b_tpconv_bwd_nodes(plan.handle, N, E, ptr(xl), ptr(Y), ptr(w), ptr(und.uid) if und is not None else None,
                                               ptr(g_mid), ptr(csr.in_ptr), ptr(csr.in_nbr), ptr(csr.in_eid), ptr(g_xl),
                                               ptr(gsh_part), ptr(gw), stream()))
            else:
                check(lib.e3b_tpconv_bwd(plan.handle, 0, N, E, ptr(xl), ptr(Y), ptr(w), ptr(g_mid), ptr(csr.in_ptr),
                                         ptr(csr.in_nbr), ptr(csr.in_eid), ptr(gx_edge), ptr(gsh_part), ptr(gw), stream()))
            if end is not None:
                end.record()
            count_launch()
        g_Y = None
        if need_Y:
            with ops.stage("b.gsh_sum"):
                g_Y = gsh_part.sum(1) if n_part > 1 else gsh_part.view(E, plan.sh_dim)
        # ---- last radial layer backward (data path): the gradient handed to the shared hidden node is already
        # multiplied by the activation derivative at h_last (epilogue 3), see `_RadialHidden`
        hs = fi.hs
        n_fc = conv.fc.n_layers
        fc_last = P["fc"][n_fc - 1] if so is None else fi.restricted_fc(so)
        g_h = None
        if need_h and und is None:
            g_h = new(E, hs[-2])
            with ops.stage("b.mlp_last"):
                ops.gemm_run([ops.gemm_problem(gw, fc_last, g_h, E, alpha=1.0 / math.sqrt(hs[-2]),
                                               epilogue=3 if n_fc > 1 else 0, H=h_last if n_fc > 1 else None,
                                               act_cst=conv.fc.cst)])
        elif need_h:
            # shared weight rows: the GEMM is linear, so it runs on the directed gradient rows and the two directions of every
            # undirected edge are folded afterwards on the small [E, 64] result, with the activation derivative in the same pass
            g_dir = new(E, hs[-2])
            with ops.stage("b.mlp_last"):
                ops.gemm_run([ops.gemm_problem(gw, fc_last, g_dir, E, alpha=1.0 / math.sqrt(hs[-2]), epilogue=0)])
                g_h = new(Ew, hs[-2])
                check(lib.e3b_pair_sum_act(ptr(g_dir), ptr(und.canon), ptr(und.rev), ptr(h_last), float(conv.fc.cst), Ew,
                                           hs[-2], ptr(g_h), stream()))
                count_launch()
        # ---- d/dx: linear_1 transposed on the reduced edge gradient + self-connection transposed
        g_x = None
        if need_x:
            if not in_kernel:
                g_xl = new(N, fi.Din)
                with ops.stage("b.segment_sum"):
                    check(lib.e3b_segment_sum(0, ptr(gx_edge), plan.x_dim, ptr(csr.out_ptr), ptr(csr.out_eid), N, ptr(g_xl), stream()))
                count_launch()
            g_x = new(N, fi.Din)
            probs, written = [], set()
            for q, (i, o, off, alpha) in enumerate(lin1.paths):
                bi, bo = fi.feat_in[i], fi.feat_in[o]
                probs.append((ops.gemm_problem(g_xl, P["lin1"][q], g_x, N * bi.ir.dim, a_off=fi.x_off[o],
                                               a_rows=(fi.Din, bo.mul, bo.ir.dim), c_off=fi.x_off[i],
                                               c_rows=(fi.Din, bi.mul, bi.ir.dim), alpha=alpha), i, False))
                written.add(i)
            sc_probs = []
            grp = ctx.grp
            sets = fi.sc_sets(grp) if grp is not None else None
            for q, (i1, i2, o, off, alpha) in enumerate(sc.paths):
                bi, bo = fi.feat_in[i1], fi.conv_out[o]
                if so is not None and q not in so["sc_q"]:
                    continue                                     # the gradient of that output block is zero
                if grp is not None:
                    g = ops.gemm_problem(g_cv, sets["bwd"][q], g_x, grp.n_virtual * bi.ir.dim, a_off=fi.c_off[o],
                                         a_rows=(fi.Dconv, bo.mul, bo.ir.dim), c_off=fi.x_off[i1],
                                         c_rows=(fi.Din, bi.mul, bi.ir.dim), alpha=alpha, groups=grp)
                else:
                    g = ops.gemm_problem(g_cv, P["sc"][q], g_x, N * bi.ir.dim, a_off=fi.c_off[o],
                                         a_rows=(fi.Dconv, bo.mul, bo.ir.dim), c_off=fi.x_off[i1],
                                         c_rows=(fi.Din, bi.mul, bi.ir.dim), alpha=alpha, epilogue=1, aux=attrs,
                                         aux_d=bi.ir.dim, aux_group=fi.Vg)
                sc_probs.append((g, i1, i1 in written))
            if len(written | {t for _, t, _ in sc_probs}) < len(fi.feat_in):
                g_x.zero_()
            with ops.stage("b.linear_1"):
                for wave in _waves(probs):
                    ops.gemm_run(wave)
            # several self-connection paths may feed from one input block (0e -> scalars and gates)
            with ops.stage("b.self_connection"):
                for wave in _waves(sc_probs):
                    ops.gemm_run(wave)
        # ---- parameter / attribute gradients (training only): library GEMMs on the saved activations
        g_params = [None] * len(ctx.needs_input_grad[9:])
        g_attrs = None
        if need_params or need_attrs:
            gw_rows = gw if und is None else gw.index_select(0, und.canon) + gw.index_select(0, und.rev.long().index_select(0, und.canon))
            g_params, g_attrs = _param_grads(fi, ctx.needs_input_grad[2:], x_imu, attrs, h_last, gw_rows, mid, g_cv, g_xl, need_attrs)
        g_x_mi = g_x_imu = None
        if need_x:
            if ctx.src_is_imu:
                g_x_imu = g_x
            else:
                g_x_mi = g_x if fi.all_scalar_in else ops.layout_convert(g_x, fi.feat_in, False)
            if so is not None and so["sup"] is not None:
                # only these input blocks received anything: the block before this one may restrict its backward as well
                (g_x_imu if g_x_imu is not None else g_x_mi)._e3b_live_blocks = (so["sup"], str(fi.feat_in))
        return (g_x_mi, g_x_imu, g_attrs, g_h, g_Y, None, None, None, None, *g_params)


def _param_grads(fi, needs, x_imu, attrs, h_last, gw, mid, g_cv, g_xl, need_attrs):
    """parameter gradients of one block (last radial layer, linear_1, post linear, self-connection): reductions over all
    edges / nodes on the split-K tcgen05 kernel (csrc/wgrad_tf32x3.cu), written straight into the flat weight layouts;
    the attribute gradient (a per-node quantity) stays a torch contraction"""
    conv = fi.conv
    n_fc = 1                                           # parameters of this node: [fc last, linear_1, post, sc]
    out = []
    N, dev = x_imu.shape[0], x_imu.device
    probs = []
    if needs[7]:
        K1, K2 = fi.hs[-2], fi.hs[-1]
        if K1 % 4 == 0 and K2 % 4 == 0:
            gw_last = torch.empty(K1, K2, dtype=torch.float32, device=dev)
            probs.append(ops.wgrad_problem(h_last, gw, gw_last, h_last.shape[0], K1, K2, alpha=1.0 / math.sqrt(K1)))
        else:
            gw_last = (h_last.t() @ gw) * (1.0 / math.sqrt(K1))
        out.append(gw_last)
    else:
        out.append(None)
    # linear_1: dW[u, w] = alpha sum_{z, m} x[z, m, u] g_xl[z, m, w]
    lin1 = conv.linear_1
    g = None
    if needs[7 + n_fc]:
        g = torch.zeros_like(lin1.weight)
        if g_xl is not None:
            seen = set()
            for i, o, off, alpha in lin1.paths:
                bi, bo = fi.feat_in[i], fi.feat_in[o]
                probs.append(ops.wgrad_problem(x_imu, g_xl, g, N * bi.ir.dim, bi.mul, bo.mul, a_off=fi.x_off[i],
                                               a_rows=(fi.Din, bi.mul, bi.ir.dim), b_off=fi.x_off[o],
                                               b_rows=(fi.Din, bo.mul, bo.ir.dim), c_off=off, alpha=alpha))
                assert off not in seen
                seen.add(off)
    out.append(g)
    # post linear: dW[kk, w] = alpha' sum_{z, k} mid[z, k, kk] g_cv[z, k, w]
    post = conv.tp.linear
    g = None
    if needs[8 + n_fc]:
        g = torch.zeros_like(post.weight)
        for i, o, off, alpha in post.paths:
            bi, bo = fi.mid[i], fi.conv_out[o]
            probs.append(ops.wgrad_problem(mid, g_cv, g, N * bi.ir.dim, bi.mul, bo.mul, a_off=fi.m_off[i],
                                           a_rows=(fi.Dmid, bi.mul, bi.ir.dim), b_off=fi.c_off[o],
                                           b_rows=(fi.Dconv, bo.mul, bo.ir.dim), c_off=off, alpha=alpha * fi.inv_sqrt_avg))
    out.append(g)
    # self-connection: dW[u, v, w] = alpha sum_{z, m} x[z, m, u] a[z, v] g[z, m, w];  da[z, v] likewise
    sc = conv.sc
    g = torch.zeros_like(sc.weight) if needs[9 + n_fc] else None
    g_attrs = torch.zeros_like(attrs) if need_attrs else None
    V = fi.V
    if g is not None:
        for i1, i2, o, off, alpha in sc.paths:
            bi, bo = fi.feat_in[i1], fi.conv_out[o]
            probs.append(ops.wgrad_problem(x_imu, g_cv, g, N * bi.ir.dim, bi.mul, bo.mul, a_off=fi.x_off[i1],
                                           a_rows=(fi.Din, bi.mul, bi.ir.dim), b_off=fi.c_off[o],
                                           b_rows=(fi.Dconv, bo.mul, bo.ir.dim), aux=attrs, aux_d=bi.ir.dim, c_off=off,
                                           c_rows=(bo.mul, V * bo.mul, bi.mul), alpha=alpha))
    ops.wgrad_run(probs, dev)
    if g_attrs is not None:
        for i1, i2, o, off, alpha in sc.paths:
            bi, bo = fi.feat_in[i1], fi.conv_out[o]
            xa = x_imu[:, fi.x_off[i1]:fi.x_off[i1] + bi.dim].reshape(N, bi.ir.dim, bi.mul)
            gb = g_cv[:, fi.c_off[o]:fi.c_off[o] + bo.dim].reshape(N, bi.ir.dim, bo.mul)
            t = torch.einsum("zmu,zmw->zuw", xa, gb)                     # [z, u, w]
            W = sc.weight[off:off + bi.mul * V * bo.mul].reshape(bi.mul, V, bo.mul)
            g_attrs += alpha * torch.einsum("zuw,uvw->zv", t, W)
    out.append(g)
    return out, g_attrs


def interaction(fi, x, attrs, er, Y, csr, group=None, edge_index=None):
    """-> (out mul_ir, out imu).  `x` is the mul_ir feature tensor; if it carries the imu twin written by the
    previous block's gate (attribute ``_e3b_imu``) that one is consumed instead, so no layout pass runs.  `group`: the
    interaction blocks of the same network (their radial hidden layers run together, see `_RadialHidden`).  On a
    symmetric radius graph in evaluation mode the radial MLP runs once per UNDIRECTED edge (`undirected`)."""
    twin = getattr(x, "_e3b_imu", None)
    params = fi._block_params()
    E = er.shape[0]
    und = None
    if (SHARED_W and not ops.DETERMINISTIC and getattr(er, "_e3b_length_only", False) and edge_index is not None
            and not fi.mp.training and E > 0 and E % 2 == 0 and csr.out_eid is None
            and csr.in_eid is not None and fi.conv.tp.plan.specialized and fi.structure.uniform_mul in (32, 64)
            and fi.conv.fc.n_layers > 1 and fi.hs[-2] % 4 == 0):
        global SHARED_CALLS
        SHARED_CALLS += 1
        und = undirected(er, edge_index, csr)
    src = und.er_u if und is not None else er
    h_last = radial_hidden(src, fi, group) if fi.conv.fc.n_layers > 1 else er
    grp = ops.species_groups_of(attrs)
    if grp is not None:
        global SPECIES_SC_CALLS
        SPECIES_SC_CALLS += 1
    if twin is not None:
        return _Interaction.apply(None, twin, attrs, h_last, Y, fi, csr, und, grp, *params)
    return _Interaction.apply(x, None, attrs, h_last, Y, fi, csr, und, grp, *params)

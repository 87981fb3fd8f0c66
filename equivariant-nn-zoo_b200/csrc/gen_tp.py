#!/usr/bin/env python3
"""Build-time generator for the fully unrolled tensor-product convolution kernels
(csrc/tp_generated.cuh) and the dense CG tables of the generic kernel (csrc/cg_tables.cuh).

For every TP structure the reference's configs produce (e3b200.plan.reference_structures) it
emits straight-line CUDA: per warp "group" (a balanced subset of the input blocks) the sparse
Clebsch-Gordan contraction with the coefficients sqrt(2 l3+1) C_ijk as immediates, the outer
products x_i * Y_j shared between the paths of one (input block, SH block) pair, accumulators
in named registers, channel = lane.  Run by __graft_entry__.build(); output is committed too.
"""
import os
import sys
from math import sqrt

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from e3b200 import cg  # noqa: E402
from e3b200.plan import generated_structures  # noqa: E402

MAX_ACC_PER_GROUP = 56  # accumulator registers per thread (forward)
PAIRED_INFO = {}        # sid -> (d/dY partial rows of the paired backward kernel, selected by default)
N_PARTS = 4             # translation units the generated kernels are spread over (E3B_TP_PART)
PAIRED_MAX_ACC = 30     # output components per group of the paired backward kernels (two registers each)


def lit(v):
    return f"T({float(v)!r})"


def coef(l1, l2, l3):
    return sqrt(2 * l3 + 1) * cg.w3j(l1, l2, l3)


def make_groups(st, max_acc=None):
    """Partition input blocks into groups with <= max_acc (default MAX_ACC_PER_GROUP) output components, keeping
    blocks whole and balancing the FMA count.  Returns list of lists of input-block indices."""
    MAX_ACC_PER_GROUP = max_acc or globals()["MAX_ACC_PER_GROUP"]
    n_in = len(st.irreps_in)
    acc = [sum(p.ir_out.dim for p in st.paths if p.i_in == b) for b in range(n_in)]
    work = [sum((abs(coef(st.irreps_in[p.i_in].ir.l, st.irreps_sh[p.i_sh].ir.l, p.ir_out.l)) > 0).sum()
                for p in st.paths if p.i_in == b) for b in range(n_in)]
    total = sum(acc)
    n_groups = max(1, -(-total // MAX_ACC_PER_GROUP))
    while True:
        groups = [[] for _ in range(n_groups)]
        gacc = [0] * n_groups
        gwork = [0] * n_groups
        ok = True
        for b in sorted(range(n_in), key=lambda b: -work[b]):
            cands = [g for g in range(n_groups) if gacc[g] + acc[b] <= MAX_ACC_PER_GROUP]
            if not cands:
                ok = False
                break
            g = min(cands, key=lambda g: gwork[g])
            groups[g].append(b)
            gacc[g] += acc[b]
            gwork[g] += work[b]
        if ok:
            return [sorted(g) for g in groups if g]
        n_groups += 1


class Emitter:
    def __init__(self):
        self.lines = []

    def __call__(self, s=""):
        self.lines.append(s)

    def text(self):
        return "\n".join(self.lines) + "\n"


def emit_structure(E, sid, st):
    groups = make_groups(st)
    xoff, xdim = st.x_comp_offsets()
    soff, sdim = st.sh_comp_offsets()
    ybase, ykst, ydim = st.y_layout()
    n_paths = len(st.paths)
    G = len(groups)

    def pairs_of(blocks):
        """(b, s) pairs in order with their paths"""
        out = []
        for b in blocks:
            for s in range(len(st.irreps_sh)):
                ps = [pi for pi, p in enumerate(st.paths) if p.i_in == b and p.i_sh == s]
                if ps:
                    out.append((b, s, ps))
        return out

    def emit_bwd_body(blocks, xl, wl, yl, stage_gx=False, paired=False, pmap=None, xmap=None, gyname=None, xbase="0"):
            pmap = pmap or (lambda pi: pi)
            xmap = xmap or (lambda b, i: xoff[b] + i)
            gyname = gyname or (lambda pi, k: f"gy_{pi}_{k}")
            used_s = sorted({s for (b, s, ps) in pairs_of(blocks)})
            for s in used_s:
                for j in range(st.irreps_sh[s].ir.dim):
                    E(f"    const T Y_{s}_{j} = {yl(s, j)};")
                    E(f"    T gY_{s}_{j} = T(0);")
            for b in blocks:
                d1 = st.irreps_in[b].ir.dim
                for i in range(d1):
                    E(f"    const T x_{b}_{i} = {xl(b, i)};")
                for i in range(d1):
                    E(f"    T gx_{b}_{i} = T(0);")
                for (bb, s, ps) in pairs_of([b]):
                    l1, l2 = st.irreps_in[b].ir.l, st.irreps_sh[s].ir.l
                    need = set()
                    for pi in ps:
                        C = coef(l1, l2, st.paths[pi].ir_out.l)
                        for i in range(2 * l1 + 1):
                            for j in range(2 * l2 + 1):
                                if abs(C[i, j]).max() > 0:
                                    need.add((i, j))
                    need = sorted(need)
                    E("    {")
                    for (i, j) in need:
                        E(f"      const T xy_{i}_{j} = x_{b}_{i} * Y_{s}_{j};")
                        E(f"      T gxy_{i}_{j} = T(0);")
                    for pi in ps:
                        l3 = st.paths[pi].ir_out.l
                        C = coef(l1, l2, l3)
                        E(f"      {{ const T w_p = {wl(pi)}; T gw_p = T(0);")
                        for k in range(2 * l3 + 1):
                            terms = [(i, j, C[i, j, k]) for i in range(2 * l1 + 1) for j in range(2 * l2 + 1) if C[i, j, k] != 0]
                            if not terms:
                                continue
                            i, j, c = terms[0]
                            expr = f"{lit(c)} * xy_{i}_{j}"
                            for (i, j, c) in terms[1:]:
                                expr = f"fma_({lit(c)}, xy_{i}_{j}, {expr})"
                            E(f"        gw_p = fma_({gyname(pi, k)}, {expr}, gw_p);")
                            E(f"        {{ const T gt = w_p * {gyname(pi, k)};")
                            for (i, j, c) in terms:
                                E(f"          gxy_{i}_{j} = fma_({lit(c)}, gt, gxy_{i}_{j});")
                            E("        }")
                        E(f"        if (active) gwr[{pmap(pi)} * mul] = gw_p; }}")
                    for (i, j) in need:
                        E(f"      gx_{b}_{i} = fma_(gxy_{i}_{j}, Y_{s}_{j}, gx_{b}_{i});")
                        E(f"      gY_{s}_{j} = fma_(gxy_{i}_{j}, x_{b}_{i}, gY_{s}_{j});")
                    E("    }")
                E("    if (a.gx_edge != nullptr && active) {")
                if paired:
                    E(f"      T* __restrict__ gxr = reinterpret_cast<T*>(a.gx_edge + eid * ROW_X) + {xbase} + u;")
                else:
                    E("      T* __restrict__ gxr = a.gx_edge + eid * x_dim + u;")
                for i in range(d1):
                    E(f"      gxr[{xmap(b, i)} * mul] = gx_{b}_{i};")
                E("    }")
                if stage_gx and paired:   # node-reduction mode, paired kernels: 8-byte vector reductions straight from registers
                    E("    if (a.gx_node != nullptr) {")
                    for i in range(d1):
                        E(f"      red_add(gxn + {xmap(b, i)} * mul, gx_{b}_{i});")
                    E("    }")
                elif stage_gx:   # node-reduction mode: the edge's gradient row is staged in shared memory for one TMA reduce-add
                    E("    if (a.gx_node != nullptr) {")
                    for i in range(d1):
                        E(f"      gxs[{xoff[b] + i} * mul + u] = gx_{b}_{i};")
                    E("    }")
            E("    if (a.gsh != nullptr) {")
            E(f"      {'float' if paired else 'T'}* __restrict__ gsr = a.gsh + (eid * a.n_part + part) * a.sh_dim;")
            sfx = "2" if paired else ""
            if paired and sdim == 9:
                vals = ["0.f"] * 9
                for s in used_s:
                    for j in range(st.irreps_sh[s].ir.dim):
                        vals[soff[s] + j] = f"pair_sum(gY_{s}_{j})"
                E("      gsh_reduce_store9(gsr, lane, " + ", ".join(vals) + ");")
            if not paired and sdim == 9:      # one channel per thread: the same halving butterfly (host emulation: plain sums)
                vals = ["T(0)"] * 9
                for s in used_s:
                    for j in range(st.irreps_sh[s].ir.dim):
                        vals[soff[s] + j] = f"gY_{s}_{j}"
                E("      E3B_GSH_STORE9(gsr, lane, " + ", ".join(vals) + ");")
            for s in ([] if sdim == 9 else range(len(st.irreps_sh))):
                for j in range(st.irreps_sh[s].ir.dim):
                    if s in used_s:
                        E(f"      E3B_GSH_STORE{sfx}(gsr, {soff[s] + j}, gY_{s}_{j});")
                    else:
                        E(f"      E3B_GSH_ZERO{sfx}(gsr, {soff[s] + j});")
            E("    }")


    # ------------------------------------------------------------------ forward
    for g, blocks in enumerate(groups):
        E(f"template <typename T, int MUL> __device__ __forceinline__ void tpf_S{sid}_g{g}(const TpArgs<T>& a, int64_t node, int u, bool active) {{")
        E("  const int mul = MUL > 0 ? MUL : a.mul;")
        E(f"  const int64_t x_dim = (int64_t){xdim} * mul, w_dim = (int64_t){n_paths} * mul, y_dim = (int64_t){ydim} * mul;")
        accs = []
        for pi, p in enumerate(st.paths):
            if p.i_in in blocks:
                for k in range(p.ir_out.dim):
                    accs.append(f"acc_{pi}_{k}")
        for i in range(0, len(accs), 8):
            E("  T " + ", ".join(f"{n} = T(0)" for n in accs[i:i + 8]) + ";")
        used_s = sorted({s_ for (b_, s_, ps_) in pairs_of(blocks)})
        names = []          # (register name, load expression given pointers xr/wr/yr)
        for b_ in blocks:
            for i in range(st.irreps_in[b_].ir.dim):
                names.append((f"x_{b_}_{i}", f"ldg(xr + {xoff[b_] + i} * mul)"))
        for s_ in used_s:
            for j in range(st.irreps_sh[s_].ir.dim):
                names.append((f"Y_{s_}_{j}", f"ldg(yr + {soff[s_] + j})"))
        for pi, p in enumerate(st.paths):
            if p.i_in in blocks:
                names.append((f"w_{pi}", f"ldg(wr + {pi} * mul)"))
        E("  const int64_t e0 = a.in_ptr[node], e1 = a.in_ptr[node + 1];")
        E("  if (e0 < e1) {")
        E("    int64_t src = a.in_nbr[e0];")
        E("    int64_t eid = a.in_eid ? (int64_t)a.in_eid[e0] : e0;")
        E("    int64_t nsrc = src, neid = eid;")
        E("    if (e0 + 1 < e1) { nsrc = a.in_nbr[e0 + 1]; neid = a.in_eid ? (int64_t)a.in_eid[e0 + 1] : e0 + 1; }")
        E("    const T* __restrict__ xr = a.x + src * x_dim + u;")
        E("    const T* __restrict__ wr = a.w + eid * w_dim + u;")
        E("    const T* __restrict__ yr = a.sh + eid * a.sh_dim;")
        for n, ex in names:
            E(f"    T {n} = {ex};")
        E("    for (int64_t kk = e0; kk < e1; ++kk) {")
        E("      // software pipeline: issue the loads of edge kk+1 (and the indices of kk+2) before the math of kk")
        E("      const bool has_next = kk + 1 < e1;")
        for n, ex in names:
            E(f"      T n{n} = {n};")
        E("      if (has_next) {")
        E("        xr = a.x + nsrc * x_dim + u; wr = a.w + neid * w_dim + u; yr = a.sh + neid * a.sh_dim;")
        for n, ex in names:
            E(f"        n{n} = {ex};")
        E("        if (kk + 2 < e1) { nsrc = a.in_nbr[kk + 2]; neid = a.in_eid ? (int64_t)a.in_eid[kk + 2] : kk + 2; }")
        E("      }")
        for (b_, s_, ps) in pairs_of(blocks):
            l1, l2 = st.irreps_in[b_].ir.l, st.irreps_sh[s_].ir.l
            need = set()
            for pi in ps:
                C = coef(l1, l2, st.paths[pi].ir_out.l)
                for i in range(2 * l1 + 1):
                    for j in range(2 * l2 + 1):
                        if abs(C[i, j]).max() > 0:
                            need.add((i, j))
            E("      {")
            for (i, j) in sorted(need):
                E(f"        const T xy_{i}_{j} = x_{b_}_{i} * Y_{s_}_{j};")
            for pi in ps:
                l3 = st.paths[pi].ir_out.l
                C = coef(l1, l2, l3)
                for k in range(2 * l3 + 1):
                    terms = [(i, j, C[i, j, k]) for i in range(2 * l1 + 1) for j in range(2 * l2 + 1) if C[i, j, k] != 0]
                    if not terms:
                        continue
                    i, j, c = terms[0]
                    expr = f"{lit(c)} * xy_{i}_{j}"
                    for (i, j, c) in terms[1:]:
                        expr = f"fma_({lit(c)}, xy_{i}_{j}, {expr})"
                    E(f"        acc_{pi}_{k} = fma_(w_{pi}, {expr}, acc_{pi}_{k});")
            E("      }")
        for n, ex in names:
            E(f"      {n} = n{n};")
        E("    }")
        E("  }")
        E("  if (active) {")
        E("    T* __restrict__ yo = a.y + node * y_dim + u;")
        for pi, p in enumerate(st.paths):
            if p.i_in in blocks:
                for k in range(p.ir_out.dim):
                    E(f"    yo[{ybase[p.slot] + k * ykst[p.slot]} * mul] = acc_{pi}_{k};")
        E("  }")
        E("}")
        E()

    # ------------------------------------------------------------------ backward
    for g, blocks in enumerate(groups):
        E(f"template <typename T, int MUL> __device__ __forceinline__ void tpb_S{sid}_g{g}(const TpArgs<T>& a, int64_t node, int u, bool active, int part, int lane) {{")
        E("  const int mul = MUL > 0 ? MUL : a.mul;")
        E(f"  const int64_t x_dim = (int64_t){xdim} * mul, w_dim = (int64_t){n_paths} * mul, y_dim = (int64_t){ydim} * mul;")
        E("  const T* __restrict__ gyr = a.gy + node * y_dim + u;")
        for pi, p in enumerate(st.paths):
            if p.i_in in blocks:
                for k in range(p.ir_out.dim):
                    E(f"  const T gy_{pi}_{k} = active ? ldg(gyr + {ybase[p.slot] + k * ykst[p.slot]} * mul) : T(0);")
        E("  const int64_t e0 = a.in_ptr[node], e1 = a.in_ptr[node + 1];")
        E("  for (int64_t kk = e0; kk < e1; ++kk) {")
        E("    const int64_t src = a.in_nbr[kk];")
        E("    const int64_t eid = a.in_eid ? (int64_t)a.in_eid[kk] : kk;")
        E("    const T* __restrict__ xr = a.x + src * x_dim + u;")
        E("    const T* __restrict__ wr = a.w + eid * w_dim + u;")
        E("    const T* __restrict__ yr = a.sh + eid * a.sh_dim;")
        E("    T* __restrict__ gwr = a.gw + eid * w_dim + u;")
        emit_bwd_body(blocks, lambda b, i: f"ldg(xr + {xoff[b] + i} * mul)", lambda pi: f"ldg(wr + {pi} * mul)",
                      lambda s_, j: f"ldg(yr + {soff[s_] + j})")
        E("  }")
        E("}")
        E()

    # ------------------------------------------------------------------ pipelined forward (TMA bulk copies)
    # One CTA per destination node.  An elected thread streams, per incoming edge, the edge's
    # weight row and the source node's feature row into a TPP_STAGES-deep shared-memory ring with
    # cp.async.bulk (TMA engine, completion on an mbarrier); the warps (group x channel-chunk)
    # consume from shared memory, so no registers are tied up by loads in flight.
    E("#ifdef __CUDACC__")
    E(f"template <int MUL> __global__ void __launch_bounds__(32 * {G} * (MUL / 32)) tpfp_S{sid}(const TpArgs<float> a) {{")
    E("  typedef float T;")
    E("  asm volatile(\"griddepcontrol.launch_dependents;\" ::: \"memory\");   // a dependent launch (the tcgen05 GEMM that follows) may set itself up while this grid drains")
    E(f"  constexpr int G = {G}, ROW_W = {n_paths} * MUL, ROW_X = {xdim} * MUL, STAGE = ROW_W + ROW_X, SH_DIM = {sdim};")
    E("  constexpr int NT = 32 * G * (MUL / 32);")
    E("  extern __shared__ __align__(128) unsigned char tpp_smem[];")
    E("  float* stages = reinterpret_cast<float*>(tpp_smem);")
    E("  int* s_src = reinterpret_cast<int*>(stages + TPP_STAGES * STAGE);")
    E("  int* s_eid = s_src + TPP_MAXSEG;")
    E("  int* s_wid = s_eid + TPP_MAXSEG;    // row of the weight tensor per slot (= edge id unless a.w_idx shares rows)")
    E("  uint64_t* full = reinterpret_cast<uint64_t*>(s_wid + TPP_MAXSEG);")
    E("  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;")
    E("  const int64_t node = blockIdx.x;")
    E(f"  const int chunk = warp / G, group = warp - chunk * G;")
    E("  const int u = chunk * 32 + lane;")
    E("  if (tid == 0) {")
    E("    for (int s = 0; s < TPP_STAGES; ++s) mbar_init(&full[s], 1);")
    E("    fence_mbar_init();")
    E("  }")
    E("  const int64_t e0 = a.in_ptr[node], e1 = a.in_ptr[node + 1];")
    accs_all = []
    for pi, p in enumerate(st.paths):
        for k in range(p.ir_out.dim):
            accs_all.append(f"acc_{pi}_{k}")
    # accumulators: each warp only uses its group's, the compiler keeps one live set per branch
    E("  uint32_t it = 0;  // running edge counter -> ring stage and mbarrier phase")
    E("  switch (group) {")
    for g, blocks in enumerate(groups):
        E(f"  case {g}: {{")
        accs = [f"acc_{pi}_{k}" for pi, p in enumerate(st.paths) if p.i_in in blocks for k in range(p.ir_out.dim)]
        for i in range(0, len(accs), 8):
            E("    T " + ", ".join(f"{n} = T(0)" for n in accs[i:i + 8]) + ";")
        E("    for (int64_t c0 = e0; c0 < e1; c0 += TPP_MAXSEG) {")
        E("      const int n = (int)((e1 - c0) < TPP_MAXSEG ? (e1 - c0) : TPP_MAXSEG);")
        E("      __syncthreads();   // previous chunk fully consumed (also orders the barrier init)")
        E("      for (int i = tid; i < n; i += NT) {")
        E("        s_src[i] = a.in_nbr[c0 + i];")
        E("        const int eid_i = a.in_eid ? a.in_eid[c0 + i] : (int)(c0 + i);")
        E("        s_eid[i] = eid_i;")
        E("        s_wid[i] = a.w_idx ? a.w_idx[eid_i] : eid_i;")
        E("      }")
        E("      __syncthreads();")
        E("      if (tid == 0) {")
        E("        const int pre = n < TPP_STAGES ? n : TPP_STAGES;")
        E("        for (int j = 0; j < pre; ++j) {")
        E("          const uint32_t s = (it + j) % TPP_STAGES;")
        E("          tpp_issue(stages + s * STAGE, &full[s], a.w + (int64_t)s_wid[j] * ROW_W, ROW_W, a.x + (int64_t)s_src[j] * ROW_X, ROW_X);")
        E("        }")
        E("      }")
        E("      T Ycur = (lane < SH_DIM) ? ldg(a.sh + (int64_t)s_eid[0] * SH_DIM + lane) : T(0);")
        E("      for (int i = 0; i < n; ++i, ++it) {")
        E("        const uint32_t s = it % TPP_STAGES, ph = (it / TPP_STAGES) & 1u;")
        E("        T Ynext = T(0);")
        E("        if (i + 1 < n && lane < SH_DIM) Ynext = ldg(a.sh + (int64_t)s_eid[i + 1] * SH_DIM + lane);")
        E("        mbar_wait(&full[s], ph);")
        E("        const T* __restrict__ sw = stages + s * STAGE + u;")
        E("        const T* __restrict__ sx = sw + ROW_W;")
        used_s = sorted({s_ for (b_, s_, ps_) in pairs_of(blocks)})
        for s_ in used_s:
            for j in range(st.irreps_sh[s_].ir.dim):
                E(f"        const T Y_{s_}_{j} = __shfl_sync(0xffffffffu, Ycur, {soff[s_] + j});")
        for b_ in blocks:
            for i in range(st.irreps_in[b_].ir.dim):
                E(f"        const T x_{b_}_{i} = sx[{xoff[b_] + i} * MUL];")
        for (b_, s_, ps) in pairs_of(blocks):
            l1, l2 = st.irreps_in[b_].ir.l, st.irreps_sh[s_].ir.l
            need = set()
            for pi in ps:
                C = coef(l1, l2, st.paths[pi].ir_out.l)
                for i in range(2 * l1 + 1):
                    for j in range(2 * l2 + 1):
                        if abs(C[i, j]).max() > 0:
                            need.add((i, j))
            E("        {")
            for (i, j) in sorted(need):
                E(f"          const T xy_{i}_{j} = x_{b_}_{i} * Y_{s_}_{j};")
            for pi in ps:
                l3 = st.paths[pi].ir_out.l
                C = coef(l1, l2, l3)
                E(f"          {{ const T w_p = sw[{pi} * MUL];")
                for k in range(2 * l3 + 1):
                    terms = [(i, j, C[i, j, k]) for i in range(2 * l1 + 1) for j in range(2 * l2 + 1) if C[i, j, k] != 0]
                    if not terms:
                        continue
                    i, j, c = terms[0]
                    expr = f"{lit(c)} * xy_{i}_{j}"
                    for (i, j, c) in terms[1:]:
                        expr = f"fma_({lit(c)}, xy_{i}_{j}, {expr})"
                    E(f"            acc_{pi}_{k} = fma_(w_p, {expr}, acc_{pi}_{k});")
                E("          }")
            E("        }")
        E("        Ycur = Ynext;")
        E("        __syncthreads();   // every warp is done with stage s")
        E("        if (tid == 0 && i + TPP_STAGES < n)")
        E("          tpp_issue(stages + s * STAGE, &full[s], a.w + (int64_t)s_wid[i + TPP_STAGES] * ROW_W, ROW_W,")
        E("                    a.x + (int64_t)s_src[i + TPP_STAGES] * ROW_X, ROW_X);")
        E("      }")
        E("    }")
        E(f"    T* __restrict__ yo = a.y + node * ({ydim} * MUL) + u;")
        for pi, p in enumerate(st.paths):
            if p.i_in in blocks:
                for k in range(p.ir_out.dim):
                    E(f"    yo[{ybase[p.slot] + k * ykst[p.slot]} * MUL] = acc_{pi}_{k};")
        E("  } break;")
    E("  }")
    E("}")
    E(f"static size_t tpfp_smem_S{sid}(int mul) {{ return (size_t)TPP_STAGES * ({n_paths} + {xdim}) * mul * 4 + 3 * TPP_MAXSEG * 4 + TPP_STAGES * 8; }}")
    # ---- pipelined backward: same ring; gy of the node lives in registers, gw / gx_edge / gsh
    # partials are stored directly (fire-and-forget)
    E(f"template <int MUL> __global__ void __launch_bounds__(32 * {G} * (MUL / 32)) tpbp_S{sid}(const TpArgs<float> a) {{")
    E("  typedef float T;")
    E("  asm volatile(\"griddepcontrol.launch_dependents;\" ::: \"memory\");   // a dependent launch (the tcgen05 GEMM that follows) may set itself up while this grid drains")
    E(f"  constexpr int G = {G}, ROW_W = {n_paths} * MUL, ROW_X = {xdim} * MUL, STAGE = ROW_W + ROW_X, SH_DIM = {sdim};")
    E("  constexpr int NT = 32 * G * (MUL / 32);")
    E("  constexpr int mul = MUL; constexpr bool active = true;")
    E("  constexpr int64_t x_dim = ROW_X, w_dim = ROW_W;")
    E("  extern __shared__ __align__(128) unsigned char tpp_smem[];")
    E("  float* stages = reinterpret_cast<float*>(tpp_smem);")
    E("  int* s_src = reinterpret_cast<int*>(stages + TPP_STAGES * STAGE);")
    E("  int* s_eid = s_src + TPP_MAXSEG;")
    E("  int* s_wid = s_eid + TPP_MAXSEG;    // row of the weight tensor per slot (= edge id unless a.w_idx shares rows)")
    E("  uint64_t* full = reinterpret_cast<uint64_t*>(s_wid + TPP_MAXSEG);")
    E("  float* gx_stage = reinterpret_cast<float*>(full + TPP_STAGES + (TPP_STAGES & 1));   // 16-byte aligned; only with a.gx_node")
    E("  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;")
    E("  const int64_t node = blockIdx.x;")
    E(f"  const int chunk = warp / G, group = warp - chunk * G, part = warp;")
    E("  const int u = chunk * 32 + lane;")
    E("  if (tid == 0) {")
    E("    for (int s = 0; s < TPP_STAGES; ++s) mbar_init(&full[s], 1);")
    E("    fence_mbar_init();")
    E("  }")
    E("  const int64_t e0 = a.in_ptr[node], e1 = a.in_ptr[node + 1];")
    E(f"  const T* __restrict__ gyr = a.gy + node * ({ydim} * MUL) + u;")
    E("  uint32_t it = 0;")
    E("  switch (group) {")
    for g, blocks in enumerate(groups):
        E(f"  case {g}: {{")
        for pi, p in enumerate(st.paths):
            if p.i_in in blocks:
                for k in range(p.ir_out.dim):
                    E(f"    const T gy_{pi}_{k} = ldg(gyr + {ybase[p.slot] + k * ykst[p.slot]} * MUL);")
        E("    for (int64_t c0 = e0; c0 < e1; c0 += TPP_MAXSEG) {")
        E("      const int n = (int)((e1 - c0) < TPP_MAXSEG ? (e1 - c0) : TPP_MAXSEG);")
        E("      __syncthreads();")
        E("      for (int i = tid; i < n; i += NT) {")
        E("        s_src[i] = a.in_nbr[c0 + i];")
        E("        const int eid_i = a.in_eid ? a.in_eid[c0 + i] : (int)(c0 + i);")
        E("        s_eid[i] = eid_i;")
        E("        s_wid[i] = a.w_idx ? a.w_idx[eid_i] : eid_i;")
        E("      }")
        E("      __syncthreads();")
        E("      if (tid == 0) {")
        E("        const int pre = n < TPP_STAGES ? n : TPP_STAGES;")
        E("        for (int j = 0; j < pre; ++j) {")
        E("          const uint32_t s = (it + j) % TPP_STAGES;")
        E("          tpp_issue(stages + s * STAGE, &full[s], a.w + (int64_t)s_wid[j] * ROW_W, ROW_W, a.x + (int64_t)s_src[j] * ROW_X, ROW_X);")
        E("        }")
        E("      }")
        E("      T Ycur = (lane < SH_DIM) ? ldg(a.sh + (int64_t)s_eid[0] * SH_DIM + lane) : T(0);")
        E("      for (int i = 0; i < n; ++i, ++it) {")
        E("        const uint32_t s = it % TPP_STAGES, ph = (it / TPP_STAGES) & 1u;")
        E("        const int64_t eid = s_eid[i];")
        E("        T Ynext = T(0);")
        E("        if (i + 1 < n && lane < SH_DIM) Ynext = ldg(a.sh + (int64_t)s_eid[i + 1] * SH_DIM + lane);")
        E("        mbar_wait(&full[s], ph);")
        E("        const T* __restrict__ sw = stages + s * STAGE + u;")
        E("        const T* __restrict__ sx = sw + ROW_W;")
        E("        T* __restrict__ gwr = a.gw + eid * w_dim + u;")
        E("        float* gxs = gx_stage + (it % TPP_GXBUF) * ROW_X;")
        E("        {")
        emit_bwd_body(blocks, lambda b, i: f"sx[{xoff[b] + i} * MUL]", lambda pi: f"sw[{pi} * MUL]",
                      lambda s_, j: f"__shfl_sync(0xffffffffu, Ycur, {soff[s_] + j})", stage_gx=True)
        E("        }")
        E("        if (a.gx_node != nullptr) {")
        E("          fence_proxy_async();                       // the staged row -> visible to the TMA engine")
        E("          if (tid == 0) bulk_wait_read<TPP_GXBUF - 2>();   // the buffer of the NEXT edge has been read out")
        E("        }")
        E("        Ycur = Ynext;")
        E("        __syncthreads();")
        E("        if (tid == 0) {")
        E("          if (a.gx_node != nullptr) {                // d/dx of this edge += into its source node's row (L2 reduction)")
        E("            bulk_reduce_add_f32(a.gx_node + (int64_t)s_src[i] * ROW_X, gxs, ROW_X * 4u);")
        E("            bulk_commit();")
        E("          }")
        E("          if (i + TPP_STAGES < n)")
        E("            tpp_issue(stages + s * STAGE, &full[s], a.w + (int64_t)s_wid[i + TPP_STAGES] * ROW_W, ROW_W,")
        E("                      a.x + (int64_t)s_src[i + TPP_STAGES] * ROW_X, ROW_X);")
        E("        }")
        E("      }")
        E("    }")
        E("  } break;")
    E("  }")
    E("  if (a.gx_node != nullptr && tid == 0) bulk_wait_all();   // the staging buffers must outlive the reductions reading them")
    E("}")
    E(f"static size_t tpbp_smem_S{sid}(int mul, bool reduce) {{ return ((tpfp_smem_S{sid}(mul) + 15) & ~(size_t)15) + 16 + (reduce ? (size_t)TPP_GXBUF * {xdim} * mul * 4 : 0); }}")
    # ---- paired pipelined kernels (multiplicity 64): a thread owns the channels (2 lane, 2 lane + 1) and computes on packed
    # fp32 pairs (FFMA2 / FMUL2: half the issue slots of the arithmetic, 8-byte shared-memory loads and global stores); the
    # input blocks are split into more groups (warps) so that the doubled live set still fits the register file.  The warps of
    # a CTA are DECOUPLED: a stage of the ring has a `full` barrier (TMA bytes) and an `empty` barrier (one arrival per
    # warp); a warp waits for data only, runs ahead of the others by up to the ring depth, and adds ITS input blocks'
    # d/dx rows into the source node's row with 8-byte vector reductions (red.global.add.v2.f32) from registers -- no
    # CTA-wide barrier per edge, no staging.  The lightest warp's lane 0
    # is the producer: one edge behind itself it waits for `empty` and refills the stage.
    acc_of = [sum(p.ir_out.dim for p in st.paths if p.i_in == b) for b in range(len(st.irreps_in))]
    for cap in (PAIRED_MAX_ACC, 36, 42, MAX_ACC_PER_GROUP):
        groups2 = make_groups(st, max(cap, max(acc_of)))
        G2 = len(groups2)
        if G2 <= 2 * G:     # the partial-sum rows of d/dY are laid out for G * 2 warps per node
            break
    else:
        groups2, G2 = groups, G
    work_of = [sum(int((abs(coef(st.irreps_in[p.i_in].ir.l, st.irreps_sh[p.i_sh].ir.l, p.ir_out.l)) > 0).sum())
                   for p in st.paths if p.i_in == b) for b in range(len(st.irreps_in))]
    g_light = min(range(G2), key=lambda g: sum(work_of[b] for b in groups2[g]))   # the warp with the least arithmetic

    def paired_head(name, bwd, nwarps=None, pw=None, pair=True, ngroups=None):
        nwarps = G2 if nwarps is None else nwarps
        pw = g_light if pw is None else pw
        E(f"__global__ void __launch_bounds__(32 * {nwarps}) {name}_S{sid}(const TpArgs<float> a) {{")
        E("  typedef F2 T;" if pair else "  typedef float T;")
        E("  asm volatile(\"griddepcontrol.launch_dependents;\" ::: \"memory\");")
        E(f"  constexpr int MUL = 64, G = {nwarps}, PW = {pw}, ROW_W = {n_paths} * MUL, ROW_X = {xdim} * MUL, STAGE = ROW_W + ROW_X, SH_DIM = {sdim};")
        E("  constexpr int NT = 32 * G;")
        E("  constexpr int mul = MUL / 2;   // row strides below are in channel PAIRS" if pair else "  constexpr int mul = MUL;")
        if bwd:
            E("  constexpr bool active = true;")
        E("  extern __shared__ __align__(128) unsigned char tpp_smem[];")
        E("  float* stages = reinterpret_cast<float*>(tpp_smem);")
        E("  const uint32_t nst = (uint32_t)a.n_stages;   // ring depth: a launch parameter")
        E("  int* s_src = reinterpret_cast<int*>(stages + nst * STAGE);")
        E("  int* s_eid = s_src + TPP2_MAXSEG;")
        E("  int* s_wid = s_eid + TPP2_MAXSEG;")
        E("  uint64_t* full = reinterpret_cast<uint64_t*>(s_wid + TPP2_MAXSEG);")
        E("  uint64_t* empty = full + nst;")
        E("  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;")
        E("  const int64_t node = blockIdx.x;")
        if pair:
            E("  const int group = warp, u = lane;" + (" const int part = warp;" if bwd else ""))
        else:
            E(f"  const int chunk = warp / {ngroups}, group = warp - chunk * {ngroups}, u = chunk * 32 + lane;" + (" const int part = warp;" if bwd else ""))
        E("  if (tid == 0) {")
        E("    for (uint32_t s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], G); }")
        E("    fence_mbar_init();")
        E("  }")
        E("  const int64_t e0 = a.in_ptr[node], e1 = a.in_ptr[node + 1];")
        E("  uint32_t it = 0, s = 0, ph = 0;   // running edge counter, ring stage and barrier phase of the edge being consumed")

    def paired_chunk_open(pair=True):
        E("    for (int64_t c0 = e0; c0 < e1; c0 += TPP2_MAXSEG) {")
        E("      const int n = (int)((e1 - c0) < TPP2_MAXSEG ? (e1 - c0) : TPP2_MAXSEG);")
        E("      __syncthreads();   // every warp is through the previous chunk: all stages free (also orders the barrier init)")
        E("      for (int i = tid; i < n; i += NT) {")
        E("        s_src[i] = a.in_nbr[c0 + i];")
        E("        const int eid_i = a.in_eid ? a.in_eid[c0 + i] : (int)(c0 + i);")
        E("        s_eid[i] = eid_i;")
        E("        s_wid[i] = a.w_idx ? a.w_idx[eid_i] : eid_i;")
        E("      }")
        E("      __syncthreads();")
        E("      if (warp == PW && lane == 0) {")
        E("        const int pre = n < (int)nst ? n : (int)nst;")
        E("        for (int j = 0; j < pre; ++j) {")
        E("          uint32_t sj = s + j; if (sj >= nst) sj -= nst;")
        E("          tpp_issue(stages + sj * STAGE, &full[sj], a.w + (int64_t)s_wid[j] * ROW_W, ROW_W, a.x + (int64_t)s_src[j] * ROW_X, ROW_X);")
        E("        }")
        E("      }")
        E("      uint32_t sl = s, phl = ph;   // stage / phase of the edge TPP2_LAG behind: the one the producer refills")
        E("      float Ycur = (lane < SH_DIM) ? ldg(a.sh + (int64_t)s_eid[0] * SH_DIM + lane) : 0.f;")
        E("      for (int i = 0; i < n; ++i, ++it) {")
        E("        if (i >= TPP2_LAG) {")
        E("          if (warp == PW && lane == 0 && i - TPP2_LAG + (int)nst < n) {   // refill the stage of edge i - LAG once every warp released it")
        E("            mbar_wait(&empty[sl], phl);")
        E("            tpp_issue(stages + sl * STAGE, &full[sl], a.w + (int64_t)s_wid[i - TPP2_LAG + nst] * ROW_W, ROW_W,")
        E("                      a.x + (int64_t)s_src[i - TPP2_LAG + nst] * ROW_X, ROW_X);")
        E("          }")
        E("          if (++sl == nst) { sl = 0; phl ^= 1u; }")
        E("        }")
        E("        float Ynext = 0.f;")
        E("        if (i + 1 < n && lane < SH_DIM) Ynext = ldg(a.sh + (int64_t)s_eid[i + 1] * SH_DIM + lane);")
        E("        mbar_wait(&full[s], ph);")
        E("        const T* __restrict__ sw = reinterpret_cast<const T*>(stages + s * STAGE) + u;")
        E("        const T* __restrict__ sx = sw + ROW_W / 2;" if pair else "        const T* __restrict__ sx = sw + ROW_W;")

    def paired_chunk_close():
        E("        Ycur = Ynext;")
        E("        __syncwarp();")
        E("        if (lane == 0) mbar_arrive(&empty[s]);   // this warp is done with the stage")
        E("        if (++s == nst) { s = 0; ph ^= 1u; }")
        E("      }")
        E("    }")

    # ---------------- backward
    # Class-shared variant: one warp per input block, and blocks of the same l with the same relative path list (the two
    # parities of an l in the hidden layers) run THE SAME unrolled code with per-warp base offsets -- the hot loops of a
    # kernel then total ~14 KB instead of 27 KB and stay inside the SM's 32 KB instruction cache (ncu: `no_instruction`
    # 2.2 -> 0.2 stall cycles per issue).  Needs: every block has paths (contiguous in the weight row, by construction of
    # the plan) -- the restricted structures keep the group variant.  MEASURED SLOWER than one channel per thread and
    # therefore not the default (E3B_TP_PAIRED_FORCE=1 runs it): 160-192 threads x 164 registers leave 2 CTAs per SM, and
    # with one CTA per destination node the start-up chain of a node (in_ptr -> indices -> TMA -> first data) is hidden
    # only by the other resident CTAs: 27 paths 609 -> 773 us, 15 paths 327 -> 413 us; ring depth 3 -> 8 recovers 5 %.
    # What it needs next is a persistent CTA over a flattened edge range.
    blk_paths = [[pi for pi, p in enumerate(st.paths) if p.i_in == b] for b in range(len(st.irreps_in))]
    sig = [(st.irreps_in[b].ir.l, tuple((st.paths[pi].i_sh, st.paths[pi].ir_out.l) for pi in blk_paths[b])) for b in range(len(st.irreps_in))]
    class_ok = (all(len(ps) > 0 and ps == list(range(ps[0], ps[0] + len(ps))) for ps in blk_paths)
                and 2 <= len(st.irreps_in) <= 8 and sdim == 9 and n_paths > 3)
    if class_ok:
        order = sorted(range(len(st.irreps_in)), key=lambda b: -work_of[b])      # warp w runs block order[w]: heavy blocks first
        classes = []
        for b in order:
            if sig[b] not in [sig[c] for c in classes]:
                classes.append(b)                                                # representative block of each class
        cls_of = [[sig[c] for c in classes].index(sig[b]) for b in order]
        W3 = len(order)
        n_gy = [sum(st.paths[pi].ir_out.dim for pi in blk_paths[b]) for b in order]
        E(f"static __device__ const int kB3Cls_S{sid}[{W3}] = {{" + ", ".join(map(str, cls_of)) + "};")
        E(f"static __device__ const int kB3Xb_S{sid}[{W3}] = {{" + ", ".join(str(xoff[b]) for b in order) + "};")
        E(f"static __device__ const int kB3Wb_S{sid}[{W3}] = {{" + ", ".join(str(blk_paths[b][0]) for b in order) + "};")
        E(f"static __device__ const int kB3Gy_S{sid}[{W3}][{max(n_gy)}] = {{")
        for b in order:
            offs = [ybase[st.paths[pi].slot] + k * ykst[st.paths[pi].slot] for pi in blk_paths[b] for k in range(st.paths[pi].ir_out.dim)]
            E("  {" + ", ".join(map(str, offs + [0] * (max(n_gy) - len(offs)))) + "},")
        E("};")
        pw3 = W3 - 1                                                             # the lightest block's warp is the producer
        paired_head("tpbp2", True, W3, pw3)
        E(f"  const int xb = kB3Xb_S{sid}[warp] * mul, wb = kB3Wb_S{sid}[warp] * mul;   // this warp's block: offsets in channel pairs")
        E(f"  const int* __restrict__ go = kB3Gy_S{sid}[warp];")
        E(f"  const T* __restrict__ gyr = reinterpret_cast<const T*>(a.gy + node * ({ydim} * MUL)) + u;")
        E(f"  switch (kB3Cls_S{sid}[warp]) {{")
        for ci, b in enumerate(classes):
            p0 = blk_paths[b][0]
            E(f"  case {ci}: {{")
            j = 0
            for pi in blk_paths[b]:
                for k in range(st.paths[pi].ir_out.dim):
                    E(f"    const T gy_{pi}_{k} = ldg(gyr + go[{j}] * mul);")
                    j += 1
            paired_chunk_open()
            E("        sw += wb; sx += xb;")
            E("        const int64_t eid = s_eid[i];")
            E("        T* __restrict__ gwr = reinterpret_cast<T*>(a.gw + eid * ROW_W) + wb + u;")
            E("        T* __restrict__ gxn = reinterpret_cast<T*>(a.gx_node + (int64_t)s_src[i] * ROW_X) + xb + u;   // d/dx of the SOURCE node (node-reduction mode)")
            E("        {")
            emit_bwd_body([b], lambda b_, i: f"sx[{i} * mul]", lambda pi: f"sw[{pi - p0} * mul]",
                          lambda s_, j_: f"T(__shfl_sync(0xffffffffu, Ycur, {soff[s_] + j_}))", stage_gx=True, paired=True,
                          pmap=lambda pi: pi - p0, xmap=lambda b_, i: i, xbase="xb")
            E("        }")
            paired_chunk_close()
            E("  } break;")
        E("  }")
        E("}")
        E(f"static const int kPairedBwdWarps_S{sid} = {W3}, kPairedBwdParts_S{sid} = {W3};")
    else:
        E(f"static const int kPairedBwdWarps_S{sid} = {G2}, kPairedBwdParts_S{sid} = {2 * G};")
    if not class_ok:
        paired_head("tpbp2", True)
        E(f"  const T* __restrict__ gyr = reinterpret_cast<const T*>(a.gy + node * ({ydim} * MUL)) + u;")
        E("  switch (group) {")
        for g, blocks in enumerate(groups2):
            E(f"  case {g}: {{")
            for pi, p in enumerate(st.paths):
                if p.i_in in blocks:
                    for k in range(p.ir_out.dim):
                        E(f"    const T gy_{pi}_{k} = ldg(gyr + {ybase[p.slot] + k * ykst[p.slot]} * mul);")
            paired_chunk_open()
            E("        const int64_t eid = s_eid[i];")
            E("        T* __restrict__ gwr = reinterpret_cast<T*>(a.gw + eid * ROW_W) + u;")
            E("        T* __restrict__ gxn = reinterpret_cast<T*>(a.gx_node + (int64_t)s_src[i] * ROW_X) + u;   // d/dx of the SOURCE node (node-reduction mode)")
            E("        {")
            emit_bwd_body(blocks, lambda b, i: f"sx[{xoff[b] + i} * mul]", lambda pi: f"sw[{pi} * mul]",
                          lambda s_, j: f"T(__shfl_sync(0xffffffffu, Ycur, {soff[s_] + j}))", stage_gx=True, paired=True)
            E("        }")
            if g == g_light and G2 != 2 * G:
                E(f"        if (a.gsh != nullptr && lane < SH_DIM)     // partial-sum rows no warp of this kernel owns")
                E(f"          for (int p = G; p < a.n_part; ++p) a.gsh[(eid * a.n_part + p) * SH_DIM + lane] = 0.f;")
            paired_chunk_close()
            E("  } break;")
        E("  }")
        E("}")
    E(f"static const int kPairedGroups_S{sid} = {G2};")
    # which variant runs by default -- measured per kernel on the W2 step (ncu launch list, paired vs one channel per thread):
    # backward 3-path structures 133 -> 88 us and 188 -> 171 us, the 15-path restriction of the full block 405 -> 371 us, but the
    # 15-path second block 327 -> 338 us and 27 paths 609 -> 646 us (the unrolled loops of the 4 warps, 27 KB, no longer fit the 32 KB instruction
    # cache of the SM: ncu `no_instruction` is the top stall); forward 27 / 30 paths 258 -> 227 us and 260 -> 248 us, smaller
    # structures unchanged.  E3B_TP_PAIRED_FORCE=1 runs the paired kernels everywhere.
    bwd_ok = not class_ok and (n_paths <= 3 or (n_paths == 15 and len(st.irreps_in) == 6))
    PAIRED_INFO[sid] = (W3 if class_ok else 2 * G, bwd_ok)
    E(f"static const bool kPairedBwdOk_S{sid} = {'true' if bwd_ok else 'false'};")
    E(f"static const bool kPairedFwdOk_S{sid} = {'true' if n_paths >= 20 else 'false'};")
    E(f"static size_t tpp2_smem_S{sid}(int nst) {{ return (size_t)nst * ({n_paths} + {xdim}) * 64 * 4 + 3 * TPP2_MAXSEG * 4 + (size_t)nst * 16; }}")

    # ---------------- backward, one channel per thread, with the decoupled pipeline of the paired kernels (no CTA barrier per
    # edge, d/dx by fire-and-forget reductions from registers instead of the staged TMA reduce-add): the default for the
    # structures whose paired variant is slower
    pw1 = min(range(G), key=lambda g: sum(work_of[b] for b in groups[g]))
    paired_head("tpbp1d", True, 2 * G, pw1, pair=False, ngroups=G)
    E(f"  const T* __restrict__ gyr = a.gy + node * ({ydim} * MUL) + u;")
    E("  switch (group) {")
    for g, blocks in enumerate(groups):
        E(f"  case {g}: {{")
        for pi, p in enumerate(st.paths):
            if p.i_in in blocks:
                for k in range(p.ir_out.dim):
                    E(f"    const T gy_{pi}_{k} = ldg(gyr + {ybase[p.slot] + k * ykst[p.slot]} * mul);")
        paired_chunk_open(pair=False)
        E("        const int64_t eid = s_eid[i];")
        E("        T* __restrict__ gwr = a.gw + eid * ROW_W + u;")
        E("        T* __restrict__ gxn = a.gx_node + (int64_t)s_src[i] * ROW_X + u;   // d/dx of the SOURCE node (node-reduction mode)")
        E("        {")
        emit_bwd_body(blocks, lambda b, i: f"sx[{xoff[b] + i} * mul]", lambda pi: f"sw[{pi} * mul]",
                      lambda s_, j: f"__shfl_sync(0xffffffffu, Ycur, {soff[s_] + j})", stage_gx=True, paired=True)
        E("        }")
        paired_chunk_close()
        E("  } break;")
    E("  }")
    E("}")

    # ---------------- forward
    paired_head("tpfp2", False)
    E("  switch (group) {")
    for g, blocks in enumerate(groups2):
        E(f"  case {g}: {{")
        accs = [f"acc_{pi}_{k}" for pi, p in enumerate(st.paths) if p.i_in in blocks for k in range(p.ir_out.dim)]
        for i in range(0, len(accs), 8):
            E("    T " + ", ".join(f"{n} = T(0.f)" for n in accs[i:i + 8]) + ";")
        paired_chunk_open()
        used_s = sorted({s_ for (b_, s_, ps_) in pairs_of(blocks)})
        for s_ in used_s:
            for j in range(st.irreps_sh[s_].ir.dim):
                E(f"        const T Y_{s_}_{j} = T(__shfl_sync(0xffffffffu, Ycur, {soff[s_] + j}));")
        for b_ in blocks:
            for i in range(st.irreps_in[b_].ir.dim):
                E(f"        const T x_{b_}_{i} = sx[{xoff[b_] + i} * mul];")
        for (b_, s_, ps) in pairs_of(blocks):
            l1, l2 = st.irreps_in[b_].ir.l, st.irreps_sh[s_].ir.l
            need = set()
            for pi in ps:
                C = coef(l1, l2, st.paths[pi].ir_out.l)
                for i in range(2 * l1 + 1):
                    for j in range(2 * l2 + 1):
                        if abs(C[i, j]).max() > 0:
                            need.add((i, j))
            E("        {")
            for (i, j) in sorted(need):
                E(f"          const T xy_{i}_{j} = x_{b_}_{i} * Y_{s_}_{j};")
            for pi in ps:
                l3 = st.paths[pi].ir_out.l
                C = coef(l1, l2, l3)
                E(f"          {{ const T w_p = sw[{pi} * mul];")
                for k in range(2 * l3 + 1):
                    terms = [(i, j, C[i, j, k]) for i in range(2 * l1 + 1) for j in range(2 * l2 + 1) if C[i, j, k] != 0]
                    if not terms:
                        continue
                    i, j, c = terms[0]
                    expr = f"{lit(c)} * xy_{i}_{j}"
                    for (i, j, c) in terms[1:]:
                        expr = f"fma_({lit(c)}, xy_{i}_{j}, {expr})"
                    E(f"            acc_{pi}_{k} = fma_(w_p, {expr}, acc_{pi}_{k});")
                E("          }")
            E("        }")
        paired_chunk_close()
        E(f"    T* __restrict__ yo = reinterpret_cast<T*>(a.y + node * ({ydim} * MUL)) + u;")
        for pi, p in enumerate(st.paths):
            if p.i_in in blocks:
                for k in range(p.ir_out.dim):
                    E(f"    yo[{ybase[p.slot] + k * ykst[p.slot]} * mul] = acc_{pi}_{k};")
        E("  } break;")
    E("  }")
    E("}")
    E("#endif  // __CUDACC__")
    E()

    # ------------------------------------------------------------------ kernels
    E("#ifdef __CUDACC__")
    for kind in ("f", "b"):
        extra = ", part, lane" if kind == "b" else ""
        E(f"template <typename T, int MUL> __global__ void __launch_bounds__(TP_THREADS) tp{kind}_S{sid}(const TpArgs<T> a) {{")
        E("  const int lane = threadIdx.x & 31;")
        E("  const int64_t item = (int64_t)blockIdx.x * (TP_THREADS / 32) + (threadIdx.x >> 5);")
        E(f"  const int per_node = a.n_chunks * {G};")
        E("  if (item >= a.n_nodes * per_node) return;")
        E("  const int64_t node = item / per_node;")
        E("  const int part = (int)(item - node * per_node);")
        E(f"  const int chunk = part / {G}, group = part - chunk * {G};")
        E("  int u = chunk * 32 + lane;")
        E("  const bool active = u < a.mul;")
        E("  if (!active) u = a.mul - 1;")
        E("  switch (group) {")
        for g in range(G):
            E(f"    case {g}: tp{kind}_S{sid}_g{g}<T, MUL>(a, node, u, active{extra}); break;")
        E("  }")
        E("}")
        E()
    E("#endif  // __CUDACC__")
    return G


def emit_launchers(E, sid, G_):
    """host launchers of structure `sid` (G_ warp groups): pick the kernel variant and its launch geometry"""
    Gs = {sid: G_}
    for kind in ("f", "b"):
        E(f"void launch_tp{kind}_S{sid}(const TpArgs<float>& a, int64_t grid, cudaStream_t s) {{")
        if kind == "b":
            E(f"  if ((a.mul == 64 || a.mul == 32) && e3b_tp_pipelined_enabled()) {{")
            E(f"    const size_t smem = tpbp_smem_S{sid}(a.mul, a.gx_node != nullptr);")
            E(f"    if (a.mul == 64 && e3b_tp_paired_enabled(kPairedBwdOk_S{sid})) {{")
            E(f"      TpArgs<float> a2 = a; a2.n_stages = e3b_tp_stages(1, tpp2_smem_S{sid}(1) - tpp2_smem_S{sid}(0), tpp2_smem_S{sid}(0));")
            E(f"      const size_t smem2 = tpp2_smem_S{sid}(a2.n_stages);")
            E(f"      if (smem2 > 48 * 1024) cudaFuncSetAttribute(tpbp2_S{sid}, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);")
            E(f"      a2.n_part = kPairedBwdParts_S{sid};")
            E(f"      tpbp2_S{sid}<<<(unsigned)a.n_nodes, 32 * kPairedBwdWarps_S{sid}, smem2, s>>>(a2); }}")
            E(f"    else if (a.mul == 64 && e3b_tp_decoupled_enabled()) {{")
            E(f"      TpArgs<float> a2 = a; a2.n_stages = e3b_tp_stages(1, tpp2_smem_S{sid}(1) - tpp2_smem_S{sid}(0), tpp2_smem_S{sid}(0));")
            E(f"      const size_t smem2 = tpp2_smem_S{sid}(a2.n_stages);")
            E(f"      if (smem2 > 48 * 1024) cudaFuncSetAttribute(tpbp1d_S{sid}, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);")
            E(f"      tpbp1d_S{sid}<<<(unsigned)a.n_nodes, 32 * {Gs[sid]} * 2, smem2, s>>>(a2); }}")
            E(f"    else if (a.mul == 64) {{ if (smem > 48 * 1024) cudaFuncSetAttribute(tpbp_S{sid}<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);")
            E(f"      tpbp_S{sid}<64><<<(unsigned)a.n_nodes, 32 * {Gs[sid]} * 2, smem, s>>>(a); }}")
            E(f"    else {{ if (smem > 48 * 1024) cudaFuncSetAttribute(tpbp_S{sid}<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);")
            E(f"      tpbp_S{sid}<32><<<(unsigned)a.n_nodes, 32 * {Gs[sid]}, smem, s>>>(a); }}")
            E("    return;")
            E("  }")
        if kind == "f":
            E(f"  if ((a.mul == 64 || a.mul == 32) && e3b_tp_pipelined_enabled()) {{")
            E(f"    const size_t smem = tpfp_smem_S{sid}(a.mul);")
            E(f"    if (a.mul == 64 && e3b_tp_paired_fwd_enabled(kPairedFwdOk_S{sid})) {{")
            E(f"      TpArgs<float> a2 = a; a2.n_stages = e3b_tp_stages(0, tpp2_smem_S{sid}(1) - tpp2_smem_S{sid}(0), tpp2_smem_S{sid}(0));")
            E(f"      const size_t smem2 = tpp2_smem_S{sid}(a2.n_stages);")
            E(f"      if (smem2 > 48 * 1024) cudaFuncSetAttribute(tpfp2_S{sid}, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);")
            E(f"      tpfp2_S{sid}<<<(unsigned)a.n_nodes, 32 * kPairedGroups_S{sid}, smem2, s>>>(a2); }}")
            E(f"    else if (a.mul == 64) {{ if (smem > 48 * 1024) cudaFuncSetAttribute(tpfp_S{sid}<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);")
            E(f"      tpfp_S{sid}<64><<<(unsigned)a.n_nodes, 32 * {Gs[sid]} * 2, smem, s>>>(a); }}")
            E(f"    else {{ if (smem > 48 * 1024) cudaFuncSetAttribute(tpfp_S{sid}<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);")
            E(f"      tpfp_S{sid}<32><<<(unsigned)a.n_nodes, 32 * {Gs[sid]}, smem, s>>>(a); }}")
            E("    return;")
            E("  }")
        E(f"  if (a.mul == 64) tp{kind}_S{sid}<float, 64><<<(unsigned)grid, TP_THREADS, 0, s>>>(a);")
        E(f"  else if (a.mul == 32) tp{kind}_S{sid}<float, 32><<<(unsigned)grid, TP_THREADS, 0, s>>>(a);")
        E(f"  else tp{kind}_S{sid}<float, 0><<<(unsigned)grid, TP_THREADS, 0, s>>>(a);")
        E("}")


def emit_tables(st_list):
    E = Emitter()
    E("// GENERATED by csrc/gen_tp.py -- do not edit.")
    E("#pragma once")
    E()
    Gs, texts = [], []
    for sid, st in enumerate(st_list):
        sub = Emitter()
        sub(f"// ---- structure S{sid}: in l={[b.ir.l for b in st.irreps_in]} sh l={[b.ir.l for b in st.irreps_sh]} "
            f"paths={len(st.paths)} out comps={st.y_layout()[2]}")
        Gs.append(emit_structure(sub, sid, st))
        sub("#ifdef __CUDACC__")
        emit_launchers(sub, sid, Gs[sid])
        sub("#endif  // __CUDACC__")
        texts.append(sub.text())
    # the structures are spread over N_PARTS translation units (tp_fast.cu = part 0, tp_fast_p{k}.cu), balanced by size; a
    # CUDA compile with E3B_TP_PART defined sees only its part, anything else (host emulation, a single-TU build) sees all
    part_of, load = {}, [0] * N_PARTS
    for sid in sorted(range(len(st_list)), key=lambda i: -len(texts[i])):
        k = min(range(N_PARTS), key=lambda q: load[q])
        part_of[sid] = k
        load[k] += len(texts[sid])
    for sid in range(len(st_list)):
        E(f"#if !defined(__CUDACC__) || !defined(E3B_TP_PART) || E3B_TP_PART == {part_of[sid]}")
        E.lines.append(texts[sid].rstrip("\n"))
        E(f"#endif  // part {part_of[sid]}")
    E("#ifdef E3B_HOST_EMU")
    E("// CPU emulation entry (tests only): runs every (node, channel, group) item serially.")
    E("template <typename T> static int emu_tp(int sid, int bwd, const TpArgs<T>& a) {")
    E("  for (int64_t node = 0; node < a.n_nodes; ++node)")
    E("    for (int u = 0; u < a.mul; ++u) {")
    E("      const int chunk = u / 32, lane = u % 32; (void)lane;")
    E("      switch (sid) {")
    for sid in range(len(st_list)):
        E(f"        case {sid}:")
        for g in range(Gs[sid]):
            E(f"          if (bwd) tpb_S{sid}_g{g}<T, 0>(a, node, u, true, chunk * {Gs[sid]} + {g}, lane); else tpf_S{sid}_g{g}<T, 0>(a, node, u, true);")
        E("          break;")
    E("        default: return -1;")
    E("      }")
    E("    }")
    E("  return 0;")
    E("}")
    E("static const int kEmuGroups[] = {" + ", ".join(str(g) for g in Gs) + "};")
    E("#endif  // E3B_HOST_EMU")
    E("#if defined(__CUDACC__) && (!defined(E3B_TP_PART) || E3B_TP_PART == 0)")
    for sid in range(len(st_list)):
        E(f"void launch_tpf_S{sid}(const TpArgs<float>&, int64_t, cudaStream_t); void launch_tpb_S{sid}(const TpArgs<float>&, int64_t, cudaStream_t);")

    def arr(v):
        return "{" + ", ".join(str(int(t)) for t in v) + "}"

    E(f"static const int kNumGenEntries = {len(st_list)};")
    E("static const GenEntry kGenEntries[] = {")
    for sid, st in enumerate(st_list):
        E("  {" + ", ".join([
            str(len(st.irreps_in)), arr(b.ir.l for b in st.irreps_in),
            str(len(st.irreps_sh)), arr(b.ir.l for b in st.irreps_sh),
            str(len(st.paths)), arr(p.i_in for p in st.paths), arr(p.i_sh for p in st.paths),
            arr(p.ir_out.l for p in st.paths), arr(p.slot for p in st.paths),
            arr(st.y_layout()[0][p.slot] for p in st.paths), arr(st.y_layout()[1][p.slot] for p in st.paths),
            str(Gs[sid]), f"launch_tpf_S{sid}", f"launch_tpb_S{sid}", str(PAIRED_INFO[sid][0]), "true" if PAIRED_INFO[sid][1] else "false"]) + "},")
    E("};")
    E("#endif  // __CUDACC__")
    return E.text()


def emit_cg_tables(lmax=3):
    E = Emitter()
    E("// GENERATED by csrc/gen_tp.py -- dense real Wigner-3j tables (analytic signs), l <= %d." % lmax)
    E("#pragma once")
    n = lmax + 1
    offs, data, signs = [], [], []
    for l1 in range(n):
        for l2 in range(n):
            for l3 in range(n):
                if abs(l1 - l2) <= l3 <= l1 + l2:
                    offs.append(len(data))
                    data.extend(cg.w3j(l1, l2, l3).reshape(-1).tolist())
                    signs.append(int(cg._preset_sign(l1, l2, l3)))
                else:
                    offs.append(-1)
                    signs.append(1)
    E(f"#define E3B_CG_LMAX {lmax}")
    E(f"static __device__ const int kCgOffset[{len(offs)}] = {{" + ", ".join(map(str, offs)) + "};")
    E(f"static const int kCgSign044[{len(signs)}] = {{" + ", ".join(map(str, signs)) + "};")
    E(f"static __device__ const double kCgData[{len(data)}] = {{")
    for i in range(0, len(data), 6):
        E("  " + ", ".join(repr(v) for v in data[i:i + 6]) + ",")
    E("};")
    return E.text()


def main():
    sts = generated_structures()
    if os.environ.get("E3B_GEN_ONLY"):      # compile-time experiments on a few structures (the output is not a working table)
        sts = [sts[int(i)] for i in os.environ["E3B_GEN_ONLY"].split(",")]
        with open(os.environ.get("E3B_GEN_OUT", "/tmp/tp_generated_only.cuh"), "w") as f:
            f.write(emit_tables(sts))
        return
    with open(os.path.join(HERE, "tp_generated.cuh"), "w") as f:
        f.write(emit_tables(sts))
    with open(os.path.join(HERE, "cg_tables.cuh"), "w") as f:
        f.write(emit_cg_tables())
    print(f"generated {len(sts)} structures:", [(len(s.paths), [b.ir.l for b in s.irreps_in]) for s in sts])


if __name__ == "__main__":
    main()

/*
 * e3b200.h -- C ABI of libe3b200.so: the B200 (sm_100a) hot path of Equivariant-NN-Zoo.
 *
 * The reference (pure Python on e3nn 0.4.4) has no FFI; the seam this library plugs into is
 * its Python module protocol.  Each entry point below names the
 * reference interface it replaces (paths relative to /root/reference).  The host side that
 * binds these symbols with ctypes is equivariant-nn-zoo_b200/e3b200/_lib.py; the reference-side
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - return 0 on success, a negative e3b_status otherwise; text via e3b_last_error()
 *     (thread-local, valid until the next call on the same thread);
 *   - no C++ exceptions cross the boundary;
 *   - the CALLER allocates every buffer (torch caching allocator); the library never
 *     allocates device memory and never synchronises;
 *   - all pointers are DEVICE pointers unless named h_*; functions are asynchronous on the
 *     supplied stream (a cudaStream_t passed as void*);
 *   - dtype: 0 = float32, 1 = float64 (the fp64 build of the same templates is the
 *     1e-10 correctness mode);
 *   - node feature rows cross this ABI in the library's "channel-fastest" layout
 *     [block][m][u] ("imu") unless stated otherwise; e3nn's "mul_ir" layout is [block][u][m].
 *     e3b_layout_convert translates.
 */
#ifndef E3B200_H
#define E3B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E3B_ABI_VERSION 1
#define E3B_API __attribute__((visibility("default")))

typedef enum {
  E3B_OK = 0,
  E3B_ERR_INVALID = -1,     /* bad argument / descriptor            */
  E3B_ERR_UNSUPPORTED = -2, /* irreps outside what the kernels cover */
  E3B_ERR_CUDA = -3,        /* launch failed; see e3b_last_error()   */
  E3B_ERR_NOMEM = -4        /* host allocation for a plan failed     */
} e3b_status;

#define E3B_F32 0
#define E3B_F64 1

E3B_API int e3b_abi_version(void);
E3B_API const char* e3b_last_error(void);
/* sizeof of the ABI structs as compiled (0 e3b_tp_desc, 1 e3b_gate_desc, 2 e3b_gemm_problem,
 * 3 e3b_gemm_pack_desc, 4 e3b_wgrad_problem) so that a binding can verify its own struct layout; -1 otherwise */
E3B_API int64_t e3b_struct_size(int which);

/* ---------------------------------------------------------------------------------------
 * Neighbour list.   Replaces computeEdgeIndex, e3_layers/data/compute_edge.py:38-113
 * (radius part + self-loop removal; the optional `criteria` edges are OR-ed in by the host
 * as an extra edge list).  Semantics: per graph g (nodes node_ptr[g]..node_ptr[g+1]) every
 * ordered pair (a,b), a != b, with  sqrtf(fmaf(dz,dz,fmaf(dy,dy,dx*dx))) < r_max  in fp32
 * (bit-identical to torch.linalg.norm on the fp32 difference, compute_edge.py:68-70).
 * Output order = the reference's: lexicographic in (edge_index[0]=a, edge_index[1]=b).
 * Two phases so that the caller allocates the output:
 *   count: deg[a] = number of neighbours of a                       (int32 [N])
 *   (caller: row_ptr = exclusive scan of deg, int64 [N+1]; E = row_ptr[N])
 *   fill : edge_index int64 [2,E]; optional rev int32 [E] with rev[e] = position of the
 *          reversed edge (b,a) -- the radius graph is symmetric, so (row_ptr, edge_index[1],
 *          rev) is also the dst-grouped CSR the convolution consumes.
 * pos_stride = floats between consecutive rows of pos (3 for a dense [N,3]).
 */
E3B_API int e3b_radius_graph_count(const float* pos, int64_t pos_stride, const int64_t* node_ptr /* [G+1] */,
                           int32_t n_graphs, int64_t n_nodes, float r_max, int32_t* deg, void* stream);
E3B_API int e3b_radius_graph_fill(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                          int64_t n_nodes, float r_max, const int64_t* row_ptr, int64_t n_edges,
                          int64_t* edge_index /* [2,E] */, int32_t* rev /* [E] or NULL */, void* stream);

/* Bucketed edge count, for replaying one captured CUDA graph over batches of different sizes (the reference has no
 * counterpart: its eager PyTorch path re-allocates per batch, data/compute_edge.py:56-75).  The caller appends n_pad
 * padding atoms to the batch as the LAST graph, laid out as isolated pairs (2p, 2p+1) closer than r_max (an odd last
 * atom stays alone), runs e3b_radius_graph_count + scan over the padded batch (n_edges_natural = row_ptr[N]) and picks
 * n_edges_total >= n_edges_natural (same parity).  This call rewrites the padding rows of row_ptr so that the edge
 * list has exactly n_edges_total entries -- the missing ones are parallel copies of the pair edges, spread evenly
 * over the pairs -- and fills edge_index [2, n_edges_total], rev and (optionally) nbr32 = int32(edge_index[1]).
 * Real atoms get exactly the edges e3b_radius_graph_fill gives them, in the same order.                       */
E3B_API int e3b_radius_graph_fill_padded(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                         int64_t n_nodes /* incl. padding */, int64_t n_pad, float r_max,
                                         int64_t* row_ptr /* [N+1], in/out */, int64_t n_edges_natural,
                                         int64_t n_edges_total, int64_t* edge_index, int32_t* rev,
                                         int32_t* nbr32 /* [n_edges_total] or NULL */, void* stream);

/* Neighbour list with the `criteria` of the protein config evaluated inside the sweep (replaces the
 * criteria(data, all_pairs) call of compute_edge.py:73-74 for e3_layers/configs/config_diffusion_CA.py:58-64):
 * pair (a,b), a != b, same graph, is an edge iff   |pos[a]-pos[b]| < r_max
 *      OR (segment != NULL and segment[a] == segment[b] and |a - b| < max_separation)
 *      OR (p_random > 0 and u(a,b) < p_random),
 * u(a,b) = uniforms[pair_ptr[g] + (a - first_g) * n_g + (b - first_g)] when `uniforms` is given (the reference's
 * torch.rand over its all-pairs list, indexable so that a seeded run can be reproduced edge for edge), else a
 * counter-based hash of (seed, a, b).  Same two phases and output order as e3b_radius_graph_*; the graph is not
 * symmetric in general, so there is no `rev` (use e3b_csr_fill).  The struct holds DEVICE pointers.            */
typedef struct {
  const int64_t* segment;     /* [N] or NULL */
  int64_t max_separation;
  float p_random;
  const float* uniforms;      /* [sum_g n_g^2] or NULL */
  const int64_t* pair_ptr;    /* [G+1], exclusive scan of n_g^2; required with uniforms */
  uint64_t seed;
} e3b_pair_criteria;
E3B_API int e3b_pair_graph_count(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                 int64_t n_nodes, float r_max, const e3b_pair_criteria* h_crit, int32_t* deg, void* stream);
E3B_API int e3b_pair_graph_fill(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                int64_t n_nodes, float r_max, const e3b_pair_criteria* h_crit, const int64_t* row_ptr,
                                int64_t n_edges, int64_t* edge_index, void* stream);

/* Cell-list variant of e3b_radius_graph_* for large graphs (same predicate, same output, bit for bit).
 *   bin  : graphs with >= min_nodes atoms get a uniform grid (cells >= 1.001 r_max, at most
 *          cell_ptr[g+1]-cell_ptr[g] cells; the caller reserves about two cells per atom); writes grids
 *          [G * E3B_CELL_GRID_BYTES], cell_of int32 [N] (-1 = graph not binned) and adds to the zeroed cell_count [C];
 *   (caller: cell_start = exclusive scan of cell_count, int64 [C+1])
 *   sort : atoms grouped by cell into sorted int32 [N] (cursor int32 [C], zeroed);
 *   count / fill: as e3b_radius_graph_count / _fill; small graphs of the batch take the all-pairs loop.        */
#define E3B_CELL_GRID_BYTES 48
E3B_API int e3b_cell_graph_bin(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                               int64_t n_nodes, float r_max, int64_t min_nodes, const int64_t* cell_ptr /* [G+1] */,
                               void* grids, int32_t* cell_of, int32_t* cell_count, void* stream);
E3B_API int e3b_cell_graph_sort(const int32_t* cell_of, int64_t n_nodes, const int64_t* cell_start, int32_t* cursor,
                                int32_t* sorted, void* stream);
E3B_API int e3b_cell_graph_count(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                 int64_t n_nodes, float r_max, const void* grids, const int64_t* cell_start,
                                 const int32_t* sorted, int32_t* deg, void* stream);
E3B_API int e3b_cell_graph_fill(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                int64_t n_nodes, float r_max, const void* grids, const int64_t* cell_start,
                                const int32_t* sorted, const int64_t* row_ptr, int64_t n_edges, int64_t* edge_index,
                                int32_t* rev /* [E] or NULL */, void* stream);

/* Dst-grouped CSR of an arbitrary edge list (any order, e.g. with `criteria` edges or a
 * user-supplied edge_index).  The caller passes row_ptr = exclusive scan of the in-degree
 * (int64 [N+1]) and a zeroed int32 cursor[N]; the kernel fills, per destination node, the
 * edge ids of its incoming edges in ASCENDING edge id (deterministic).              */
E3B_API int e3b_csr_fill(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, int which_row /*0|1*/,
                 const int64_t* row_ptr, int32_t* cursor /* [N], zeroed */, int32_t* eid /* [E] */, void* stream);

/* ---------------------------------------------------------------------------------------
 * Edge geometry.  Replaces computeEdgeVector (data/compute_edge.py:13-36),
 * SphericalEncoding (nn/embedding.py:130-178; e3nn o3.SphericalHarmonics, normalize=True,
 * 'component', l <= 2 -- every in-scope config uses 1x0e+1x1o+1x2e) and RadialBasisEncoding (nn/embedding.py:181-219: Bessel x polynomial
 * cutoff p, or the symmetric cutoff of :26-29 when cutoff_kind = 1).
 */
E3B_API int e3b_edge_vectors_fwd(int dtype, const void* pos, const int64_t* edge_index, int64_t n_edges,
                         void* vec /* [E,3] */, void* len /* [E] or NULL */, void* stream);
/* gpos[n] += sum_{e: dst=n} g[e] - sum_{e: src=n} g[e]  via the two CSR views, no atomics.
 * g = gvec + glen * vec/len (either may be NULL). */
E3B_API int e3b_edge_vectors_bwd(int dtype, const void* gvec, const void* glen, const void* vec, const void* len,
                         int64_t n_nodes, const int64_t* in_ptr, const int32_t* in_eid,
                         const int64_t* out_ptr, const int32_t* out_eid, void* gpos /* [N,3] */, void* stream);

E3B_API int e3b_sh_fwd(int dtype, const void* vec, int64_t n, int lmax, int normalize, void* sh /* [n,(lmax+1)^2] */,
               void* stream);
E3B_API int e3b_sh_bwd(int dtype, const void* vec, const void* gsh, int64_t n, int lmax, int normalize,
               void* gvec /* [n,3] */, void* stream);

E3B_API int e3b_radial_fwd(int dtype, const void* r, int64_t n, const void* bessel_w, int n_basis, double r_max,
                   double r_min, int one_over_r, int cutoff_kind, double p, void* out /* [n,n_basis] */,
                   void* stream);
/* gr [n]; gw_partial [n_blocks, n_basis] partial sums of d/d bessel_w (caller sums rows;
 * n_blocks = e3b_radial_bwd_blocks(n)). */
E3B_API int64_t e3b_radial_bwd_blocks(int64_t n);
E3B_API int e3b_radial_bwd(int dtype, const void* r, const void* gout, int64_t n, const void* bessel_w, int n_basis,
                   double r_max, double r_min, int one_over_r, int cutoff_kind, double p, void* gr,
                   void* gw_partial, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused tensor-product convolution.  Replaces, in FactorizedConvolution.forward
 * (nn/message_passing.py:104-109): x[edge_src] gather, TensorProductExpansion.tp
 * (o3.TensorProduct 'uvu', external per-edge weights, nn/pointwise.py:61-85,98) and
 * scatter(edge_features, edge_dst) -- with the post-TP o3.Linear (pointwise.py:99) moved
 * AFTER the reduction by linearity.
 *
 *   y[n, p, k, u] = sum_{e in in(n)} w[e, p, u] * sqrt(2 l3+1) *
 *                   sum_{ij} C^{l1 l2 l3}_{ijk} x[src(e), b1(p), i, u] * Y[e, b2(p), j]
 *
 * A plan fixes the irreps: input blocks (l, parity) all with multiplicity `mul`, SH blocks,
 * and the path list in e3nn instruction order (weight layout) with each path's slot in the
 * sorted output (irreps_mid).  Layouts: x [N][sum_b (2l_b+1)][mul]; Y [E][sum (2l+1)];
 * w [E][n_paths][mul] (== e3nn's weight layout for uvu with mul_in2 = 1);
 * y [N][...][mul]: output slots sorted by irrep; the slots of one irrep (l3, parity) form a
 * group stored [k][slot-in-group][u], i.e. the imu layout of the simplified irreps_mid block.
 */
#define E3B_MAX_BLOCKS 16
#define E3B_MAX_PATHS 96

typedef struct {
  int32_t mul;                       /* common multiplicity of every input block        */
  int32_t n_in;
  int32_t in_l[E3B_MAX_BLOCKS];      /* l of input block b                              */
  int32_t in_p[E3B_MAX_BLOCKS];      /* parity (+1/-1); only used to group output slots  */
  int32_t n_sh;
  int32_t sh_l[E3B_MAX_BLOCKS];
  int32_t sh_p[E3B_MAX_BLOCKS];      /* parity of SH block (output parity = in_p * sh_p)  */
  int32_t n_paths;
  int32_t path_in[E3B_MAX_PATHS];    /* input block index                               */
  int32_t path_sh[E3B_MAX_PATHS];    /* SH block index                                  */
  int32_t path_lout[E3B_MAX_PATHS];  /* l3                                              */
  int32_t path_slot[E3B_MAX_PATHS];  /* position of this path's output block in y       */
  int32_t w3j_sign_preset;           /* 0 = analytic (e3nn >= 0.5), 1 = e3nn044 (R1)    */
} e3b_tp_desc;

typedef struct e3b_tp_plan e3b_tp_plan;

E3B_API int e3b_tp_plan_create(const e3b_tp_desc* desc, e3b_tp_plan** out);
E3B_API void e3b_tp_plan_destroy(e3b_tp_plan* plan);
/* 1 if a generated (fully unrolled) kernel matches this plan, 0 if the generic one runs */
E3B_API int e3b_tp_plan_is_specialized(const e3b_tp_plan* plan);
/* row widths in scalars, and n_part = number of partial rows per edge the f32 backward writes
 * into gsh (1 for the generic kernel, which accumulates with atomics into a ZEROED buffer) */
E3B_API int e3b_tp_plan_dims(const e3b_tp_plan* plan, int32_t* x_dim, int32_t* sh_dim, int32_t* w_dim, int32_t* y_dim,
                     int32_t* n_part_f32);

/* in_ptr/in_nbr/in_eid: dst-grouped CSR (in_nbr[k] = source node, in_eid[k] = row of w / Y;
 * in_eid may be NULL when the edge arrays are already in CSR order).                        */
E3B_API int e3b_tpconv_fwd(const e3b_tp_plan* plan, int dtype, int64_t n_nodes, int64_t n_edges, const void* x,
                   const void* sh, const void* w, const int64_t* in_ptr, const int32_t* in_nbr,
                   const int32_t* in_eid, void* y, void* stream);
/* First-order backward.  gx_edge [E, x_dim] receives the per-edge contribution to d/dx[src]
 * (row = edge id); the caller reduces it by source with e3b_segment_sum over the src-grouped
 * CSR (deterministic, no atomics).  gw [E, w_dim] is written directly.  gsh is
 * [E, n_part, sh_dim]: each warp role writes its own partial row (caller sums over n_part).
 * When the plan is not specialized (or dtype is f64) the generic kernel runs: gx_edge and gsh
 * must then be ZERO-filled by the caller (it accumulates with atomics, n_part = 1).
 * gx_edge and gsh may be NULL when those gradients are not needed.                         */
E3B_API int e3b_tpconv_bwd(const e3b_tp_plan* plan, int dtype, int64_t n_nodes, int64_t n_edges, const void* x,
                   const void* sh, const void* w, const void* gy, const int64_t* in_ptr, const int32_t* in_nbr,
                   const int32_t* in_eid, void* gx_edge, void* gsh, void* gw, void* stream);

/* out[n, :] = sum_{k in [ptr[n], ptr[n+1])} src[ids ? ids[k] : k, :]  (rows of `width` scalars).
 * Replaces torch_runstats scatter at nn/message_passing.py:109 / nn/output.py:69 (Pooling). */
/* Backward with d/dx reduced per SOURCE node inside the kernel (fp32 only; generated structures with multiplicity 32 / 64):
 * every edge's gradient row is staged in shared memory and added into gx_node[src] by one TMA reduce-add
 * (cp.reduce.async.bulk), so neither the [E, x_dim] per-edge buffer nor the segment sum of e3b_tpconv_bwd's caller is needed.
 * gx_node [N, x_dim] must be ZEROED by the caller; the order of the additions into a row is not fixed (results repeat to
 * fp32 rounding, ~1e-7 relative, not bit for bit) -- callers that need bit-reproducible gradients use e3b_tpconv_bwd.   */
E3B_API int e3b_tpconv_bwd_nodes(const e3b_tp_plan* plan, int64_t n_nodes, int64_t n_edges, const float* x, const float* sh,
                                 const float* w, const int32_t* w_idx /* [E] or NULL, see e3b_tpconv_fwd_shared */,
                                 const float* gy, const int64_t* in_ptr, const int32_t* in_nbr, const int32_t* in_eid,
                                 float* gx_node, float* gsh /* [E, n_part, sh_dim] or NULL */, float* gw, void* stream);
/* Forward with SHARED weight rows: edge e reads w[w_idx[e]] instead of w[e].  The radial weights are a function of the edge
 * LENGTH (nn/message_passing.py:93 on nn/embedding.py:181-219), so the two directions of an undirected edge of the radius
 * graph carry bit-identical rows: the radial MLP is evaluated once per undirected edge ([E/2, W] rows, half the GEMM work and
 * half the HBM write) and both directions read the same row (the second read usually hits L2).  Same restrictions as
 * e3b_tpconv_bwd_nodes (fp32, generated structures with multiplicity 32 / 64).                                          */
E3B_API int e3b_tpconv_fwd_shared(const e3b_tp_plan* plan, int64_t n_nodes, int64_t n_edges, const float* x, const float* sh,
                                  const float* w, const int32_t* w_idx, const int64_t* in_ptr, const int32_t* in_nbr,
                                  const int32_t* in_eid, float* y, void* stream);
/* out[u, k] = (g[canon[u], k] + g[rev[canon[u]], k]) * d/dz[cst * ssp](z), derivative from the stored activation
 * h[u, k] = cst * ssp(z) (h == NULL: factor 1): folds the gradients of the two directions of every undirected edge and
 * applies the activation derivative of the last hidden radial layer in one pass.  width % 4 == 0.                       */
E3B_API int e3b_pair_sum_act(const float* g /* [E, width] */, const int64_t* canon /* [n_unique] */, const int32_t* rev /* [E] */,
                             const float* h /* [n_unique, width] or NULL */, float cst, int64_t n_unique, int32_t width,
                             float* out /* [n_unique, width] */, void* stream);
E3B_API int e3b_segment_sum(int dtype, const void* src, int64_t width, const int64_t* ptr, const int32_t* ids,
                    int64_t n_out, void* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Gate non-linearity.  Replaces e3nn nn.Gate as built at nn/message_passing.py:191-207:
 * input [scalars | gates | gated] in mul_ir layout; act codes: 0 identity, 1 silu, 2 tanh,
 * 3 ssp, 4 tanhlu, 5 abs; each multiplied by its normalize2mom constant `cst`.
 */
typedef struct {
  int32_t n_scalar_blocks;
  int32_t scalar_mul[E3B_MAX_BLOCKS];
  int32_t scalar_act[E3B_MAX_BLOCKS];
  double scalar_cst[E3B_MAX_BLOCKS];
  int32_t n_gated_blocks;
  int32_t gated_mul[E3B_MAX_BLOCKS];
  int32_t gated_l[E3B_MAX_BLOCKS];
  int32_t gate_act[E3B_MAX_BLOCKS];
  double gate_cst[E3B_MAX_BLOCKS];
} e3b_gate_desc;

E3B_API int e3b_gate_fwd(const e3b_gate_desc* desc, int dtype, const void* in, int64_t n, void* out, void* stream);
E3B_API int e3b_gate_bwd(const e3b_gate_desc* desc, int dtype, const void* in, const void* gout, int64_t n, void* gin,
                 void* stream);

/* Adjoint of e3b_gate_bwd (second order, needed when the reference differentiates its position
 * gradient again: nn/output.py:39-43, create_graph=self.training).  With gin = gate_bwd(in, gout)
 * and a cotangent ggin [n, in_dim] of gin:  g_in = d<ggin, gin>/d in  [n, in_dim],
 * g_gout = d<ggin, gin>/d gout  [n, out_dim].                                                   */
E3B_API int e3b_gate_bwd2(const e3b_gate_desc* desc, int dtype, const void* in, const void* gout, const void* ggin,
                          int64_t n, void* g_in, void* g_gout, void* stream);

/* The same gate on the library's channel-fastest layout: `in` is [scalars | gates | gated] with the
 * gated blocks stored [m][u]; the result is written in the imu layout (out_imu) and / or e3nn's
 * mul_ir layout (out_mul_ir); either may be NULL.  The backward accepts the gradient in either or
 * both layouts (summed) and returns d/d in in the imu layout.                                  */
E3B_API int e3b_gate_imu_fwd(const e3b_gate_desc* desc, int dtype, const void* in, int64_t n, void* out_mul_ir,
                             void* out_imu, void* stream);
E3B_API int e3b_gate_imu_bwd(const e3b_gate_desc* desc, int dtype, const void* in, const void* gout_mul_ir,
                             const void* gout_imu, int64_t n, void* gin, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense contractions on the tcgen05 tensor cores, fp32-faithful (3xTF32 split, fp32 TMEM
 * accumulators, accumulation chains cut every 64 floats of K).  Replaces the cuBLAS/einsum
 * calls behind e3nn o3.Linear (nn/message_passing.py:58-63,102; nn/pointwise.py:87-92,99),
 * nn.FullyConnectedNet (nn/message_passing.py:74-79,93) and o3.FullyConnectedTensorProduct
 * (nn/message_passing.py:83-87), forward and data-gradient backward.
 *
 *   acc[r, n] = sum_k A[r, k] B[n, k]             A: activation (HBM), B: weight (packed once)
 *   epilogue 0:  C[r, n]  = alpha * acc
 *   epilogue 1:  C[r, oc] = alpha * sum_{v < V} aux[(r / aux_d) * aux_ld + v] * acc[r, oc * V + v]
 *                (the self-connection: B rows ordered (w, v), v fastest; N = n_out * V)
 *   epilogue 2:  C[r, n]  = act_cst * ssp(alpha * acc)          (radial MLP hidden layer, forward)
 *   epilogue 3:  C[r, n]  = alpha * acc * d/dz[act_cst * ssp](z) (radial MLP backward), the
 *                derivative evaluated from the stored forward output H[r * h_ld + n]
 *   accumulate != 0: the value is ADDED to what C holds.
 * Row r of A starts at A + (r / a_d) * a_s1 + (r % a_d) * a_s2 (k contiguous); element (r, n) of
 * C is at C + (r / c_d) * c_s1 + (r % c_d) * c_s2 + n * c_s3.  K, a_s1, a_s2 and the base of A
 * must be multiples of 4 floats.
 *
 * B is consumed in a packed form (TF32 hi/lo split, UMMA canonical tiles) produced by
 * e3b_gemm_pack from any strided view of the weight: element (n, k), n = n1 * d + n2, is read
 * from src[n1 * s1 + n2 * s2 + k * sk] (zero when n2 >= n2_valid > 0).  dst needs e3b_gemm_packed_floats(N, K) floats,
 * 128-byte aligned; repack only when the weight changes.
 * e3b_gemm_run launches up to E3B_GEMM_MAX_GROUP independent problems (e.g. the irreps blocks
 * of one o3.Linear) as ONE persistent kernel; they must share the tile shape, i.e. agree on
 * (K <= 64) and on (N <= 64 or K > 64)  [e3b_gemm_tile_n returns the column-tile width].     */
#define E3B_GEMM_MAX_GROUP 16

typedef struct {
  const float* A;
  int64_t a_s1, a_s2;
  int32_t a_d;
  const float* B_packed;
  float* C;
  int64_t c_s1, c_s2, c_s3;
  int32_t c_d;
  const float* aux;      /* epilogue 1 */
  int64_t aux_ld;
  int32_t aux_d, V;
  int32_t aux_cols;      /* valid columns of aux (<= V; 0 means V): narrower attributes are zero-extended */
  const float* H;        /* epilogue 3 */
  int64_t h_ld;
  int32_t M, N, K;
  int32_t epilogue, accumulate;
  float alpha, act_cst;
  const int32_t* row_map;  /* grouped rows: virtual row group -> actual row group of A and C, < 0 = padding (or NULL) */
  const int32_t* b_sel;    /* weight set of every block of 128 virtual row groups (or NULL) */
  const int32_t* n_blocks; /* device scalar (or NULL): blocks of 128 virtual row groups actually in use; row tiles past
                              them are skipped (M, a static bound, keeps the launch shape independent of the batch) */
  int64_t b_set_stride;    /* floats between consecutive weight sets in B_packed */
} e3b_gemm_problem;

typedef struct {
  const float* src;
  int64_t s1, s2, sk;
  int32_t d;
  int32_t n2_valid;      /* rows with n2 >= n2_valid are zero (0 means all valid): pads a group to V */
  int32_t N, K;
  float* dst;
  int32_t n_sets;        /* > 1: that many weights of the same shape, set i read from src + i * set_stride and written */
  int64_t set_stride;    /*      to dst + i * e3b_gemm_packed_floats(N, K) (grouped rows of e3b_gemm_problem)        */
} e3b_gemm_pack_desc;

E3B_API int e3b_gemm_tile_n(int32_t N, int32_t K);
E3B_API int64_t e3b_gemm_packed_floats(int32_t N, int32_t K);
E3B_API int e3b_gemm_pack(const e3b_gemm_pack_desc* descs, int32_t n, void* stream);
E3B_API int e3b_gemm_run(const e3b_gemm_problem* problems, int32_t n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Weight gradients on the tcgen05 tensor cores (3xTF32, fp32 accumulate, deterministic split-K): the
 * reductions over all rows that autograd produces for e3nn o3.Linear / nn.FullyConnectedNet /
 * o3.FullyConnectedTensorProduct weights (reference: the torch.einsum / matmul backward behind
 * nn/message_passing.py:58-63,74-87 and nn/pointwise.py:87-92 when nn/output.py:39-43 or a loss is differentiated).
 *
 *   C[m, n] (+)= alpha * sum_{r < R} A'[r, m] * B[r, n]
 *   V == 0:  A'[r, m] = A[r, m],                                   m < K1
 *   V  > 0:  A'[r, v * K1 + u] = A[r, u] * aux[(r / aux_d) * aux_ld + v],  m = v * K1 + u < V * K1
 *            (self-connection weights W[u, v, w]: the feature x attribute outer product is formed in the loader)
 * Row r of A starts at A + (r / a_d) * a_s1 + (r % a_d) * a_s2 (K1 contiguous floats), of B likewise (K2
 * contiguous floats); C[m, n] is at C + (m / c_d) * c_s1 + (m % c_d) * c_s2 + n * c_s3.  K1, K2, the row strides
 * and the bases of A and B must be multiples of 4 floats; R < 2^31.
 * Up to E3B_WGRAD_MAX_GROUP problems per launch; the caller provides a workspace of
 * e3b_wgrad_workspace_floats() floats (partial tiles, summed in a fixed order by a second kernel).             */
#define E3B_WGRAD_MAX_GROUP 8
typedef struct {
  const float* A;
  int64_t a_s1, a_s2;
  int32_t a_d;
  const float* B;
  int64_t b_s1, b_s2;
  int32_t b_d;
  const float* aux;
  int64_t aux_ld;
  int32_t aux_d, V;
  float* C;
  int64_t c_s1, c_s2, c_s3;
  int32_t c_d;
  int64_t R;
  int32_t K1, K2;
  float alpha;
  int32_t accumulate;
} e3b_wgrad_problem;
E3B_API int64_t e3b_wgrad_workspace_floats(const e3b_wgrad_problem* problems, int32_t n);   /* -1: invalid / unsupported */
E3B_API int e3b_wgrad_run(const e3b_wgrad_problem* problems, int32_t n, float* workspace, void* stream);

/* ---------------------------------------------------------------------------------------
 * LayerNormalization.  Replaces nn/pointwise.py:32-51 (used when normalize=True,
 * nn/message_passing.py:238-240,255-257; config_diffusion_CA.py:126): per node and irreps
 * block b (all mul_b (2 l_b + 1) entries, mul_ir layout)
 *   y = x * rinv * std[b],  rinv = (sum x^2 / mul_b + eps)^-1/2     (rinv [n, n_blocks] is kept for the backward)
 * Backward: g_x, and per-CTA partial sums of d/d std ([e3b_layernorm_bwd_blocks(n), n_blocks],
 * summed by the caller; may be NULL).                                                            */
E3B_API int e3b_layernorm_fwd(int dtype, const void* x, int64_t n, int32_t n_blocks, const int32_t* h_mul,
                              const int32_t* h_l, const void* std_w, double eps, void* y, void* rinv, void* stream);
E3B_API int64_t e3b_layernorm_bwd_blocks(int64_t n);
E3B_API int e3b_layernorm_bwd(int dtype, const void* x, const void* gy, const void* rinv, int64_t n, int32_t n_blocks,
                              const int32_t* h_mul, const int32_t* h_l, const void* std_w, void* g_x,
                              void* g_std_partial, void* stream);

/* ---------------------------------------------------------------------------------------
 * Optimiser step on a flat fp32 parameter buffer: Adam (torch.optim.Adam semantics: the reference's optimiser,
 * run/trainer.py:370-386, train.py:103-107) fused with the exponential moving average of the parameters
 * (torch_ema as used in run/trainer.py and run/sde_utils.py:232-248; `ema` may be NULL).  grad_scale: optional
 * DEVICE float multiplying the gradient (clipping coefficient); skip: optional DEVICE int, non-zero = the Adam
 * update is skipped (non-finite gradients, run/sde_utils.py:240-246) while the moving average still advances.
 * step_in / step_out: optional pair of DISTINCT DEVICE int64 scalars holding the number of updates applied so
 * far (read from step_in, written to step_out; a skipped update does not count, as in torch.optim.Adam); when
 * NULL, `step` >= 1 is the number of this update (bias corrections).                                          */
E3B_API int e3b_adam_ema_step(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, void* ema, int64_t n,
                              float lr, double beta1, double beta2, float eps, float weight_decay, int64_t step,
                              float ema_decay, const void* grad_scale, const void* skip, const void* step_in,
                              void* step_out, void* stream);

/* mul_ir <-> imu layout conversion of feature rows (blocks: mul, l). to_imu = 1: [u][m]->[m][u] */
E3B_API int e3b_layout_convert(int dtype, const void* in, int64_t n, int32_t n_blocks, const int32_t* mul,
                       const int32_t* l, int to_imu, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* E3B200_H */

// libe3b200: C ABI + the non-generated kernels (neighbour list, edge geometry, generic tensor
// product, segmented sums, gate, layout conversion).  sm_100a only.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "cg_tables.cuh"
#include "common.cuh"
#include "tp_fast.h"

// ------------------------------------------------------------------------------------------
// errors
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// shared with the other translation units of the library (hidden visibility)
int e3b_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(E3B_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return E3B_OK;
}

extern "C" int e3b_abi_version(void) { return E3B_ABI_VERSION; }
extern "C" const char* e3b_last_error(void) { return g_err; }
extern "C" int64_t e3b_struct_size(int which) {
  switch (which) {
    case 0: return (int64_t)sizeof(e3b_tp_desc);
    case 1: return (int64_t)sizeof(e3b_gate_desc);
    case 2: return (int64_t)sizeof(e3b_gemm_problem);
    case 3: return (int64_t)sizeof(e3b_gemm_pack_desc);
    case 4: return (int64_t)sizeof(e3b_wgrad_problem);
    default: return -1;
  }
}

static inline unsigned blocks_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

#define DISPATCH_DTYPE(dtype, ...)                                  \
  if ((dtype) == E3B_F32) { typedef float T; __VA_ARGS__ }          \
  else if ((dtype) == E3B_F64) { typedef double T; __VA_ARGS__ }    \
  else return fail(E3B_ERR_INVALID, "dtype must be 0 (f32) or 1 (f64)");

// ------------------------------------------------------------------------------------------
// Neighbour list (data/compute_edge.py:38-113).  One warp per atom a; lanes sweep the atoms b
// of a's graph in ascending order, so edges come out sorted by (a, b) with ballot/popc
// compaction -- no atomics, deterministic, bit-exact predicate.
__device__ __forceinline__ bool within(const float* __restrict__ pos, int64_t stride, float ax, float ay, float az,
                                       int64_t b, float r_max) {
  const float dx = __fsub_rn(ax, pos[b * stride + 0]);
  const float dy = __fsub_rn(ay, pos[b * stride + 1]);
  const float dz = __fsub_rn(az, pos[b * stride + 2]);
  // torch.linalg.norm over 3 fp32 elements on CPU == sqrt(fma(z,z,fma(y,y,x*x)))
  const float d = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
  return d < r_max;
}

__device__ __forceinline__ int find_graph(const int64_t* __restrict__ node_ptr, int n_graphs, int64_t a) {
  int lo = 0, hi = n_graphs;  // node_ptr[lo] <= a < node_ptr[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (node_ptr[mid] <= a) lo = mid; else hi = mid;
  }
  return lo;
}

template <bool FILL>
__global__ void __launch_bounds__(256) radius_graph_kernel(const float* __restrict__ pos, int64_t stride,
                                                           const int64_t* __restrict__ node_ptr, int n_graphs,
                                                           int64_t n_nodes, float r_max, int32_t* __restrict__ deg,
                                                           const int64_t* __restrict__ row_ptr, int64_t n_edges,
                                                           int64_t* __restrict__ edge_index) {
  const int lane = threadIdx.x & 31;
  const int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= n_nodes) return;
  const int g = find_graph(node_ptr, n_graphs, a);
  const int64_t b0 = node_ptr[g], b1 = node_ptr[g + 1];
  const float ax = pos[a * stride], ay = pos[a * stride + 1], az = pos[a * stride + 2];
  int64_t cursor = FILL ? row_ptr[a] : 0;
  int count = 0;
  for (int64_t base = b0; base < b1; base += 32) {
    const int64_t b = base + lane;
    bool hit = false;
    if (b < b1 && b != a) hit = within(pos, stride, ax, ay, az, b, r_max);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (FILL) {
      if (hit) {
        const int64_t p = cursor + __popc(m & ((1u << lane) - 1u));
        edge_index[p] = a;
        edge_index[n_edges + p] = b;
      }
      cursor += __popc(m);
    } else {
      count += __popc(m);
    }
  }
  if (!FILL && lane == 0) deg[a] = count;
}

// rev[e] = index of the reversed edge (b, a): binary search for a among b's sorted neighbours.
// `count` edges are visited, `n_edges` is the row pitch of edge_index; nbr32 (optional) receives int32(b).
__global__ void reverse_edge_kernel(const int64_t* __restrict__ edge_index, const int64_t* __restrict__ row_ptr,
                                    int64_t count, int64_t n_edges, int32_t* __restrict__ rev, int32_t* __restrict__ nbr32) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  const int64_t a = edge_index[e], b = edge_index[n_edges + e];
  int64_t lo = row_ptr[b], hi = row_ptr[b + 1] - 1;
  int64_t found = -1;
  while (lo <= hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = edge_index[n_edges + mid];
    if (v == a) { found = mid; break; }
    if (v < a) lo = mid + 1; else hi = mid - 1;
  }
  rev[e] = (int32_t)found;
  if (nbr32) nbr32[e] = (int32_t)b;
}

// Bucketed edge count (CUDA-graph replay of batches with different sizes): the last n_pad atoms of the batch are
// padding atoms in a graph of their own, laid out as isolated pairs (2p, 2p + 1) (an odd last one has no partner).
// The natural radius graph gives every paired atom one neighbour; here the tail of row_ptr is rewritten so that the edge
// list has exactly n_total entries: the P = n_total - natural missing ones are parallel copies of the pair edges,
// spread evenly over the pairs (P / 2 per direction).  One thread per padding atom writes its whole row.
__global__ void pad_rows_kernel(int64_t* __restrict__ row_ptr, int64_t n_real, int64_t n_pad, int64_t n_total,
                                int64_t* __restrict__ edge_index, int32_t* __restrict__ rev, int32_t* __restrict__ nbr32) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pad) return;
  const int64_t pairs = n_pad >> 1;
  const int64_t e_real = row_ptr[n_real];                 // never rewritten (start of the first padding row)
  const int64_t half = (n_total - e_real - 2 * pairs) >> 1;   // extra edges per direction
  const int64_t base = pairs ? half / pairs : 0, rem = pairs ? half % pairs : 0;
  const int64_t p = j >> 1, s = j & 1;
  int64_t start, deg = 0, partner_start = 0;
  if (p < pairs) {
    const int64_t before = 2 * (p * (1 + base) + (p < rem ? p : rem));
    deg = 1 + base + (p < rem ? 1 : 0);
    start = e_real + before + s * deg;
    partner_start = e_real + before + (1 - s) * deg;
  } else {
    start = n_total;                                        // unpaired last atom: empty row
  }
  if (j > 0) row_ptr[n_real + j] = start;
  if (j == n_pad - 1) row_ptr[n_real + n_pad] = n_total;
  const int64_t a = n_real + j, b = n_real + (j ^ 1);
  for (int64_t i = 0; i < deg; ++i) {
    edge_index[start + i] = a;
    edge_index[n_total + start + i] = b;
    rev[start + i] = (int32_t)(partner_start + i);
    if (nbr32) nbr32[start + i] = (int32_t)b;
  }
}

extern "C" int e3b_radius_graph_count(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                      int64_t n_nodes, float r_max, int32_t* deg, void* stream) {
  if (n_nodes == 0) return E3B_OK;
  if (!pos || !node_ptr || !deg || n_graphs <= 0) return fail(E3B_ERR_INVALID, "radius_graph_count: null/empty argument");
  radius_graph_kernel<false><<<blocks_for(n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
      pos, pos_stride, node_ptr, n_graphs, n_nodes, r_max, deg, nullptr, 0, nullptr);
  return check_launch("radius_graph_count");
}

extern "C" int e3b_radius_graph_fill(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                     int64_t n_nodes, float r_max, const int64_t* row_ptr, int64_t n_edges,
                                     int64_t* edge_index, int32_t* rev, void* stream) {
  if (n_nodes == 0 || n_edges == 0) return E3B_OK;
  if (!pos || !node_ptr || !row_ptr || !edge_index) return fail(E3B_ERR_INVALID, "radius_graph_fill: null argument");
  radius_graph_kernel<true><<<blocks_for(n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
      pos, pos_stride, node_ptr, n_graphs, n_nodes, r_max, nullptr, row_ptr, n_edges, edge_index);
  int rc = check_launch("radius_graph_fill");
  if (rc) return rc;
  if (rev) {
    reverse_edge_kernel<<<blocks_for(n_edges, 256), 256, 0, (cudaStream_t)stream>>>(edge_index, row_ptr, n_edges, n_edges, rev,
                                                                                  nullptr);
    rc = check_launch("radius_graph_reverse");
  }
  return rc;
}

extern "C" int e3b_radius_graph_fill_padded(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                            int64_t n_nodes, int64_t n_pad, float r_max, int64_t* row_ptr,
                                            int64_t n_edges_natural, int64_t n_edges_total, int64_t* edge_index, int32_t* rev,
                                            int32_t* nbr32, void* stream) {
  if (!pos || !node_ptr || !row_ptr || !edge_index || !rev) return fail(E3B_ERR_INVALID, "radius_graph_fill_padded: null argument");
  const int64_t pairs = n_pad / 2, n_real = n_nodes - n_pad, e_real = n_edges_natural - 2 * pairs;
  if (n_pad < 0 || n_real < 0 || n_graphs < 2 || e_real < 0 || n_edges_total < n_edges_natural ||
      ((n_edges_total - n_edges_natural) & 1) || (n_edges_total > n_edges_natural && pairs == 0))
    return fail(E3B_ERR_INVALID, "radius_graph_fill_padded: need >= 1 padding pair and an even number of extra edges");
  cudaStream_t s = (cudaStream_t)stream;
  if (n_pad) {
    pad_rows_kernel<<<blocks_for(n_pad, 128), 128, 0, s>>>(row_ptr, n_real, n_pad, n_edges_total, edge_index, rev, nbr32);
    int rc = check_launch("radius_graph_pad_rows");
    if (rc) return rc;
  }
  if (n_real && e_real) {
    // the real atoms: graphs 0 .. n_graphs - 2; edge_index has row pitch n_edges_total
    radius_graph_kernel<true><<<blocks_for(n_real, 8), 256, 0, s>>>(pos, pos_stride, node_ptr, n_graphs - 1, n_real, r_max,
                                                                     nullptr, row_ptr, n_edges_total, edge_index);
    int rc = check_launch("radius_graph_fill");
    if (rc) return rc;
    reverse_edge_kernel<<<blocks_for(e_real, 256), 256, 0, s>>>(edge_index, row_ptr, e_real, n_edges_total, rev, nbr32);
    rc = check_launch("radius_graph_reverse");
    if (rc) return rc;
  }
  return E3B_OK;
}

// Undirected-edge bookkeeping of the radial MLP: out[u, k] = (g[canon[u], k] + g[rev[canon[u]], k]) * d/dz[cst ssp](z)
// with the derivative evaluated from the stored activation h[u, k] = cst * ssp(z) (h == null: factor 1).
__global__ void pair_sum_act_kernel(const float* __restrict__ g, const int64_t* __restrict__ canon, const int32_t* __restrict__ rev,
                                    const float* __restrict__ h, float cst, int64_t n_unique, int width, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one float4 per thread
  const int w4 = width >> 2;
  if (idx >= n_unique * w4) return;
  const int64_t u = idx / w4;
  const int c = (int)(idx - u * w4) * 4;
  const int64_t e = canon[u], r = rev[e];
  const float4 a = *reinterpret_cast<const float4*>(g + e * width + c), b = *reinterpret_cast<const float4*>(g + r * width + c);
  float4 o = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  if (h) {
    const float4 hv = *reinterpret_cast<const float4*>(h + u * width + c);
    const float ic = 1.f / cst;
    o.x *= cst * (1.f - 0.5f * __expf(-hv.x * ic)); o.y *= cst * (1.f - 0.5f * __expf(-hv.y * ic));
    o.z *= cst * (1.f - 0.5f * __expf(-hv.z * ic)); o.w *= cst * (1.f - 0.5f * __expf(-hv.w * ic));
  }
  *reinterpret_cast<float4*>(out + u * width + c) = o;
}

extern "C" int e3b_pair_sum_act(const float* g, const int64_t* canon, const int32_t* rev, const float* h, float cst,
                                int64_t n_unique, int32_t width, float* out, void* stream) {
  if (n_unique == 0) return E3B_OK;
  if (!g || !canon || !rev || !out || width <= 0 || (width & 3)) return fail(E3B_ERR_INVALID, "pair_sum_act: bad argument");
  pair_sum_act_kernel<<<blocks_for(n_unique * (width / 4), 256), 256, 0, (cudaStream_t)stream>>>(g, canon, rev, h, cst, n_unique,
                                                                                                 width, out);
  return check_launch("pair_sum_act");
}

// ------------------------------------------------------------------------------------------
// Neighbour list with pair predicates evaluated inside the sweep (the `criteria` callable of
// config_diffusion_CA.py:58-64: same chain and |a - b| < 5, OR a Bernoulli(p) draw per ordered pair) and, for
// large radius-only graphs, a cell list.  Same warp-per-atom sweep, same exact predicate, same output order.
__device__ __forceinline__ float pair_uniform(unsigned long long seed, int64_t a, int64_t b) {
  // counter-based: splitmix64 finaliser of (seed, a, b) -> 24 random bits -> [0, 1)
  unsigned long long x = seed ^ ((unsigned long long)a * 0x9E3779B97F4A7C15ull + (unsigned long long)b * 0xC2B2AE3D27D4EB4Full +
                                 0x165667B19E3779F9ull);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (float)(x >> 40) * (1.0f / 16777216.0f);
}

struct PairCrit {
  const int64_t* seg;        // per-node segment id (chain), or NULL
  long long max_sep;         // |a - b| < max_sep within a segment
  float p;                   // Bernoulli probability per ordered pair, 0 = off
  const float* uniforms;     // explicit uniforms indexed like the reference's all-pairs list, or NULL (hash RNG)
  const int64_t* pair_ptr;   // [G+1] exclusive scan of n_g^2 (with uniforms)
  unsigned long long seed;
};

__device__ __forceinline__ bool pair_extra(const PairCrit& c, int64_t a, int64_t b, int g, int64_t b0, int64_t n_g) {
  bool hit = false;
  if (c.seg) {
    const long long d = a > b ? a - b : b - a;
    hit = c.seg[a] == c.seg[b] && d < c.max_sep;
  }
  if (!hit && c.p > 0.f) {
    const float u = c.uniforms ? c.uniforms[c.pair_ptr[g] + (a - b0) * n_g + (b - b0)] : pair_uniform(c.seed, a, b);
    hit = u < c.p;
  }
  return hit;
}

template <bool FILL>
__global__ void __launch_bounds__(256) pair_graph_kernel(const float* __restrict__ pos, int64_t stride,
                                                         const int64_t* __restrict__ node_ptr, int n_graphs,
                                                         int64_t n_nodes, float r_max, PairCrit crit, int32_t* __restrict__ deg,
                                                         const int64_t* __restrict__ row_ptr, int64_t n_edges,
                                                         int64_t* __restrict__ edge_index) {
  const int lane = threadIdx.x & 31;
  const int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= n_nodes) return;
  const int g = find_graph(node_ptr, n_graphs, a);
  const int64_t b0 = node_ptr[g], b1 = node_ptr[g + 1];
  const float ax = pos[a * stride], ay = pos[a * stride + 1], az = pos[a * stride + 2];
  int64_t cursor = FILL ? row_ptr[a] : 0;
  int count = 0;
  for (int64_t base = b0; base < b1; base += 32) {
    const int64_t b = base + lane;
    bool hit = false;
    if (b < b1 && b != a) hit = within(pos, stride, ax, ay, az, b, r_max) || pair_extra(crit, a, b, g, b0, b1 - b0);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (FILL) {
      if (hit) {
        const int64_t p = cursor + __popc(m & ((1u << lane) - 1u));
        edge_index[p] = a;
        edge_index[n_edges + p] = b;
      }
      cursor += __popc(m);
    } else {
      count += __popc(m);
    }
  }
  if (!FILL && lane == 0) deg[a] = count;
}

static PairCrit to_crit(const e3b_pair_criteria* c) {
  PairCrit k;
  k.seg = c->segment; k.max_sep = c->max_separation; k.p = c->p_random; k.uniforms = c->uniforms;
  k.pair_ptr = c->pair_ptr; k.seed = c->seed;
  return k;
}

extern "C" int e3b_pair_graph_count(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                    int64_t n_nodes, float r_max, const e3b_pair_criteria* crit, int32_t* deg, void* stream) {
  if (n_nodes == 0) return E3B_OK;
  if (!pos || !node_ptr || !deg || !crit || n_graphs <= 0) return fail(E3B_ERR_INVALID, "pair_graph_count: null/empty argument");
  if (crit->uniforms && !crit->pair_ptr) return fail(E3B_ERR_INVALID, "pair_graph_count: uniforms need pair_ptr");
  pair_graph_kernel<false><<<blocks_for(n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
      pos, pos_stride, node_ptr, n_graphs, n_nodes, r_max, to_crit(crit), deg, nullptr, 0, nullptr);
  return check_launch("pair_graph_count");
}

extern "C" int e3b_pair_graph_fill(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                   int64_t n_nodes, float r_max, const e3b_pair_criteria* crit, const int64_t* row_ptr,
                                   int64_t n_edges, int64_t* edge_index, void* stream) {
  if (n_nodes == 0 || n_edges == 0) return E3B_OK;
  if (!pos || !node_ptr || !row_ptr || !edge_index || !crit) return fail(E3B_ERR_INVALID, "pair_graph_fill: null argument");
  pair_graph_kernel<true><<<blocks_for(n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
      pos, pos_stride, node_ptr, n_graphs, n_nodes, r_max, to_crit(crit), nullptr, row_ptr, n_edges, edge_index);
  return check_launch("pair_graph_fill");
}

// ---- cell list.  Per graph with >= min_nodes atoms: bounding box -> uniform grid with cells of >= 1.001 r_max
// (so atoms closer than r_max sit in adjacent cells whatever the fp32 rounding of the cell coordinate), at most
// cell_ptr[g+1] - cell_ptr[g] cells (the caller reserves ~2 cells per atom; the cells grow if the box is sparse).
// Atoms are binned with integer atomics (the order inside a cell is arbitrary); the sweep collects the hits of the
// 27 surrounding cells in shared memory and writes them in ascending b by rank, so the output is bit-identical to
// the all-pairs sweep.  Smaller graphs of the same batch take the all-pairs loop inside the same kernel.
struct CellGrid {
  float ox, oy, oz, ix, iy, iz;   // origin, cells per unit length
  int nx, ny, nz;                 // nx == 0: graph not binned
  int pad;
  long long base;                 // first cell of this graph
};
static_assert(sizeof(CellGrid) == E3B_CELL_GRID_BYTES, "e3b200.h: E3B_CELL_GRID_BYTES");

__global__ void __launch_bounds__(256) cell_grid_kernel(const float* __restrict__ pos, int64_t stride,
                                                        const int64_t* __restrict__ node_ptr, int n_graphs, float r_max,
                                                        int64_t min_nodes, const int64_t* __restrict__ cell_ptr,
                                                        CellGrid* __restrict__ grids) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= n_graphs) return;
  const int64_t b0 = node_ptr[g], b1 = node_ptr[g + 1];
  CellGrid cg;
  cg.ox = cg.oy = cg.oz = cg.ix = cg.iy = cg.iz = 0.f;
  cg.nx = cg.ny = cg.nz = 0; cg.pad = 0; cg.base = cell_ptr[g];
  const int64_t cap = cell_ptr[g + 1] - cell_ptr[g];
  if (b1 - b0 >= min_nodes && cap >= 1) {
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int64_t b = b0 + lane; b < b1; b += 32)
      for (int d = 0; d < 3; ++d) { const float v = pos[b * stride + d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
    for (int o = 16; o; o >>= 1)
      for (int d = 0; d < 3; ++d) {
        lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
        hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
      }
    int dcap = 1;
    while ((int64_t)(dcap + 1) * (dcap + 1) * (dcap + 1) <= cap && dcap < 1024) ++dcap;
    int n[3];
    float inv[3];
    const float cell = r_max * 1.001f;
    for (int d = 0; d < 3; ++d) {
      const float ext = hi[d] - lo[d];
      int k = ext > 0.f && cell > 0.f ? (int)fminf(floorf(ext / cell), (float)dcap) : 1;
      if (k < 1) k = 1;
      n[d] = k;
      inv[d] = ext > 0.f ? (float)k / ext : 0.f;
      // (x - lo) * inv <= k (1 + 2^-22); the coordinate is clamped to k - 1 in cell_coord
    }
    cg.ox = lo[0]; cg.oy = lo[1]; cg.oz = lo[2];
    cg.ix = inv[0]; cg.iy = inv[1]; cg.iz = inv[2];
    cg.nx = n[0]; cg.ny = n[1]; cg.nz = n[2];
  }
  if (lane == 0) grids[g] = cg;
}

__device__ __forceinline__ void cell_coord(const CellGrid& cg, float x, float y, float z, int& cx, int& cy, int& cz) {
  cx = min(cg.nx - 1, max(0, (int)((x - cg.ox) * cg.ix)));
  cy = min(cg.ny - 1, max(0, (int)((y - cg.oy) * cg.iy)));
  cz = min(cg.nz - 1, max(0, (int)((z - cg.oz) * cg.iz)));
}

__global__ void cell_assign_kernel(const float* __restrict__ pos, int64_t stride, const int64_t* __restrict__ node_ptr,
                                   int n_graphs, int64_t n_nodes, const CellGrid* __restrict__ grids,
                                   int32_t* __restrict__ cell_of, int32_t* __restrict__ cell_count) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_nodes) return;
  const int g = find_graph(node_ptr, n_graphs, a);
  const CellGrid cg = grids[g];
  if (cg.nx == 0) { cell_of[a] = -1; return; }
  int cx, cy, cz;
  cell_coord(cg, pos[a * stride], pos[a * stride + 1], pos[a * stride + 2], cx, cy, cz);
  const int32_t c = (int32_t)(cg.base + ((long long)cx * cg.ny + cy) * cg.nz + cz);
  cell_of[a] = c;
  atomicAdd(&cell_count[c], 1);
}

__global__ void cell_scatter_kernel(const int32_t* __restrict__ cell_of, int64_t n_nodes, const int64_t* __restrict__ cell_start,
                                    int32_t* __restrict__ cursor, int32_t* __restrict__ sorted) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_nodes) return;
  const int32_t c = cell_of[a];
  if (c < 0) return;
  sorted[cell_start[c] + atomicAdd(&cursor[c], 1)] = (int32_t)a;
}

constexpr int CELL_BUF = 512;   // hits buffered per warp in the fill pass (beyond: all-pairs loop for that atom)

template <bool FILL>
__global__ void __launch_bounds__(256) cell_graph_kernel(const float* __restrict__ pos, int64_t stride,
                                                         const int64_t* __restrict__ node_ptr, int n_graphs,
                                                         int64_t n_nodes, float r_max, const CellGrid* __restrict__ grids,
                                                         const int64_t* __restrict__ cell_start,
                                                         const int32_t* __restrict__ sorted, int32_t* __restrict__ deg,
                                                         const int64_t* __restrict__ row_ptr, int64_t n_edges,
                                                         int64_t* __restrict__ edge_index) {
  __shared__ int32_t buf_all[FILL ? 8 * CELL_BUF : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (a >= n_nodes) return;
  const int g = find_graph(node_ptr, n_graphs, a);
  const int64_t b0 = node_ptr[g], b1 = node_ptr[g + 1];
  const CellGrid cg = grids[g];
  const float ax = pos[a * stride], ay = pos[a * stride + 1], az = pos[a * stride + 2];
  const int64_t first = FILL ? row_ptr[a] : 0;
  const int64_t want = FILL ? row_ptr[a + 1] - first : 0;
  if (cg.nx == 0 || (FILL && want > CELL_BUF)) {      // all-pairs sweep (small graph, or more hits than the buffer holds)
    int64_t cursor = first;
    int count = 0;
    for (int64_t base = b0; base < b1; base += 32) {
      const int64_t b = base + lane;
      bool hit = false;
      if (b < b1 && b != a) hit = within(pos, stride, ax, ay, az, b, r_max);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (FILL) {
        if (hit) {
          const int64_t p = cursor + __popc(m & ((1u << lane) - 1u));
          edge_index[p] = a;
          edge_index[n_edges + p] = b;
        }
        cursor += __popc(m);
      } else {
        count += __popc(m);
      }
    }
    if (!FILL && lane == 0) deg[a] = count;
    return;
  }
  int32_t* buf = buf_all + (FILL ? warp * CELL_BUF : 0);
  int cx, cy, cz;
  cell_coord(cg, ax, ay, az, cx, cy, cz);
  int count = 0;
  for (int dx = -1; dx <= 1; ++dx) {
    const int x = cx + dx;
    if (x < 0 || x >= cg.nx) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = cy + dy;
      if (y < 0 || y >= cg.ny) continue;
      // the three cells z-1, z, z+1 are consecutive in memory: one contiguous range of `sorted`
      const int z0 = max(cz - 1, 0), z1 = min(cz + 1, cg.nz - 1);
      const long long c0 = cg.base + ((long long)x * cg.ny + y) * cg.nz;
      const int64_t i0 = cell_start[c0 + z0], i1 = cell_start[c0 + z1 + 1];
      for (int64_t base = i0; base < i1; base += 32) {
        const int64_t i = base + lane;
        int64_t b = -1;
        bool hit = false;
        if (i < i1) {
          b = sorted[i];
          if (b != a) hit = within(pos, stride, ax, ay, az, b, r_max);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (FILL && hit) buf[count + __popc(m & ((1u << lane) - 1u))] = (int32_t)b;
        count += __popc(m);
      }
    }
  }
  if (!FILL) {
    if (lane == 0) deg[a] = count;
    return;
  }
  __syncwarp();
  // rank = number of smaller hits (they are distinct): ascending b, the reference's order
  for (int i = lane; i < count; i += 32) {
    const int32_t v = buf[i];
    int r = 0;
    for (int j = 0; j < count; ++j) r += buf[j] < v;
    edge_index[first + r] = a;
    edge_index[n_edges + first + r] = v;
  }
}

extern "C" int e3b_cell_graph_bin(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                  int64_t n_nodes, float r_max, int64_t min_nodes, const int64_t* cell_ptr, void* grids,
                                  int32_t* cell_of, int32_t* cell_count, void* stream) {
  if (n_nodes == 0) return E3B_OK;
  if (!pos || !node_ptr || !cell_ptr || !grids || !cell_of || !cell_count || n_graphs <= 0)
    return fail(E3B_ERR_INVALID, "cell_graph_bin: null/empty argument");
  cell_grid_kernel<<<blocks_for(n_graphs, 8), 256, 0, (cudaStream_t)stream>>>(pos, pos_stride, node_ptr, n_graphs, r_max,
                                                                             min_nodes, cell_ptr, (CellGrid*)grids);
  int rc = check_launch("cell_grid");
  if (rc) return rc;
  cell_assign_kernel<<<blocks_for(n_nodes, 256), 256, 0, (cudaStream_t)stream>>>(pos, pos_stride, node_ptr, n_graphs, n_nodes,
                                                                                 (const CellGrid*)grids, cell_of, cell_count);
  return check_launch("cell_assign");
}

extern "C" int e3b_cell_graph_sort(const int32_t* cell_of, int64_t n_nodes, const int64_t* cell_start, int32_t* cursor,
                                   int32_t* sorted, void* stream) {
  if (n_nodes == 0) return E3B_OK;
  if (!cell_of || !cell_start || !cursor || !sorted) return fail(E3B_ERR_INVALID, "cell_graph_sort: null argument");
  cell_scatter_kernel<<<blocks_for(n_nodes, 256), 256, 0, (cudaStream_t)stream>>>(cell_of, n_nodes, cell_start, cursor, sorted);
  return check_launch("cell_scatter");
}

extern "C" int e3b_cell_graph_count(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                    int64_t n_nodes, float r_max, const void* grids, const int64_t* cell_start,
                                    const int32_t* sorted, int32_t* deg, void* stream) {
  if (n_nodes == 0) return E3B_OK;
  if (!pos || !node_ptr || !grids || !cell_start || !sorted || !deg) return fail(E3B_ERR_INVALID, "cell_graph_count: null argument");
  cell_graph_kernel<false><<<blocks_for(n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
      pos, pos_stride, node_ptr, n_graphs, n_nodes, r_max, (const CellGrid*)grids, cell_start, sorted, deg, nullptr, 0, nullptr);
  return check_launch("cell_graph_count");
}

extern "C" int e3b_cell_graph_fill(const float* pos, int64_t pos_stride, const int64_t* node_ptr, int32_t n_graphs,
                                   int64_t n_nodes, float r_max, const void* grids, const int64_t* cell_start,
                                   const int32_t* sorted, const int64_t* row_ptr, int64_t n_edges, int64_t* edge_index,
                                   int32_t* rev, void* stream) {
  if (n_nodes == 0 || n_edges == 0) return E3B_OK;
  if (!pos || !node_ptr || !grids || !cell_start || !sorted || !row_ptr || !edge_index)
    return fail(E3B_ERR_INVALID, "cell_graph_fill: null argument");
  cell_graph_kernel<true><<<blocks_for(n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
      pos, pos_stride, node_ptr, n_graphs, n_nodes, r_max, (const CellGrid*)grids, cell_start, sorted, nullptr, row_ptr,
      n_edges, edge_index);
  int rc = check_launch("cell_graph_fill");
  if (rc) return rc;
  if (rev) {
    reverse_edge_kernel<<<blocks_for(n_edges, 256), 256, 0, (cudaStream_t)stream>>>(edge_index, row_ptr, n_edges, n_edges, rev,
                                                                                  nullptr);
    rc = check_launch("radius_graph_reverse");
  }
  return rc;
}

// Grouped CSR of an arbitrary edge list: node n's segment lists, in ascending edge id, the
// edges whose `which_row` endpoint is n.  One warp per node sweeps... no: the edge list is
// unsorted, so: pass 1 scatter with an atomic cursor, pass 2 per-segment sort (ascending id).
__global__ void csr_scatter_kernel(const int64_t* __restrict__ key, int64_t n_edges, const int64_t* __restrict__ row_ptr,
                                   int32_t* __restrict__ cursor, int32_t* __restrict__ eid) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int64_t n = key[e];
  const int32_t slot = atomicAdd(&cursor[n], 1);
  eid[row_ptr[n] + slot] = (int32_t)e;
}

__global__ void csr_sort_kernel(const int64_t* __restrict__ row_ptr, int64_t n_nodes, int32_t* __restrict__ eid) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const int64_t a = row_ptr[n], b = row_ptr[n + 1];
  // Shell sort (gaps 57, 23, 10, 4, 1): segments are tens to a few hundred entries.
  const int gaps[5] = {57, 23, 10, 4, 1};
  for (int gi = 0; gi < 5; ++gi) {
    const int64_t gap = gaps[gi];
    for (int64_t i = a + gap; i < b; ++i) {
      const int32_t v = eid[i];
      int64_t j = i;
      while (j - gap >= a && eid[j - gap] > v) { eid[j] = eid[j - gap]; j -= gap; }
      eid[j] = v;
    }
  }
}

extern "C" int e3b_csr_fill(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, int which_row,
                            const int64_t* row_ptr, int32_t* cursor, int32_t* eid, void* stream) {
  if (n_edges == 0 || n_nodes == 0) return E3B_OK;
  if (!edge_index || !row_ptr || !cursor || !eid || (which_row != 0 && which_row != 1))
    return fail(E3B_ERR_INVALID, "csr_fill: bad argument");
  const int64_t* key = edge_index + (int64_t)which_row * n_edges;
  csr_scatter_kernel<<<blocks_for(n_edges, 256), 256, 0, (cudaStream_t)stream>>>(key, n_edges, row_ptr, cursor, eid);
  int rc = check_launch("csr_scatter");
  if (rc) return rc;
  csr_sort_kernel<<<blocks_for(n_nodes, 128), 128, 0, (cudaStream_t)stream>>>(row_ptr, n_nodes, eid);
  return check_launch("csr_sort");
}

// ------------------------------------------------------------------------------------------
// Edge vectors (data/compute_edge.py:13-36)
template <typename T> __device__ __forceinline__ T sqrt_(T v);
template <> __device__ __forceinline__ float sqrt_<float>(float v) { return sqrtf(v); }
template <> __device__ __forceinline__ double sqrt_<double>(double v) { return sqrt(v); }

template <typename T>
__global__ void edge_vectors_fwd_kernel(const T* __restrict__ pos, const int64_t* __restrict__ ei, int64_t n_edges,
                                        T* __restrict__ vec, T* __restrict__ len) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int64_t s = ei[e], d = ei[n_edges + e];
  const T x = pos[d * 3] - pos[s * 3], y = pos[d * 3 + 1] - pos[s * 3 + 1], z = pos[d * 3 + 2] - pos[s * 3 + 2];
  vec[e * 3] = x; vec[e * 3 + 1] = y; vec[e * 3 + 2] = z;
  if (len) len[e] = sqrt_<T>(fma_(z, z, fma_(y, y, x * x)));
}

template <typename T>
__global__ void edge_vectors_bwd_kernel(const T* __restrict__ gvec, const T* __restrict__ glen, const T* __restrict__ vec,
                                        const T* __restrict__ len, int64_t n_nodes, const int64_t* __restrict__ in_ptr,
                                        const int32_t* __restrict__ in_eid, const int64_t* __restrict__ out_ptr,
                                        const int32_t* __restrict__ out_eid, T* __restrict__ gpos) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  T ax = 0, ay = 0, az = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int64_t* ptr = pass == 0 ? in_ptr : out_ptr;
    const int32_t* ids = pass == 0 ? in_eid : out_eid;
    const T sgn = pass == 0 ? T(1) : T(-1);  // vec = pos[dst] - pos[src]
    for (int64_t k = ptr[n]; k < ptr[n + 1]; ++k) {
      const int64_t e = ids ? (int64_t)ids[k] : k;
      T gx = 0, gy = 0, gz = 0;
      if (gvec) { gx = gvec[e * 3]; gy = gvec[e * 3 + 1]; gz = gvec[e * 3 + 2]; }
      if (glen) {
        const T l = len[e];
        const T f = l > T(0) ? glen[e] / l : T(0);
        gx = fma_(f, vec[e * 3], gx); gy = fma_(f, vec[e * 3 + 1], gy); gz = fma_(f, vec[e * 3 + 2], gz);
      }
      ax = fma_(sgn, gx, ax); ay = fma_(sgn, gy, ay); az = fma_(sgn, gz, az);
    }
  }
  gpos[n * 3] = ax; gpos[n * 3 + 1] = ay; gpos[n * 3 + 2] = az;
}

extern "C" int e3b_edge_vectors_fwd(int dtype, const void* pos, const int64_t* edge_index, int64_t n_edges, void* vec,
                                    void* len, void* stream) {
  if (n_edges == 0) return E3B_OK;
  if (!pos || !edge_index || !vec) return fail(E3B_ERR_INVALID, "edge_vectors_fwd: null argument");
  DISPATCH_DTYPE(dtype, edge_vectors_fwd_kernel<T><<<blocks_for(n_edges, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)pos, edge_index, n_edges, (T*)vec, (T*)len);)
  return check_launch("edge_vectors_fwd");
}

extern "C" int e3b_edge_vectors_bwd(int dtype, const void* gvec, const void* glen, const void* vec, const void* len,
                                    int64_t n_nodes, const int64_t* in_ptr, const int32_t* in_eid,
                                    const int64_t* out_ptr, const int32_t* out_eid, void* gpos, void* stream) {
  if (n_nodes == 0) return E3B_OK;
  if (!in_ptr || !out_ptr || !gpos || (glen && (!vec || !len)))
    return fail(E3B_ERR_INVALID, "edge_vectors_bwd: null argument");
  DISPATCH_DTYPE(dtype, edge_vectors_bwd_kernel<T><<<blocks_for(n_nodes, 128), 128, 0, (cudaStream_t)stream>>>(
                            (const T*)gvec, (const T*)glen, (const T*)vec, (const T*)len, n_nodes, in_ptr, in_eid,
                            out_ptr, out_eid, (T*)gpos);)
  return check_launch("edge_vectors_bwd");
}

// ------------------------------------------------------------------------------------------
// Spherical harmonics, l <= 2, 'component' normalisation (nn/embedding.py:163-165)
template <typename T>
__global__ void sh_fwd_kernel(const T* __restrict__ vec, int64_t n, int lmax, int normalize, T* __restrict__ sh) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  T x = vec[e * 3], y = vec[e * 3 + 1], z = vec[e * 3 + 2];
  if (normalize) {
    T r = sqrt_<T>(x * x + y * y + z * z);
    r = r > T(1e-12) ? r : T(1e-12);
    x /= r; y /= r; z /= r;
  }
  const int dim = (lmax + 1) * (lmax + 1);
  T* o = sh + e * dim;
  o[0] = T(1);
  if (lmax >= 1) {
    const T s3 = T(1.7320508075688772);
    o[1] = s3 * x; o[2] = s3 * y; o[3] = s3 * z;
  }
  if (lmax >= 2) {
    const T s3 = T(1.7320508075688772), s5 = T(2.23606797749979);
    o[4] = s5 * s3 * x * z;
    o[5] = s5 * s3 * x * y;
    o[6] = s5 * (y * y - T(0.5) * (x * x + z * z));
    o[7] = s5 * s3 * y * z;
    o[8] = s5 * (s3 / T(2)) * (z * z - x * x);
  }
}

template <typename T>
__global__ void sh_bwd_kernel(const T* __restrict__ vec, const T* __restrict__ gsh, int64_t n, int lmax, int normalize,
                              T* __restrict__ gvec) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  T x = vec[e * 3], y = vec[e * 3 + 1], z = vec[e * 3 + 2];
  T r = T(1);
  bool clamped = false;
  if (normalize) {
    r = sqrt_<T>(x * x + y * y + z * z);
    if (!(r > T(1e-12))) { r = T(1e-12); clamped = true; }
    x /= r; y /= r; z /= r;
  }
  const int dim = (lmax + 1) * (lmax + 1);
  const T* g = gsh + e * dim;
  T gx = 0, gy = 0, gz = 0;
  if (lmax >= 1) {
    const T s3 = T(1.7320508075688772);
    gx += s3 * g[1]; gy += s3 * g[2]; gz += s3 * g[3];
  }
  if (lmax >= 2) {
    const T s3 = T(1.7320508075688772), s5 = T(2.23606797749979);
    gx += s5 * (s3 * z * g[4] + s3 * y * g[5] - x * g[6] - s3 * x * g[8]);
    gy += s5 * (s3 * x * g[5] + T(2) * y * g[6] + s3 * z * g[7]);
    gz += s5 * (s3 * x * g[4] - z * g[6] + s3 * y * g[7] + s3 * z * g[8]);
  }
  if (normalize) {
    if (clamped) {  // n = v / 1e-12: plain scaling
      gx /= r; gy /= r; gz /= r;
    } else {        // d(v/|v|): (g - n (n.g)) / |v|
      const T dot = gx * x + gy * y + gz * z;
      gx = (gx - x * dot) / r; gy = (gy - y * dot) / r; gz = (gz - z * dot) / r;
    }
  }
  gvec[e * 3] = gx; gvec[e * 3 + 1] = gy; gvec[e * 3 + 2] = gz;
}

extern "C" int e3b_sh_fwd(int dtype, const void* vec, int64_t n, int lmax, int normalize, void* sh, void* stream) {
  if (n == 0) return E3B_OK;
  if (lmax < 0 || lmax > 2) return fail(E3B_ERR_UNSUPPORTED, "sh: lmax %d not supported (0..2)", lmax);
  if (!vec || !sh) return fail(E3B_ERR_INVALID, "sh_fwd: null argument");
  DISPATCH_DTYPE(dtype, sh_fwd_kernel<T><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)vec, n, lmax,
                                                                                              normalize, (T*)sh);)
  return check_launch("sh_fwd");
}

extern "C" int e3b_sh_bwd(int dtype, const void* vec, const void* gsh, int64_t n, int lmax, int normalize, void* gvec,
                          void* stream) {
  if (n == 0) return E3B_OK;
  if (lmax < 0 || lmax > 2) return fail(E3B_ERR_UNSUPPORTED, "sh: lmax %d not supported (0..2)", lmax);
  if (!vec || !gsh || !gvec) return fail(E3B_ERR_INVALID, "sh_bwd: null argument");
  DISPATCH_DTYPE(dtype, sh_bwd_kernel<T><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)vec, (const T*)gsh, n, lmax, normalize, (T*)gvec);)
  return check_launch("sh_bwd");
}

// ------------------------------------------------------------------------------------------
// Radial basis (nn/embedding.py:181-219, 74-127, 26-40)
template <typename T> __device__ __forceinline__ void sincos_(T a, T* s, T* c);
template <> __device__ __forceinline__ void sincos_<float>(float a, float* s, float* c) { sincosf(a, s, c); }
template <> __device__ __forceinline__ void sincos_<double>(double a, double* s, double* c) { sincos(a, s, c); }
template <typename T> __device__ __forceinline__ T pow_(T a, T b);
template <> __device__ __forceinline__ float pow_<float>(float a, float b) { return powf(a, b); }
template <> __device__ __forceinline__ double pow_<double>(double a, double b) { return pow(a, b); }

struct RadialParams {
  double r_max, r_min, p;
  int n_basis, one_over_r, cutoff_kind;
};

template <typename T>
__device__ __forceinline__ void cutoff_eval(T r, const RadialParams& P, T* c, T* dc) {
  const T f = T(1.0 / P.r_max);
  const T x = r * f;
  if (P.cutoff_kind == 1) {  // symmetricCutoff: (x-1)^2 (x+1)^2 [|x| < 1]
    if (fabs((double)x) < 1.0) {
      const T q = (x - T(1)) * (x + T(1));
      *c = q * q;
      *dc = T(4) * x * q * f;
    } else { *c = T(0); *dc = T(0); }
    return;
  }
  if (x < T(1)) {
    const T p = T(P.p);
    const T xp = pow_<T>(x, p);          // x^p
    const T xm = xp / x;                 // x^(p-1)  (x > 0 for distances)
    *c = T(1) - (p + T(1)) * (p + T(2)) / T(2) * xp + p * (p + T(2)) * xp * x - p * (p + T(1)) / T(2) * xp * x * x;
    *dc = (x > T(0)) ? f * p * (p + T(1)) * (p + T(2)) / T(2) * (-xm + T(2) * xp - xp * x) : T(0);
  } else { *c = T(0); *dc = T(0); }
}

template <typename T>
__global__ void radial_fwd_kernel(const T* __restrict__ r_in, int64_t n, const T* __restrict__ bw, RadialParams P,
                                  T* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * P.n_basis) return;
  const int64_t e = idx / P.n_basis;
  const int k = (int)(idx - e * P.n_basis);
  const T r = r_in[e];
  T c, dc;
  cutoff_eval<T>(r, P, &c, &dc);
  const T span = T(P.r_max - P.r_min);
  T s, co;
  sincos_<T>(bw[k] * r / span, &s, &co);
  T b = T(2.0) / span * s;
  if (P.one_over_r) b = b / r;
  out[idx] = b * c;
}

#define RADIAL_BWD_ROWS 256  // edges per block in the backward
template <typename T>
__global__ void __launch_bounds__(256) radial_bwd_kernel(const T* __restrict__ r_in, const T* __restrict__ gout,
                                                         int64_t n, const T* __restrict__ bw, RadialParams P,
                                                         T* __restrict__ gr, T* __restrict__ gw_partial) {
  // thread t handles edge blockIdx*256 + t (all basis functions); gw reduced over the block
  extern __shared__ unsigned char smem_raw[];
  T* red = reinterpret_cast<T*>(smem_raw);  // [8 warps][n_basis]
  const int64_t e = (int64_t)blockIdx.x * RADIAL_BWD_ROWS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool valid = e < n;
  T r = valid ? r_in[e] : T(1);
  T c = 0, dc = 0;
  if (valid) cutoff_eval<T>(r, P, &c, &dc);
  const T span = T(P.r_max - P.r_min);
  const T pref = T(2.0) / span;
  T g_r = 0;
  for (int k = 0; k < P.n_basis; ++k) {
    T gwk = 0;
    if (valid) {
      const T a = bw[k] / span;
      T s, co;
      sincos_<T>(a * r, &s, &co);
      const T go = gout[e * P.n_basis + k];
      T b, db, dbw;
      if (P.one_over_r) {
        b = pref * s / r;
        db = pref * (a * co / r - s / (r * r));
        dbw = pref * co / span;  // d/dw [sin(w r/span)/r] = cos * (r/span)/r
      } else {
        b = pref * s;
        db = pref * a * co;
        dbw = pref * co * r / span;
      }
      g_r += go * (db * c + b * dc);
      gwk = go * dbw * c;
    }
    gwk = warp_sum(gwk);
    if (lane == 0) red[warp * P.n_basis + k] = gwk;
  }
  if (valid && gr) gr[e] = g_r;
  __syncthreads();
  if (gw_partial) {
    for (int k = threadIdx.x; k < P.n_basis; k += blockDim.x) {
      T s = 0;
      for (int w = 0; w < 8; ++w) s += red[w * P.n_basis + k];
      gw_partial[(int64_t)blockIdx.x * P.n_basis + k] = s;
    }
  }
}

extern "C" int e3b_radial_fwd(int dtype, const void* r, int64_t n, const void* bessel_w, int n_basis, double r_max,
                              double r_min, int one_over_r, int cutoff_kind, double p, void* out, void* stream) {
  if (n == 0) return E3B_OK;
  if (!r || !bessel_w || !out || n_basis <= 0) return fail(E3B_ERR_INVALID, "radial_fwd: bad argument");
  RadialParams P{r_max, r_min, p, n_basis, one_over_r, cutoff_kind};
  DISPATCH_DTYPE(dtype, radial_fwd_kernel<T><<<blocks_for(n * n_basis, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)r, n, (const T*)bessel_w, P, (T*)out);)
  return check_launch("radial_fwd");
}

extern "C" int64_t e3b_radial_bwd_blocks(int64_t n) { return (n + RADIAL_BWD_ROWS - 1) / RADIAL_BWD_ROWS; }

extern "C" int e3b_radial_bwd(int dtype, const void* r, const void* gout, int64_t n, const void* bessel_w, int n_basis,
                              double r_max, double r_min, int one_over_r, int cutoff_kind, double p, void* gr,
                              void* gw_partial, void* stream) {
  if (n == 0) return E3B_OK;
  if (!r || !gout || !bessel_w || n_basis <= 0) return fail(E3B_ERR_INVALID, "radial_bwd: bad argument");
  RadialParams P{r_max, r_min, p, n_basis, one_over_r, cutoff_kind};
  const unsigned grid = (unsigned)e3b_radial_bwd_blocks(n);
  DISPATCH_DTYPE(dtype, radial_bwd_kernel<T><<<grid, 256, 8 * n_basis * sizeof(T), (cudaStream_t)stream>>>(
                            (const T*)r, (const T*)gout, n, (const T*)bessel_w, P, (T*)gr, (T*)gw_partial);)
  return check_launch("radial_bwd");
}

// ------------------------------------------------------------------------------------------
// Tensor-product convolution: plan + generic kernels (any multiplicities, l <= 3, f32/f64).
struct e3b_tp_plan {
  e3b_tp_desc desc;
  const GenEntry* gen;             // matching generated kernel or null
  int32_t mul[E3B_MAX_BLOCKS];     // generic: per input block multiplicity (== desc.mul)
  int64_t x_dim, sh_dim, w_dim, y_dim;
  int32_t x_off[E3B_MAX_BLOCKS], sh_off[E3B_MAX_BLOCKS];
  int32_t w_off[E3B_MAX_PATHS], y_off[E3B_MAX_PATHS], y_kstride[E3B_MAX_PATHS];
  float sign[E3B_MAX_PATHS];
};

struct GenericTp {
  int32_t n_paths, mul;
  int32_t l1[E3B_MAX_PATHS], l2[E3B_MAX_PATHS], l3[E3B_MAX_PATHS];
  int32_t x_off[E3B_MAX_PATHS], sh_off[E3B_MAX_PATHS], w_off[E3B_MAX_PATHS], y_off[E3B_MAX_PATHS], y_kstride[E3B_MAX_PATHS];
  float sign[E3B_MAX_PATHS];
};

extern "C" int e3b_tp_plan_create(const e3b_tp_desc* d, e3b_tp_plan** out) {
  if (!d || !out) return fail(E3B_ERR_INVALID, "tp_plan_create: null argument");
  if (d->mul <= 0 || d->n_in <= 0 || d->n_in > E3B_MAX_BLOCKS || d->n_sh <= 0 || d->n_sh > E3B_MAX_BLOCKS ||
      d->n_paths <= 0 || d->n_paths > E3B_MAX_PATHS)
    return fail(E3B_ERR_INVALID, "tp_plan_create: sizes out of range (mul %d, n_in %d, n_sh %d, n_paths %d)", d->mul,
                d->n_in, d->n_sh, d->n_paths);
  e3b_tp_plan* p = new (std::nothrow) e3b_tp_plan();
  if (!p) return fail(E3B_ERR_NOMEM, "tp_plan_create: out of host memory");
  p->desc = *d;
  int off = 0;
  for (int b = 0; b < d->n_in; ++b) {
    if (d->in_l[b] < 0 || d->in_l[b] > E3B_CG_LMAX) { delete p; return fail(E3B_ERR_UNSUPPORTED, "input l=%d > %d", d->in_l[b], E3B_CG_LMAX); }
    p->x_off[b] = off;
    off += 2 * d->in_l[b] + 1;
  }
  p->x_dim = (int64_t)off * d->mul;
  off = 0;
  for (int s = 0; s < d->n_sh; ++s) {
    if (d->sh_l[s] < 0 || d->sh_l[s] > E3B_CG_LMAX) { delete p; return fail(E3B_ERR_UNSUPPORTED, "sh l=%d > %d", d->sh_l[s], E3B_CG_LMAX); }
    p->sh_off[s] = off;
    off += 2 * d->sh_l[s] + 1;
  }
  p->sh_dim = off;
  p->w_dim = (int64_t)d->n_paths * d->mul;
  // output blocks by slot; consecutive slots with the same (l, parity) form one group laid out
  // [k][slot-in-group][u]
  int slot_l[E3B_MAX_PATHS], slot_p[E3B_MAX_PATHS];
  for (int q = 0; q < d->n_paths; ++q) slot_l[q] = -1;
  for (int q = 0; q < d->n_paths; ++q) {
    const int b = d->path_in[q], s = d->path_sh[q], l3 = d->path_lout[q], slot = d->path_slot[q];
    if (b < 0 || b >= d->n_in || s < 0 || s >= d->n_sh || slot < 0 || slot >= d->n_paths || slot_l[slot] != -1 ||
        l3 > E3B_CG_LMAX || l3 < abs(d->in_l[b] - d->sh_l[s]) || l3 > d->in_l[b] + d->sh_l[s]) {
      delete p;
      return fail(E3B_ERR_INVALID, "tp_plan_create: bad path %d (in %d, sh %d, l_out %d, slot %d)", q, b, s, l3, slot);
    }
    slot_l[slot] = l3;
    slot_p[slot] = d->in_p[b] * d->sh_p[s];
  }
  int ybase[E3B_MAX_PATHS], ykst[E3B_MAX_PATHS];
  off = 0;
  for (int s = 0; s < d->n_paths;) {
    int e = s;
    while (e < d->n_paths && slot_l[e] == slot_l[s] && slot_p[e] == slot_p[s]) ++e;
    for (int t = s; t < e; ++t) { ybase[t] = off + (t - s); ykst[t] = e - s; }
    off += (e - s) * (2 * slot_l[s] + 1);
    s = e;
  }
  p->y_dim = (int64_t)off * d->mul;
  const int n = E3B_CG_LMAX + 1;
  for (int q = 0; q < d->n_paths; ++q) {
    p->w_off[q] = q;
    p->y_off[q] = ybase[d->path_slot[q]];
    p->y_kstride[q] = ykst[d->path_slot[q]];
    const int l1 = d->in_l[d->path_in[q]], l2 = d->sh_l[d->path_sh[q]], l3 = d->path_lout[q];
    p->sign[q] = (d->w3j_sign_preset == 1) ? (float)kCgSign044[(l1 * n + l2) * n + l3] : 1.0f;
  }
  p->gen = (d->w3j_sign_preset == 0) ? e3b_find_generated(d, p->y_off, p->y_kstride) : nullptr;
  *out = p;
  return E3B_OK;
}

extern "C" void e3b_tp_plan_destroy(e3b_tp_plan* p) { delete p; }
extern "C" int e3b_tp_plan_is_specialized(const e3b_tp_plan* p) { return p && p->gen ? 1 : 0; }
extern "C" int e3b_tp_plan_dims(const e3b_tp_plan* p, int32_t* x_dim, int32_t* sh_dim, int32_t* w_dim, int32_t* y_dim,
                                int32_t* n_part_f32) {
  if (!p) return fail(E3B_ERR_INVALID, "tp_plan_dims: null plan");
  if (x_dim) *x_dim = (int32_t)p->x_dim;
  if (sh_dim) *sh_dim = (int32_t)p->sh_dim;
  if (w_dim) *w_dim = (int32_t)p->w_dim;
  if (y_dim) *y_dim = (int32_t)p->y_dim;
  if (n_part_f32) *n_part_f32 = p->gen ? e3b_gen_bwd_parts(p->gen, p->desc.mul) : 1;
  return E3B_OK;
}

static GenericTp make_generic(const e3b_tp_plan* p) {
  GenericTp g;
  g.n_paths = p->desc.n_paths;
  g.mul = p->desc.mul;
  for (int q = 0; q < g.n_paths; ++q) {
    const int b = p->desc.path_in[q], s = p->desc.path_sh[q];
    g.l1[q] = p->desc.in_l[b]; g.l2[q] = p->desc.sh_l[s]; g.l3[q] = p->desc.path_lout[q];
    g.x_off[q] = p->x_off[b]; g.sh_off[q] = p->sh_off[s]; g.w_off[q] = p->w_off[q]; g.y_off[q] = p->y_off[q]; g.y_kstride[q] = p->y_kstride[q];
    g.sign[q] = p->sign[q];
  }
  return g;
}

__device__ __forceinline__ const double* cg_table(int l1, int l2, int l3) {
  const int n = E3B_CG_LMAX + 1;
  return kCgData + kCgOffset[(l1 * n + l2) * n + l3];
}

// one thread per (node, path, channel)
template <typename T>
__global__ void __launch_bounds__(128) tp_generic_fwd_kernel(const __grid_constant__ GenericTp g, TpArgs<T> a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_node = (int64_t)g.n_paths * g.mul;
  if (idx >= a.n_nodes * per_node) return;
  const int64_t node = idx / per_node;
  const int rem = (int)(idx - node * per_node);
  const int q = rem / g.mul, u = rem - q * g.mul;
  const int l1 = g.l1[q], l2 = g.l2[q], l3 = g.l3[q];
  const int d1 = 2 * l1 + 1, d2 = 2 * l2 + 1, d3 = 2 * l3 + 1;
  const double* C = cg_table(l1, l2, l3);
  const T scale = T(sqrt((double)d3)) * T(g.sign[q]);
  T acc[2 * E3B_CG_LMAX + 1];
  for (int k = 0; k < d3; ++k) acc[k] = T(0);
  for (int64_t kk = a.in_ptr[node]; kk < a.in_ptr[node + 1]; ++kk) {
    const int64_t src = a.in_nbr[kk];
    const int64_t eid = a.in_eid ? (int64_t)a.in_eid[kk] : kk;
    const T* xr = a.x + src * a.x_dim + (int64_t)g.x_off[q] * g.mul + u;
    const T* yr = a.sh + eid * a.sh_dim + g.sh_off[q];
    const T wv = a.w[eid * a.w_dim + (int64_t)g.w_off[q] * g.mul + u] * scale;
    for (int i = 0; i < d1; ++i) {
      const T xi = xr[(int64_t)i * g.mul] * wv;
      for (int j = 0; j < d2; ++j) {
        const T xy = xi * yr[j];
        const double* c = C + (i * d2 + j) * d3;
        for (int k = 0; k < d3; ++k) acc[k] = fma_(T(c[k]), xy, acc[k]);
      }
    }
  }
  T* yo = a.y + node * a.y_dim + (int64_t)g.y_off[q] * g.mul + u;
  for (int k = 0; k < d3; ++k) yo[(int64_t)k * g.y_kstride[q] * g.mul] = acc[k];
}

template <typename T>
__global__ void __launch_bounds__(128) tp_generic_bwd_kernel(const __grid_constant__ GenericTp g, TpArgs<T> a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_node = (int64_t)g.n_paths * g.mul;
  if (idx >= a.n_nodes * per_node) return;
  const int64_t node = idx / per_node;
  const int rem = (int)(idx - node * per_node);
  const int q = rem / g.mul, u = rem - q * g.mul;
  const int l1 = g.l1[q], l2 = g.l2[q], l3 = g.l3[q];
  const int d1 = 2 * l1 + 1, d2 = 2 * l2 + 1, d3 = 2 * l3 + 1;
  const double* C = cg_table(l1, l2, l3);
  const T scale = T(sqrt((double)d3)) * T(g.sign[q]);
  T gy[2 * E3B_CG_LMAX + 1];
  const T* gyr = a.gy + node * a.y_dim + (int64_t)g.y_off[q] * g.mul + u;
  for (int k = 0; k < d3; ++k) gy[k] = gyr[(int64_t)k * g.y_kstride[q] * g.mul];
  for (int64_t kk = a.in_ptr[node]; kk < a.in_ptr[node + 1]; ++kk) {
    const int64_t src = a.in_nbr[kk];
    const int64_t eid = a.in_eid ? (int64_t)a.in_eid[kk] : kk;
    const T* xr = a.x + src * a.x_dim + (int64_t)g.x_off[q] * g.mul + u;
    const T* yr = a.sh + eid * a.sh_dim + g.sh_off[q];
    const int64_t wi = eid * a.w_dim + (int64_t)g.w_off[q] * g.mul + u;
    const T wv = a.w[wi] * scale;
    T gw = T(0);
    for (int i = 0; i < d1; ++i) {
      const T xi = xr[(int64_t)i * g.mul];
      T gxi = T(0);
      for (int j = 0; j < d2; ++j) {
        const double* c = C + (i * d2 + j) * d3;
        T m = T(0);  // sum_k C_ijk gy_k
        for (int k = 0; k < d3; ++k) m = fma_(T(c[k]), gy[k], m);
        const T yj = yr[j];
        gw = fma_(m, xi * yj, gw);
        gxi = fma_(m, yj, gxi);
        if (a.gsh && m != T(0)) atomicAdd(a.gsh + eid * a.sh_dim + g.sh_off[q] + j, m * xi * wv);
      }
      if (a.gx_edge) atomicAdd(a.gx_edge + eid * a.x_dim + (int64_t)(g.x_off[q] + i) * g.mul + u, gxi * wv);
    }
    a.gw[wi] = gw * scale;
  }
}

template <typename T>
static TpArgs<T> make_args(const e3b_tp_plan* p, int64_t n_nodes, const void* x, const void* sh, const void* w,
                           const void* gy, const int64_t* in_ptr, const int32_t* in_nbr, const int32_t* in_eid, void* y,
                           void* gx_edge, void* gsh, void* gw) {
  TpArgs<T> a;
  a.x = (const T*)x; a.sh = (const T*)sh; a.w = (const T*)w; a.gy = (const T*)gy;
  a.y = (T*)y; a.gx_edge = (T*)gx_edge; a.gx_node = nullptr; a.gsh = (T*)gsh; a.gw = (T*)gw;
  a.in_ptr = in_ptr; a.in_nbr = in_nbr; a.in_eid = in_eid; a.w_idx = nullptr;
  a.n_nodes = n_nodes;
  a.x_dim = p->x_dim; a.sh_dim = p->sh_dim; a.w_dim = p->w_dim; a.y_dim = p->y_dim;
  a.mul = p->desc.mul;
  a.n_chunks = (p->desc.mul + 31) / 32;
  a.n_part = p->gen ? e3b_gen_bwd_parts(p->gen, p->desc.mul) : 1;
  a.n_stages = 0;
  return a;
}

extern "C" int e3b_tpconv_fwd(const e3b_tp_plan* plan, int dtype, int64_t n_nodes, int64_t n_edges, const void* x,
                              const void* sh, const void* w, const int64_t* in_ptr, const int32_t* in_nbr,
                              const int32_t* in_eid, void* y, void* stream) {
  if (!plan) return fail(E3B_ERR_INVALID, "tpconv_fwd: null plan");
  if (n_nodes == 0) return E3B_OK;
  if (!x || !in_ptr || !y || (n_edges > 0 && (!sh || !w || !in_nbr)))
    return fail(E3B_ERR_INVALID, "tpconv_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == E3B_F32 && plan->gen) {
    TpArgs<float> a = make_args<float>(plan, n_nodes, x, sh, w, nullptr, in_ptr, in_nbr, in_eid, y, nullptr, nullptr, nullptr);
    const int64_t items = n_nodes * a.n_part;
    plan->gen->fwd(a, (items + TP_THREADS / 32 - 1) / (TP_THREADS / 32), st);
    return check_launch("tpconv_fwd (generated)");
  }
  const GenericTp g = make_generic(plan);
  const int64_t threads = n_nodes * g.n_paths * g.mul;
  DISPATCH_DTYPE(dtype, TpArgs<T> a = make_args<T>(plan, n_nodes, x, sh, w, nullptr, in_ptr, in_nbr, in_eid, y, nullptr,
                                                   nullptr, nullptr);
                 tp_generic_fwd_kernel<T><<<blocks_for(threads, 128), 128, 0, st>>>(g, a);)
  return check_launch("tpconv_fwd (generic)");
}

bool e3b_tp_pipelined_enabled();   // tp_fast.cu

static bool pipelined_plan(const e3b_tp_plan* plan) {
  return plan->gen && (plan->desc.mul == 64 || plan->desc.mul == 32) && e3b_tp_pipelined_enabled();
}

extern "C" int e3b_tpconv_fwd_shared(const e3b_tp_plan* plan, int64_t n_nodes, int64_t n_edges, const float* x, const float* sh,
                                     const float* w, const int32_t* w_idx, const int64_t* in_ptr, const int32_t* in_nbr,
                                     const int32_t* in_eid, float* y, void* stream) {
  if (!plan) return fail(E3B_ERR_INVALID, "tpconv_fwd_shared: null plan");
  if (!pipelined_plan(plan))
    return fail(E3B_ERR_UNSUPPORTED, "tpconv_fwd_shared: needs a generated structure with multiplicity 32 or 64");
  if (n_nodes == 0) return E3B_OK;
  if (!x || !in_ptr || !y || !w_idx || (n_edges > 0 && (!sh || !w || !in_nbr))) return fail(E3B_ERR_INVALID, "tpconv_fwd_shared: null argument");
  TpArgs<float> a = make_args<float>(plan, n_nodes, x, sh, w, nullptr, in_ptr, in_nbr, in_eid, y, nullptr, nullptr, nullptr);
  a.w_idx = w_idx;
  plan->gen->fwd(a, 0, (cudaStream_t)stream);
  return check_launch("tpconv_fwd_shared (generated)");
}

extern "C" int e3b_tpconv_bwd_nodes(const e3b_tp_plan* plan, int64_t n_nodes, int64_t n_edges, const float* x, const float* sh,
                                    const float* w, const int32_t* w_idx, const float* gy, const int64_t* in_ptr,
                                    const int32_t* in_nbr, const int32_t* in_eid, float* gx_node, float* gsh, float* gw,
                                    void* stream) {
  if (!plan) return fail(E3B_ERR_INVALID, "tpconv_bwd_nodes: null plan");
  if (!pipelined_plan(plan))
    return fail(E3B_ERR_UNSUPPORTED, "tpconv_bwd_nodes: needs a generated structure with multiplicity 32 or 64");
  if (n_nodes == 0 || n_edges == 0) return E3B_OK;
  if (!x || !sh || !w || !gy || !in_ptr || !in_nbr || !gw || !gx_node) return fail(E3B_ERR_INVALID, "tpconv_bwd_nodes: null argument");
  if ((reinterpret_cast<uintptr_t>(gx_node) & 15) || ((plan->x_dim * 4) & 15))
    return fail(E3B_ERR_INVALID, "tpconv_bwd_nodes: gx_node rows must be 16-byte aligned");
  TpArgs<float> a = make_args<float>(plan, n_nodes, x, sh, w, gy, in_ptr, in_nbr, in_eid, nullptr, nullptr, gsh, gw);
  a.gx_node = gx_node;
  a.w_idx = w_idx;
  plan->gen->bwd(a, 0, (cudaStream_t)stream);
  return check_launch("tpconv_bwd_nodes (generated)");
}

extern "C" int e3b_tpconv_bwd(const e3b_tp_plan* plan, int dtype, int64_t n_nodes, int64_t n_edges, const void* x,
                              const void* sh, const void* w, const void* gy, const int64_t* in_ptr,
                              const int32_t* in_nbr, const int32_t* in_eid, void* gx_edge, void* gsh, void* gw,
                              void* stream) {
  if (!plan) return fail(E3B_ERR_INVALID, "tpconv_bwd: null plan");
  if (n_nodes == 0 || n_edges == 0) return E3B_OK;
  if (!x || !sh || !w || !gy || !in_ptr || !in_nbr || !gw) return fail(E3B_ERR_INVALID, "tpconv_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == E3B_F32 && plan->gen) {
    TpArgs<float> a = make_args<float>(plan, n_nodes, x, sh, w, gy, in_ptr, in_nbr, in_eid, nullptr, gx_edge, gsh, gw);
    const int64_t items = n_nodes * a.n_part;
    plan->gen->bwd(a, (items + TP_THREADS / 32 - 1) / (TP_THREADS / 32), st);
    return check_launch("tpconv_bwd (generated)");
  }
  const GenericTp g = make_generic(plan);
  const int64_t threads = n_nodes * g.n_paths * g.mul;
  DISPATCH_DTYPE(dtype, TpArgs<T> a = make_args<T>(plan, n_nodes, x, sh, w, gy, in_ptr, in_nbr, in_eid, nullptr, gx_edge,
                                                   gsh, gw);
                 tp_generic_bwd_kernel<T><<<blocks_for(threads, 128), 128, 0, st>>>(g, a);)
  return check_launch("tpconv_bwd (generic)");
}

// ------------------------------------------------------------------------------------------
// Segmented row sum (scatter-sum by sorted segments; nn/message_passing.py:109, nn/output.py:69)
template <typename T>
__global__ void segment_sum_kernel(const T* __restrict__ src, int64_t width, const int64_t* __restrict__ ptr,
                                   const int32_t* __restrict__ ids, int64_t n_out, T* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_out * width) return;
  const int64_t n = idx / width, c = idx - n * width;
  T acc = T(0);
  for (int64_t k = ptr[n]; k < ptr[n + 1]; ++k) {
    const int64_t row = ids ? (int64_t)ids[k] : k;
    acc += src[row * width + c];
  }
  out[idx] = acc;
}

extern "C" int e3b_segment_sum(int dtype, const void* src, int64_t width, const int64_t* ptr, const int32_t* ids,
                               int64_t n_out, void* out, void* stream) {
  if (n_out == 0 || width == 0) return E3B_OK;
  if (!ptr || !out) return fail(E3B_ERR_INVALID, "segment_sum: null argument");
  DISPATCH_DTYPE(dtype, segment_sum_kernel<T><<<blocks_for(n_out * width, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)src, width, ptr, ids, n_out, (T*)out);)
  return check_launch("segment_sum");
}

// ------------------------------------------------------------------------------------------
// Gate (e3nn nn.Gate as wired at nn/message_passing.py:191-207)
// fp32: SFU approximations (ex2 / lg2 based, relative error ~1e-6 on the values the gate sees: the gate is
// evaluated once per feature element and would otherwise spend most of its time in libm range reduction)
template <typename T> __device__ __forceinline__ T exp_(T v);
template <> __device__ __forceinline__ float exp_<float>(float v) { return __expf(v); }
template <> __device__ __forceinline__ double exp_<double>(double v) { return exp(v); }
template <typename T> __device__ __forceinline__ T tanh_(T v);
template <> __device__ __forceinline__ float tanh_<float>(float v) {
  const float a = fminf(fabsf(v), 15.f), e = __expf(2.f * a);          // tanh|v| = 1 - 2 / (e^{2|v|} + 1)
  return copysignf(1.f - __fdividef(2.f, e + 1.f), v);
}
template <> __device__ __forceinline__ double tanh_<double>(double v) { return tanh(v); }
template <typename T> __device__ __forceinline__ T log1p_(T v);
template <> __device__ __forceinline__ float log1p_<float>(float v) { return __logf(1.f + v); }
template <> __device__ __forceinline__ double log1p_<double>(double v) { return log1p(v); }

template <typename T>
__device__ __forceinline__ void act_eval(int code, T x, T* f, T* df) {
  switch (code) {
    case 1: {  // silu
      const T s = T(1) / (T(1) + exp_<T>(-x));
      *f = x * s; *df = s * (T(1) + x * (T(1) - s));
    } break;
    case 2: {  // tanh
      const T t = tanh_<T>(x);
      *f = t; *df = T(1) - t * t;
    } break;
    case 3: {  // shifted softplus (torch softplus threshold 20)
      const T sp = x > T(20) ? x : log1p_<T>(exp_<T>(x));
      *f = sp - T(0.6931471805599453);
      *df = T(1) / (T(1) + exp_<T>(-x));
    } break;
    case 4: {  // tanhlu = tanh(x) |x|
      const T t = tanh_<T>(x), ax = x < T(0) ? -x : x, sg = x > T(0) ? T(1) : (x < T(0) ? T(-1) : T(0));
      *f = t * ax; *df = (T(1) - t * t) * ax + t * sg;
    } break;
    case 5: {  // abs
      *f = x < T(0) ? -x : x; *df = x > T(0) ? T(1) : (x < T(0) ? T(-1) : T(0));
    } break;
    default: *f = x; *df = T(1);
  }
}

struct GateLayout {
  e3b_gate_desc d;
  int32_t n_scalars, n_gates, in_dim, out_dim;
};

static GateLayout gate_layout(const e3b_gate_desc* d) {
  GateLayout L;
  L.d = *d;
  L.n_scalars = 0; L.n_gates = 0;
  int gated = 0;
  for (int i = 0; i < d->n_scalar_blocks; ++i) L.n_scalars += d->scalar_mul[i];
  for (int i = 0; i < d->n_gated_blocks; ++i) { L.n_gates += d->gated_mul[i]; gated += d->gated_mul[i] * (2 * d->gated_l[i] + 1); }
  L.in_dim = L.n_scalars + L.n_gates + gated;
  L.out_dim = L.n_scalars + gated;
  return L;
}

// one thread per output element; BWD recomputes the activation
template <typename T, bool BWD>
__global__ void gate_kernel(const __grid_constant__ GateLayout L, const T* __restrict__ in, const T* __restrict__ gout,
                            int64_t n, T* __restrict__ out, T* __restrict__ gin) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * L.out_dim) return;
  const int64_t row = idx / L.out_dim;
  int c = (int)(idx - row * L.out_dim);
  const T* xin = in + row * L.in_dim;
  if (c < L.n_scalars) {
    int b = 0, o = c;
    while (o >= L.d.scalar_mul[b]) { o -= L.d.scalar_mul[b]; ++b; }
    T f, df;
    act_eval<T>(L.d.scalar_act[b], xin[c], &f, &df);
    const T cst = T(L.d.scalar_cst[b]);
    if (BWD) gin[row * L.in_dim + c] = gout[idx] * cst * df;
    else out[idx] = cst * f;
    return;
  }
  c -= L.n_scalars;
  int b = 0, goff = 0;  // gated block b; goff = gate index offset
  int dim = 2 * L.d.gated_l[0] + 1;
  while (c >= L.d.gated_mul[b] * dim) {
    c -= L.d.gated_mul[b] * dim;
    goff += L.d.gated_mul[b];
    ++b;
    dim = 2 * L.d.gated_l[b] + 1;
  }
  const int u = c / dim;
  const int in_col = (int)(L.n_scalars + L.n_gates + (idx - row * L.out_dim - L.n_scalars));
  const T gate_in = xin[L.n_scalars + goff + u];
  T f, df;
  act_eval<T>(L.d.gate_act[b], gate_in, &f, &df);
  const T cst = T(L.d.gate_cst[b]);
  if (!BWD) {
    out[idx] = xin[in_col] * (cst * f);
  } else {
    const T go = gout[idx];
    gin[row * L.in_dim + in_col] = go * (cst * f);
    // d/d gate = sum_m go_m * x_m * cst * f'  -> thread with m == 0 sums the 2l+1 components
    if (c - u * dim == 0) {
      T s = T(0);
      for (int m = 0; m < dim; ++m) s = fma_(gout[idx + m], xin[in_col + m], s);
      gin[row * L.in_dim + L.n_scalars + goff + u] = s * cst * df;
    }
  }
}

static int gate_check(const e3b_gate_desc* d) {
  if (!d || d->n_scalar_blocks < 0 || d->n_scalar_blocks > E3B_MAX_BLOCKS || d->n_gated_blocks < 0 ||
      d->n_gated_blocks > E3B_MAX_BLOCKS)
    return fail(E3B_ERR_INVALID, "gate: bad descriptor");
  return E3B_OK;
}

extern "C" int e3b_gate_fwd(const e3b_gate_desc* desc, int dtype, const void* in, int64_t n, void* out, void* stream) {
  int rc = gate_check(desc);
  if (rc) return rc;
  if (n == 0) return E3B_OK;
  if (!in || !out) return fail(E3B_ERR_INVALID, "gate_fwd: null argument");
  const GateLayout L = gate_layout(desc);
  if (L.out_dim == 0) return E3B_OK;
  DISPATCH_DTYPE(dtype, gate_kernel<T, false><<<blocks_for(n * L.out_dim, 256), 256, 0, (cudaStream_t)stream>>>(
                            L, (const T*)in, nullptr, n, (T*)out, nullptr);)
  return check_launch("gate_fwd");
}

extern "C" int e3b_gate_bwd(const e3b_gate_desc* desc, int dtype, const void* in, const void* gout, int64_t n,
                            void* gin, void* stream) {
  int rc = gate_check(desc);
  if (rc) return rc;
  if (n == 0) return E3B_OK;
  if (!in || !gout || !gin) return fail(E3B_ERR_INVALID, "gate_bwd: null argument");
  const GateLayout L = gate_layout(desc);
  if (L.out_dim == 0) return E3B_OK;
  DISPATCH_DTYPE(dtype, gate_kernel<T, true><<<blocks_for(n * L.out_dim, 256), 256, 0, (cudaStream_t)stream>>>(
                            L, (const T*)in, (const T*)gout, n, nullptr, (T*)gin);)
  return check_launch("gate_bwd");
}

// second derivative of the activations (force-matching training differentiates the gate backward again)
template <typename T>
__device__ __forceinline__ void act_eval2(int code, T x, T* f, T* df, T* d2f) {
  act_eval<T>(code, x, f, df);
  switch (code) {
    case 1: {  // silu: s (1 - s) (2 + x (1 - 2 s))
      const T s = T(1) / (T(1) + exp_<T>(-x));
      *d2f = s * (T(1) - s) * (T(2) + x * (T(1) - T(2) * s));
    } break;
    case 2: {  // tanh
      const T t = *f;
      *d2f = T(-2) * t * (T(1) - t * t);
    } break;
    case 3: {  // shifted softplus
      const T s = *df;
      *d2f = s * (T(1) - s);
    } break;
    case 4: {  // tanh(x) |x|
      const T t = tanh_<T>(x), ax = x < T(0) ? -x : x, sg = x > T(0) ? T(1) : (x < T(0) ? T(-1) : T(0));
      *d2f = (T(1) - t * t) * (T(2) * sg - T(2) * t * ax);
    } break;
    default: *d2f = T(0);
  }
}

// Adjoint of the gate backward: with gin = B(in, gout) and a cotangent ggin of gin, writes
//   g_in = d<ggin, B>/d in   [n, in_dim]      g_gout = d<ggin, B>/d gout   [n, out_dim]
// one thread per output element, same indexing as gate_kernel
template <typename T>
__global__ void gate_bwd2_kernel(const __grid_constant__ GateLayout L, const T* __restrict__ in,
                                 const T* __restrict__ gout, const T* __restrict__ ggin, int64_t n,
                                 T* __restrict__ g_in, T* __restrict__ g_gout) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * L.out_dim) return;
  const int64_t row = idx / L.out_dim;
  int c = (int)(idx - row * L.out_dim);
  const T* xin = in + row * L.in_dim;
  const T* hin = ggin + row * L.in_dim;
  T* oin = g_in + row * L.in_dim;
  if (c < L.n_scalars) {
    int b = 0, o = c;
    while (o >= L.d.scalar_mul[b]) { o -= L.d.scalar_mul[b]; ++b; }
    T f, df, d2f;
    act_eval2<T>(L.d.scalar_act[b], xin[c], &f, &df, &d2f);
    const T cst = T(L.d.scalar_cst[b]);
    oin[c] = hin[c] * cst * d2f * gout[idx];
    g_gout[idx] = hin[c] * cst * df;
    return;
  }
  c -= L.n_scalars;
  int b = 0, goff = 0;
  int dim = 2 * L.d.gated_l[0] + 1;
  while (c >= L.d.gated_mul[b] * dim) {
    c -= L.d.gated_mul[b] * dim;
    goff += L.d.gated_mul[b];
    ++b;
    dim = 2 * L.d.gated_l[b] + 1;
  }
  const int u = c / dim;
  const int in_col = (int)(L.n_scalars + L.n_gates + (idx - row * L.out_dim - L.n_scalars));
  const int gate_col = L.n_scalars + goff + u;
  T f, df, d2f;
  act_eval2<T>(L.d.gate_act[b], xin[gate_col], &f, &df, &d2f);
  const T cst = T(L.d.gate_cst[b]);
  const T hg = hin[gate_col];
  g_gout[idx] = hin[in_col] * (cst * f) + hg * (cst * df) * xin[in_col];
  oin[in_col] = hg * (cst * df) * gout[idx];
  if (c - u * dim == 0) {
    T s1 = T(0), s2 = T(0);
    for (int m = 0; m < dim; ++m) {
      s1 = fma_(hin[in_col + m], gout[idx + m], s1);
      s2 = fma_(xin[in_col + m], gout[idx + m], s2);
    }
    oin[gate_col] = cst * (df * s1 + hg * d2f * s2);
  }
}

extern "C" int e3b_gate_bwd2(const e3b_gate_desc* desc, int dtype, const void* in, const void* gout, const void* ggin,
                             int64_t n, void* g_in, void* g_gout, void* stream) {
  int rc = gate_check(desc);
  if (rc) return rc;
  if (n == 0) return E3B_OK;
  if (!in || !gout || !ggin || !g_in || !g_gout) return fail(E3B_ERR_INVALID, "gate_bwd2: null argument");
  const GateLayout L = gate_layout(desc);
  if (L.out_dim == 0) return E3B_OK;
  DISPATCH_DTYPE(dtype, gate_bwd2_kernel<T><<<blocks_for(n * L.out_dim, 256), 256, 0, (cudaStream_t)stream>>>(
                            L, (const T*)in, (const T*)gout, (const T*)ggin, n, (T*)g_in, (T*)g_gout);)
  return check_launch("gate_bwd2");
}

// the gate on the channel-fastest ("imu") layout: gated blocks are [m][u].  One block per node row; a
// thread takes one scalar, or one (gated block, channel u) pair and walks its 2l+1 components, so the gate
// activation and the block lookup are evaluated once per channel; all accesses are coalesced over u.
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) gate_imu_kernel(const __grid_constant__ GateLayout L, const T* __restrict__ in,
                                                       const T* __restrict__ g_mi, const T* __restrict__ g_imu, int64_t n,
                                                       T* __restrict__ out_mi, T* __restrict__ out_imu, T* __restrict__ gin) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // lets a dependent launch set itself up while this grid drains
  const int n_items = L.n_scalars + L.n_gates;
  for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
    const T* xin = in + row * L.in_dim;
    const int64_t obase = row * L.out_dim;
    T* grow = BWD ? gin + row * L.in_dim : nullptr;
    for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
      if (it < L.n_scalars) {
        int b = 0, o = it;
        while (o >= L.d.scalar_mul[b]) { o -= L.d.scalar_mul[b]; ++b; }
        T f, df;
        act_eval<T>(L.d.scalar_act[b], xin[it], &f, &df);
        const T cst = T(L.d.scalar_cst[b]);
        if (BWD) {
          const T go = (g_mi ? g_mi[obase + it] : T(0)) + (g_imu ? g_imu[obase + it] : T(0));
          grow[it] = go * cst * df;
        } else {
          if (out_mi) out_mi[obase + it] = cst * f;
          if (out_imu) out_imu[obase + it] = cst * f;
        }
        continue;
      }
      const int gi = it - L.n_scalars;               // index of the gate scalar
      int b = 0, goff = 0, boff = 0;                 // gated block b; its first gate; its offset in the gated part
      while (gi >= goff + L.d.gated_mul[b]) {
        goff += L.d.gated_mul[b];
        boff += L.d.gated_mul[b] * (2 * L.d.gated_l[b] + 1);
        ++b;
      }
      const int mul = L.d.gated_mul[b], dim = 2 * L.d.gated_l[b] + 1, u = gi - goff;
      T f, df;
      act_eval<T>(L.d.gate_act[b], xin[L.n_scalars + gi], &f, &df);
      const T cst = T(L.d.gate_cst[b]);
      const int in0 = L.n_scalars + L.n_gates + boff + u;                 // component m at + m * mul
      const int64_t o_imu = obase + L.n_scalars + boff + u;               // component m at + m * mul
      const int64_t o_mi = obase + L.n_scalars + boff + (int64_t)u * dim; // component m at + m
      if (!BWD) {
        for (int m = 0; m < dim; ++m) {
          const T v = xin[in0 + m * mul] * (cst * f);
          if (out_mi) out_mi[o_mi + m] = v;
          if (out_imu) out_imu[o_imu + (int64_t)m * mul] = v;
        }
      } else {
        T s = T(0);
        for (int m = 0; m < dim; ++m) {
          const T go = (g_mi ? g_mi[o_mi + m] : T(0)) + (g_imu ? g_imu[o_imu + (int64_t)m * mul] : T(0));
          grow[in0 + m * mul] = go * (cst * f);
          s = fma_(go, xin[in0 + m * mul], s);
        }
        grow[L.n_scalars + gi] = s * cst * df;       // d/d gate = sum_m go_m x_m cst f'
      }
    }
  }
}

extern "C" int e3b_gate_imu_fwd(const e3b_gate_desc* desc, int dtype, const void* in, int64_t n, void* out_mul_ir,
                                void* out_imu, void* stream) {
  int rc = gate_check(desc);
  if (rc) return rc;
  if (n == 0) return E3B_OK;
  if (!in || (!out_mul_ir && !out_imu)) return fail(E3B_ERR_INVALID, "gate_imu_fwd: null argument");
  const GateLayout L = gate_layout(desc);
  if (L.out_dim == 0) return E3B_OK;
  DISPATCH_DTYPE(dtype, gate_imu_kernel<T, false><<<(unsigned)(n < 148 * 64 ? n : 148 * 64), 256, 0, (cudaStream_t)stream>>>(
                            L, (const T*)in, nullptr, nullptr, n, (T*)out_mul_ir, (T*)out_imu, nullptr);)
  return check_launch("gate_imu_fwd");
}

extern "C" int e3b_gate_imu_bwd(const e3b_gate_desc* desc, int dtype, const void* in, const void* gout_mul_ir,
                                const void* gout_imu, int64_t n, void* gin, void* stream) {
  int rc = gate_check(desc);
  if (rc) return rc;
  if (n == 0) return E3B_OK;
  if (!in || !gin || (!gout_mul_ir && !gout_imu)) return fail(E3B_ERR_INVALID, "gate_imu_bwd: null argument");
  const GateLayout L = gate_layout(desc);
  if (L.out_dim == 0) return E3B_OK;
  DISPATCH_DTYPE(dtype, gate_imu_kernel<T, true><<<(unsigned)(n < 148 * 64 ? n : 148 * 64), 256, 0, (cudaStream_t)stream>>>(
                            L, (const T*)in, (const T*)gout_mul_ir, (const T*)gout_imu, n, nullptr, nullptr, (T*)gin);)
  return check_launch("gate_imu_bwd");
}

// ------------------------------------------------------------------------------------------
// mul_ir <-> imu layout conversion
struct LayoutDesc {
  int32_t n_blocks, dim;
  int32_t mul[E3B_MAX_BLOCKS], l[E3B_MAX_BLOCKS], off[E3B_MAX_BLOCKS];
};

template <typename T>
__global__ void layout_kernel(const __grid_constant__ LayoutDesc L, const T* __restrict__ in, int64_t n, int to_imu,
                              T* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * L.dim) return;
  const int64_t row = idx / L.dim;
  const int c = (int)(idx - row * L.dim);
  int b = 0;
  while (b + 1 < L.n_blocks && c >= L.off[b + 1]) ++b;
  const int d = 2 * L.l[b] + 1, r = c - L.off[b];
  // idx addresses the OUTPUT element
  int src;
  if (to_imu) { const int m = r / L.mul[b], u = r - m * L.mul[b]; src = L.off[b] + u * d + m; }
  else        { const int u = r / d, m = r - u * d;               src = L.off[b] + m * L.mul[b] + u; }
  out[idx] = in[row * L.dim + src];
}

extern "C" int e3b_layout_convert(int dtype, const void* in, int64_t n, int32_t n_blocks, const int32_t* h_mul,
                                  const int32_t* h_l, int to_imu, void* out, void* stream) {
  if (n_blocks <= 0 || n_blocks > E3B_MAX_BLOCKS || !h_mul || !h_l) return fail(E3B_ERR_INVALID, "layout_convert: bad blocks");
  if (n == 0) return E3B_OK;
  if (!in || !out) return fail(E3B_ERR_INVALID, "layout_convert: null argument");
  LayoutDesc L;
  L.n_blocks = n_blocks;
  int off = 0;
  for (int b = 0; b < n_blocks; ++b) { L.mul[b] = h_mul[b]; L.l[b] = h_l[b]; L.off[b] = off; off += h_mul[b] * (2 * h_l[b] + 1); }
  L.dim = off;
  if (off == 0) return E3B_OK;
  DISPATCH_DTYPE(dtype, layout_kernel<T><<<blocks_for(n * L.dim, 256), 256, 0, (cudaStream_t)stream>>>(
                            L, (const T*)in, n, to_imu, (T*)out);)
  return check_launch("layout_convert");
}

// ------------------------------------------------------------------------------------------
// LayerNormalization (nn/pointwise.py:32-51): per node and irreps block b (all mul (2l+1) entries)
//   y = x * rinv * std_b,   rinv = (sum x^2 / mul_b + eps)^-1/2
// One warp per (node, block); 4 warps per CTA walk the blocks of a node round-robin.
template <typename T> __device__ __forceinline__ T rsqrt_(T v);
template <> __device__ __forceinline__ float rsqrt_<float>(float v) { return 1.0f / sqrtf(v); }
template <> __device__ __forceinline__ double rsqrt_<double>(double v) { return 1.0 / sqrt(v); }

#define LN_ROWS 8   // nodes per CTA in the backward (their std partials are reduced in shared memory)

template <typename T>
__global__ void __launch_bounds__(128) layernorm_fwd_kernel(const __grid_constant__ LayoutDesc L, const T* __restrict__ x,
                                                            const T* __restrict__ stdw, T eps, int64_t n,
                                                            T* __restrict__ y, T* __restrict__ rinv) {
  const int64_t row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* xr = x + row * L.dim;
  T* yr = y + row * L.dim;
  for (int b = warp; b < L.n_blocks; b += 4) {
    const int len = L.mul[b] * (2 * L.l[b] + 1);
    T s = 0;
    for (int i = lane; i < len; i += 32) { const T v = xr[L.off[b] + i]; s = fma_(v, v, s); }
    s = warp_sum(s);
    s = __shfl_sync(0xffffffffu, s, 0);
    const T r = rsqrt_<T>(s / T(L.mul[b]) + eps);
    const T f = r * stdw[b];
    for (int i = lane; i < len; i += 32) yr[L.off[b] + i] = xr[L.off[b] + i] * f;
    if (lane == 0) rinv[row * L.n_blocks + b] = r;
  }
}

// g_x = std (gy r - x r^3 / mul <gy, x>);  g_std partial per CTA = sum over its rows of r <gy, x>
template <typename T>
__global__ void __launch_bounds__(128) layernorm_bwd_kernel(const __grid_constant__ LayoutDesc L, const T* __restrict__ x,
                                                            const T* __restrict__ gy, const T* __restrict__ rinv,
                                                            const T* __restrict__ stdw, int64_t n, T* __restrict__ gx,
                                                            T* __restrict__ gstd_partial) {
  __shared__ double part[E3B_MAX_BLOCKS];
  if (threadIdx.x < E3B_MAX_BLOCKS) part[threadIdx.x] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < LN_ROWS; ++k) {
    const int64_t row = (int64_t)blockIdx.x * LN_ROWS + k;
    if (row >= n) break;
    const T* xr = x + row * L.dim;
    const T* gr = gy + row * L.dim;
    T* or_ = gx + row * L.dim;
    for (int b = warp; b < L.n_blocks; b += 4) {
      const int len = L.mul[b] * (2 * L.l[b] + 1);
      T dot = 0;
      for (int i = lane; i < len; i += 32) dot = fma_(gr[L.off[b] + i], xr[L.off[b] + i], dot);
      dot = warp_sum(dot);
      dot = __shfl_sync(0xffffffffu, dot, 0);
      const T r = rinv[row * L.n_blocks + b], sd = stdw[b];
      const T c1 = sd * r, c2 = sd * r * r * r / T(L.mul[b]) * dot;
      for (int i = lane; i < len; i += 32) or_[L.off[b] + i] = gr[L.off[b] + i] * c1 - xr[L.off[b] + i] * c2;
      if (lane == 0) atomicAdd(&part[b], (double)(r * dot));   // shared-memory atomic, <= 8 rows x few blocks
    }
  }
  __syncthreads();
  if (gstd_partial && threadIdx.x < L.n_blocks) gstd_partial[(int64_t)blockIdx.x * L.n_blocks + threadIdx.x] = (T)part[threadIdx.x];
}

static int layernorm_desc(int32_t n_blocks, const int32_t* h_mul, const int32_t* h_l, LayoutDesc* L) {
  if (n_blocks <= 0 || n_blocks > E3B_MAX_BLOCKS || !h_mul || !h_l) return fail(E3B_ERR_INVALID, "layernorm: bad blocks");
  L->n_blocks = n_blocks;
  int off = 0;
  for (int b = 0; b < n_blocks; ++b) { L->mul[b] = h_mul[b]; L->l[b] = h_l[b]; L->off[b] = off; off += h_mul[b] * (2 * h_l[b] + 1); }
  L->dim = off;
  return E3B_OK;
}

extern "C" int64_t e3b_layernorm_bwd_blocks(int64_t n) { return (n + LN_ROWS - 1) / LN_ROWS; }

extern "C" int e3b_layernorm_fwd(int dtype, const void* x, int64_t n, int32_t n_blocks, const int32_t* h_mul,
                                 const int32_t* h_l, const void* std_w, double eps, void* y, void* rinv, void* stream) {
  LayoutDesc L;
  int rc = layernorm_desc(n_blocks, h_mul, h_l, &L);
  if (rc) return rc;
  if (n == 0 || L.dim == 0) return E3B_OK;
  if (!x || !std_w || !y || !rinv) return fail(E3B_ERR_INVALID, "layernorm_fwd: null argument");
  DISPATCH_DTYPE(dtype, layernorm_fwd_kernel<T><<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(
                            L, (const T*)x, (const T*)std_w, (T)eps, n, (T*)y, (T*)rinv);)
  return check_launch("layernorm_fwd");
}

extern "C" int e3b_layernorm_bwd(int dtype, const void* x, const void* gy, const void* rinv, int64_t n, int32_t n_blocks,
                                 const int32_t* h_mul, const int32_t* h_l, const void* std_w, void* gx,
                                 void* gstd_partial, void* stream) {
  LayoutDesc L;
  int rc = layernorm_desc(n_blocks, h_mul, h_l, &L);
  if (rc) return rc;
  if (n == 0 || L.dim == 0) return E3B_OK;
  if (!x || !gy || !rinv || !std_w || !gx) return fail(E3B_ERR_INVALID, "layernorm_bwd: null argument");
  DISPATCH_DTYPE(dtype, layernorm_bwd_kernel<T><<<(unsigned)e3b_layernorm_bwd_blocks(n), 128, 0, (cudaStream_t)stream>>>(
                            L, (const T*)x, (const T*)gy, (const T*)rinv, (const T*)std_w, n, (T*)gx, (T*)gstd_partial);)
  return check_launch("layernorm_bwd");
}

// ------------------------------------------------------------------------------------------
// Optimiser step on the flat parameter buffer: Adam (torch.optim.Adam semantics, the
// reference's optimiser: run/trainer.py:370-386, train.py:103-107) + the exponential moving average of the
// parameters the reference keeps with torch_ema (run/trainer.py, run/sde_utils.py:232-248), one pass over
// (param, grad, exp_avg, exp_avg_sq, ema): 5 reads + 4 writes per element instead of ~10 multi-tensor launches.
// grad_scale (device scalar, may be NULL): gradient-clipping coefficient; skip (device int, may be NULL): non-zero
// when the gradients hold Inf/NaN -> the step is skipped (sde_utils.py:240-246) without a host round trip.
struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, bias1, bias2_sqrt, ema_decay;
  double beta1d, beta2d;
};

// step_in / step_out (device int64, may be NULL): number of updates APPLIED so far.  torch.optim.Adam does not
// advance its step when the caller skips an update, so with a device-side skip flag the count lives on the device
// too: the bias corrections use *step_in + 1 and thread 0 writes *step_out = *step_in + (skipped ? 0 : 1) (two
// distinct locations, swapped by the caller, so no thread reads a value written by this launch).  A skipped step
// still moves the moving average (torch_ema's update follows every step of the reference loop, skipped or not).
__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v,
                                                       float* __restrict__ ema, int64_t n, AdamArgs a,
                                                       const float* __restrict__ grad_scale, const int* __restrict__ skip,
                                                       const long long* __restrict__ step_in, long long* __restrict__ step_out) {
  const bool skipped = skip && *skip;
  float bias1 = a.bias1, bias2_sqrt = a.bias2_sqrt;
  if (step_in) {
    const long long t = *step_in + 1;
    bias1 = (float)(1.0 - pow(a.beta1d, (double)t));
    bias2_sqrt = (float)sqrt(1.0 - pow(a.beta2d, (double)t));
    if (blockIdx.x == 0 && threadIdx.x == 0) *step_out = *step_in + (skipped ? 0 : 1);
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (skipped) {
    if (ema)
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        ema[i] = fmaf(a.ema_decay, ema[i] - p[i], p[i]);
    return;
  }
  const float gs = grad_scale ? *grad_scale : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float w = p[i];
    float gi = g[i] * gs;
    if (a.weight_decay != 0.f) gi = fmaf(a.weight_decay, w, gi);
    const float mi = fmaf(a.beta1, m[i], (1.f - a.beta1) * gi);
    const float vi = fmaf(a.beta2, v[i], (1.f - a.beta2) * gi * gi);
    const float denom = sqrtf(vi) / bias2_sqrt + a.eps;
    const float wn = w - (a.lr / bias1) * (mi / denom);
    m[i] = mi; v[i] = vi; p[i] = wn;
    if (ema) ema[i] = fmaf(a.ema_decay, ema[i] - wn, wn);      // decay * ema + (1 - decay) * w
  }
}

extern "C" int e3b_adam_ema_step(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, void* ema, int64_t n,
                                 float lr, double beta1, double beta2, float eps, float weight_decay, int64_t step,
                                 float ema_decay, const void* grad_scale, const void* skip, const void* step_in,
                                 void* step_out, void* stream) {
  if (n == 0) return E3B_OK;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail(E3B_ERR_INVALID, "adam_ema_step: bad argument");
  if ((step_in == nullptr) != (step_out == nullptr) || (step_in && step_in == step_out))
    return fail(E3B_ERR_INVALID, "adam_ema_step: step_in / step_out must be two distinct device scalars (or both NULL)");
  if (!step_in && step < 1) return fail(E3B_ERR_INVALID, "adam_ema_step: step >= 1 needed without a device step counter");
  AdamArgs a;
  a.lr = lr; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = eps; a.weight_decay = weight_decay; a.ema_decay = ema_decay;
  a.beta1d = beta1; a.beta2d = beta2;
  a.bias1 = (float)(1.0 - pow(beta1, (double)(step < 1 ? 1 : step)));
  a.bias2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)(step < 1 ? 1 : step)));
  const int64_t want = (n + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  adam_ema_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float*)param, (const float*)grad, (float*)exp_avg,
                                                          (float*)exp_avg_sq, (float*)ema, n, a, (const float*)grad_scale,
                                                          (const int*)skip, (const long long*)step_in, (long long*)step_out);
  return check_launch("adam_ema_step");
}

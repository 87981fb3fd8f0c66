// Shared definitions for libe3b200 (sm_100a).  Also compiled by g++ with E3B_HOST_EMU for the
// CPU emulation of the GENERATED contraction code that tests/ use to check the generator's
// algebra without a GPU (never linked into the product library).
#pragma once
#include <stdint.h>

#include <cmath>

#include "../../include/e3b200.h"

#ifdef E3B_HOST_EMU
#define __device__
#define __forceinline__ inline
#define __restrict__
template <typename T> static inline T ldg(const T* p) { return *p; }
static inline float fma_(float a, float b, float c) { return fmaf(a, b, c); }
static inline double fma_(double a, double b, double c) { return fma(a, b, c); }
#define E3B_GSH_STORE(ptr, idx, val) do { if (active) (ptr)[idx] += (val); } while (0)
#define E3B_GSH_ZERO(ptr, idx) do { } while (0)
#define E3B_GSH_STORE9(ptr, lane, v0, v1, v2, v3, v4, v5, v6, v7, v8) do { if (active) { (ptr)[0] += (v0); (ptr)[1] += (v1); \
  (ptr)[2] += (v2); (ptr)[3] += (v3); (ptr)[4] += (v4); (ptr)[5] += (v5); (ptr)[6] += (v6); (ptr)[7] += (v7); (ptr)[8] += (v8); } } while (0)
#else
#include <cuda_runtime.h>
template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#define E3B_GSH_STORE(ptr, idx, val) do { const T r_ = warp_sum(val); if (lane == 0) (ptr)[idx] = r_; } while (0)
#define E3B_GSH_ZERO(ptr, idx) do { if (lane == 0) (ptr)[idx] = T(0); } while (0)
#define E3B_GSH_STORE9(ptr, lane, ...) gsh_reduce_store9(ptr, lane, __VA_ARGS__)
#endif

#define TP_THREADS 128
#define TPP_STAGES 3     // shared-memory ring depth of the pipelined forward kernel
#define TPP_MAXSEG 256   // edges of one destination segment staged per index chunk
#define TPP_GXBUF 3      // staging rows of the backward kernel's node-reduction mode (TMA reduce-add in flight)
#define TPP2_MAXSEG 128  // the same for the paired kernels, whose ring depth is a launch parameter (TpArgs::n_stages)
#define TPP2_LAG 1      // the producer refills the stage of the edge this many behind the one its own warp consumes

#if defined(__CUDACC__) && !defined(E3B_HOST_EMU)
// ---- mbarrier + TMA bulk-copy primitives (sm_90+; SASS: SYNCS / UBLKCP) ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one edge -> one ring stage: [weight row | source feature row]
__device__ __forceinline__ void tpp_issue(float* stage, uint64_t* bar, const float* w_row, int row_w,
                                          const float* x_row, int row_x) {
  mbar_expect_tx(bar, (uint32_t)(row_w + row_x) * 4u);
  bulk_g2s(stage, w_row, (uint32_t)row_w * 4u, bar);
  bulk_g2s(stage + row_w, x_row, (uint32_t)row_x * 4u, bar);
}
// ---- TMA reduce-add of a shared-memory row into global memory (sm_90+; SASS: UBLKRED), bulk async groups
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
bool e3b_tp_pipelined_enabled();
bool e3b_tp_paired_enabled(bool preferred);
bool e3b_tp_paired_fwd_enabled(bool preferred);
bool e3b_tp_decoupled_enabled();
int e3b_tp_stages(int bwd, size_t per_stage_bytes, size_t fixed_bytes);
// ---- two channels per thread: packed fp32 pairs (sm_100 FFMA2 / FMUL2; a scalar or an immediate broadcasts to both halves,
// so Clebsch-Gordan literals and the spherical harmonics of the edge need no register pairs)
struct __align__(8) F2 {
  float2 v;
  __device__ __forceinline__ F2() {}
  __device__ __forceinline__ explicit F2(float s) : v(make_float2(s, s)) {}
  __device__ __forceinline__ explicit F2(float2 t) : v(t) {}
};
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return F2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 fma_(F2 a, F2 b, F2 c) { return F2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ void red_add(F2* p, F2 v) {   // fire-and-forget L2 reduction of a channel pair (sm_90+)
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.v.x), "f"(v.v.y) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float pair_sum(F2 v) { return v.v.x + v.v.y; }
__device__ __forceinline__ float pair_sum(float v) { return v; }
__device__ __forceinline__ F2 ldg(const F2* p) { return F2(__ldg(reinterpret_cast<const float2*>(p))); }
#define E3B_GSH_STORE2(ptr, idx, val) do { const float r_ = warp_sum((val).v.x + (val).v.y); if (lane == 0) (ptr)[idx] = r_; } while (0)
// sums over the 32 lanes of NINE values per lane (the d/dY partials of one edge, sh_dim = 9) and their store: a halving butterfly
// (after the exchange over lane bit b a lane keeps half of its values) takes 14 shuffles instead of 45; lane 4 i ends up with the
// total of value i (i < 8), the ninth value goes through the plain butterfly.  NOT inlined: the unrolled edge loops of the
// warps of a kernel compete for the 32 KB instruction cache of the SM, one copy serves them all.
static __device__ __noinline__ void gsh_reduce_store9(float* __restrict__ row, int lane, float r0, float r1, float r2, float r3, float r4,
                                               float r5, float r6, float r7, float r8) {
  const float r[9] = {r0, r1, r2, r3, r4, r5, r6, r7, r8};
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float t[4], s2[2];
#pragma unroll
  for (int k = 0; k < 4; ++k) t[k] = (b4 ? r[k + 4] : r[k]) + __shfl_xor_sync(0xffffffffu, b4 ? r[k] : r[k + 4], 16);
#pragma unroll
  for (int k = 0; k < 2; ++k) s2[k] = (b3 ? t[k + 2] : t[k]) + __shfl_xor_sync(0xffffffffu, b3 ? t[k] : t[k + 2], 8);
  float q = (b2 ? s2[1] : s2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? s2[0] : s2[1], 4);
  q += __shfl_xor_sync(0xffffffffu, q, 2);
  q += __shfl_xor_sync(0xffffffffu, q, 1);
  const float s8 = warp_sum(r[8]);
  if ((lane & 3) == 0) row[(lane >> 4) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = q;
  if (lane == 1) row[8] = s8;
}
#define E3B_GSH_ZERO2(ptr, idx) do { if (lane == 0) (ptr)[idx] = 0.f; } while (0)
#endif

// Arguments of the tensor-product convolution kernels.  Dims are in scalars per row.
template <typename T>
struct TpArgs {
  const T* x;    // [N, x_dim]   imu layout
  const T* sh;   // [E, sh_dim]
  const T* w;    // [E, w_dim]
  const T* gy;   // [N, y_dim]   (backward)
  T* y;          // [N, y_dim]
  T* gx_edge;    // [E, x_dim]   (backward, may be null)
  T* gx_node;    // [N, x_dim]   (backward, pipelined kernels only, may be null): d/dx reduced into the SOURCE node's row
  T* gsh;        // [E, n_part, sh_dim] (backward, may be null)
  T* gw;         // [E, w_dim]   (backward)
  const int64_t* in_ptr;
  const int32_t* in_nbr;
  const int32_t* in_eid;
  const int32_t* w_idx;   // [E] or null: row of w (forward and backward read) of edge id e -- edges that share a weight row
                          // (the two directions of an undirected edge); pipelined kernels only
  int64_t n_nodes;
  int64_t x_dim, sh_dim, w_dim, y_dim;
  int32_t mul, n_chunks, n_part;
  int32_t n_stages;       // paired pipelined kernels: depth of the shared-memory ring
};

// fp32-faithful dense contractions on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   C[r, n] = epilogue( sum_k A[r, k] * B[n, k] )          A, C fp32 in HBM, B a (small) weight
//
// 3xTF32: every fp32 operand is split into a TF32 "hi" part and a TF32 "lo" part (the rounded
// remainder) and three tcgen05.mma.kind::tf32 products (lo*hi + hi*lo + hi*hi) accumulate in
// TMEM in fp32 -- relative error ~1e-6, which the 1e-5 parity budget of the interaction block
// needs (plain TF32 gives ~1e-3).
//
// B is a weight: it is split and laid out ONCE by e3b_gemm_pack into the UMMA canonical K-major /
// no-swizzle layout, tile by tile, so that the kernel fetches a whole (hi | lo) B chunk with one
// TMA bulk copy.  A is an activation: converter warps stream it with cp.async (16 B per thread,
// coalesced, affine row addressing so the irreps layouts are read in place) into a padded raw
// ring, read their own accumulator row back, split it in registers and store the hi / lo parts
// straight into TENSOR MEMORY (tcgen05.st); the MMAs take A from TMEM, so of the operands only B
// crosses shared memory (the kernel was bound by shared-memory bandwidth with A staged there).
//
// One persistent CTA = 19 warps, warp-specialised:
//   warps 0-7  epilogue   TMEM -> registers (tcgen05.ld) -> epilogue math -> global
//   warps 8-15 converter  global -(cp.async)-> raw ring -> split hi/lo -> A ring in TMEM
//   warp  16   producer   TMA bulk copies of packed B chunks into the B ring
//   warps 17-18 MMA       one elected thread each issues tcgen05.mma and commits to the mbarriers (long K: the two
//                         alternate accumulation chains, halving the per-chunk issue overhead; else warp 17 alone)
// Three pipelines (mbarrier full/empty pairs): A ring, B ring, two TMEM accumulators.  When the
// whole K extent of a row tile fits the A ring it stays RESIDENT across the column tiles.
// The tensor core accumulates with truncation, so an unbroken chain over a long K drifts
// (1.5e-5 at K = 1920, measured): chains are cut every 64 floats of K and the partial sums are
// added in fp32 registers by the epilogue warps while the next chain runs in the other buffer.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/e3b200.h"

int e3b_fail(int code, const char* fmt, ...);  // e3b200.cu

namespace {

constexpr int BM = 128;     // UMMA M
constexpr int BK = 32;      // floats per K chunk = 4 MMA K-steps of 8 = 8 sixteen-byte columns
constexpr int KSEG = 2;     // chunks per accumulation chain (64 floats of K)
constexpr int A_COLS = 2 * BK;   // TMEM columns of one A stage: 32 hi | 32 lo
constexpr int NTHREADS = 608;
constexpr int NCVT = 256;    // converter threads (warps 4-11)
constexpr int CVT0 = 256;   // first converter thread
constexpr int NEPI = 256;   // epilogue threads (warps 0-7)
constexpr int MAXG = E3B_GEMM_MAX_GROUP;

struct Problem {
  e3b_gemm_problem p;
  int32_t m_tiles, n_tiles, k_chunks;
  int32_t n_ctas;        // CTAs of this problem; CTA i walks the tiles [i * T / n, (i + 1) * T / n), T = m_tiles * n_tiles,
                         // in row-major order (tile t = row tile t / n_tiles, column tile t % n_tiles)
  int32_t cta_begin;     // first CTA (blockIdx.x) of this problem
  uint64_t a_mul;        // ceil(2^40 / a_d): r / a_d == (r * a_mul) >> 40 for r < 2^31, a_d < 512
  uint64_t c_mul;        // the same for c_d
  int32_t tma;           // the A tile is fetched by TMA tensor copies (plain row-major A streamed along a long K): one
                         // instruction per 128 x 32 chunk into a 128-byte-swizzled slot instead of 2048 16-byte copies
                         // issued by the converter threads, and up to RS chunks in flight per SM
  int32_t tma_c;         // C is a plain row-major matrix written (not accumulated) by epilogue 0 / 2: every epilogue warp
                         // stages its 32 x 32 block in a swizzled tile and ONE TMA tensor store writes it (the 32 x
                         // 16-byte stores per lane it replaces bound the [E,64] x [64,1920] launch: 310 us, 162 without)
  int32_t dbg;           // E3B_GEMM_DEBUG bisection bits: 1 skip A load+convert, 2 skip MMA, 4 skip B loads, 8 skip stores
};
struct Batch {
  Problem pr[MAXG];
  int32_t n;
};
struct TmapBatch {
  alignas(64) CUtensorMap a[MAXG];   // A of problem i as a 2-D tensor {K, M}, box {32, 128}, 128-byte swizzle (when pr[i].tma)
  alignas(64) CUtensorMap c[MAXG];   // C of problem i as {N, M}, box {32, 32}, 128-byte swizzle (when pr[i].tma_c)
};
constexpr int RS = 6;                // TMA mode: 16 KB slots of the raw ring
constexpr uint32_t RAW_SLOT_BYTES = BM * BK * 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B (128 B
// contiguous); LBO = bytes between the two 16-byte K columns of one MMA, SBO = bytes between
// consecutive 8-row groups; both encoded >> 4; bits 46-47 = 1 (Blackwell descriptor version).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// A operand from tensor memory (lane = row, one tf32 per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
        "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
        "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
        "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int32_t c0, int32_t c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_addr(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// fp32 -> (tf32 hi, tf32 lo) with round-to-nearest (cvt.rna): hi has its low 13 bits cleared, so
// the tensor core sees it exactly whether it truncates or rounds; lo = rna(x - hi) (x - hi is
// exact in fp32).  Rounding (not masking) keeps the residual unbiased.
__host__ __device__ __forceinline__ float rna_tf32(float x) {
#ifdef __CUDA_ARCH__
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
#else
  return x;
#endif
}
__device__ __forceinline__ void split4(const float4 x, float4* hi, float4* lo) {
  hi->x = rna_tf32(x.x); hi->y = rna_tf32(x.y); hi->z = rna_tf32(x.z); hi->w = rna_tf32(x.w);
  lo->x = rna_tf32(x.x - hi->x); lo->y = rna_tf32(x.y - hi->y);
  lo->z = rna_tf32(x.z - hi->z); lo->w = rna_tf32(x.w - hi->w);
}
// The hot-loop version (converter warps): round-to-nearest by integer arithmetic on the bit pattern
// (add half an ulp of tf32, clear the 13 low bits; ties away from zero, no Inf/NaN special cases) = 2
// instructions instead of cvt.rna's 4; lo = x - hi is exact and is left unrounded: whatever the tensor
// core does with its low 13 bits is an error of 2^-22 |x|, below the dropped lo*lo term.
__device__ __forceinline__ float rn_tf32_fast(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split4_fast(const float4 x, float4* hi, float4* lo) {
  hi->x = rn_tf32_fast(x.x); hi->y = rn_tf32_fast(x.y); hi->z = rn_tf32_fast(x.z); hi->w = rn_tf32_fast(x.w);
  lo->x = x.x - hi->x; lo->y = x.y - hi->y; lo->z = x.z - hi->z; lo->w = x.w - hi->w;
}

// ShiftedSoftPlus pieces (e3nn nn.FullyConnectedNet activation): ssp(z) = softplus(z) - ln 2.
// Only four epilogue warps per SM evaluate these, so they use the SFU approximations (ex2 / lg2):
// absolute error <= ~2e-7 on values of order one, inside the 1e-5 budget of the block.
__device__ __forceinline__ float ssp_f(float z) {
  return (z > 15.f ? z : __logf(1.f + __expf(z))) - 0.6931471805599453f;
}
// d/dz [cst * ssp(z)] expressed through the stored output h = cst * ssp(z):
// sigmoid(z) = 1 - exp(-softplus(z)) = 1 - 0.5 * exp(-h / cst)
__device__ __forceinline__ float dssp_from_out(float h, float cst, float inv_cst) {
  return cst * (1.f - 0.5f * __expf(-h * inv_cst));
}

template <int BN, int PRAW, int SB>
struct Smem {
  static constexpr int B_STAGE = 2 * BN * BK;      // floats: hi | lo
  static constexpr int RAW_ROW = BK + 4;           // padded row (144 B): conflict-free row-per-lane reads
  static constexpr int RAW_STAGE = BM * RAW_ROW;
  static constexpr int EPI_STAGE = 32 * 40;        // per epilogue warp: 32 rows x (32 + 4 pad) floats, or a swizzled
                                                   // 32 x 32 tile for a TMA store (stride kept a multiple of 1024 B)
  static constexpr size_t BYTES = (size_t)(SB * B_STAGE + PRAW * RAW_STAGE + 8 * EPI_STAGE) * 4 + 1024 /*align slack*/;
};

struct EpiCtx {
  const Problem* P;
  const CUtensorMap* tmap_c;
  uint64_t* acc_full;
  uint64_t* acc_empty;
  float* stg;            // warp-private staging tile [32][36]
  uint32_t tmem_base;
  int t0, t1, k_chunks;
};

template <int EPI> __device__ __forceinline__ float epi_apply(float o, float h, float cst) {
  if (EPI == 2) return cst * ssp_f(o);
  if (EPI == 3) return o * dssp_from_out(h, cst, 1.f / cst);
  return o;
}

// Epilogue warps (8): warp w drains TMEM lanes 32 (w % 4) .. +31 (thread = one accumulator row) of the column
// half w / 4 of every tile.  DENSE outputs (unit column stride, N % 4 == 0) are transposed through a
// warp-private staging tile so that every store instruction writes whole 128-byte lines (4 rows x 128 B per
// warp instruction) instead of 32 partial sectors.
template <int BN, bool MULTI, int NACC, int EPI, bool DENSE>
__device__ __forceinline__ void epilogue_role(const EpiCtx& c) {
  constexpr int HB = BN / 2;                     // columns of a tile handled by this warp
  const Problem& P = *c.P;
  const e3b_gemm_problem& g = P.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cb0 = (warp >> 2) * HB;
  const int n_seg = (c.k_chunks + KSEG - 1) / KSEG;
  uint32_t acc_it = 0;
  const uint32_t t_lane = c.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb0;
  const int t_row = lane >> 3, t_c4 = (lane & 7) * 4;     // transposed role: rows t_row + 4 i, columns t_c4..+3
  const float alpha = g.alpha, cst = g.act_cst;
  const bool accumulate = g.accumulate != 0;
  int m_cur = -1;
  bool row_ok = false;
  float* c_row = nullptr;
  const float* h_row = nullptr;
  float aux[EPI == 1 ? 32 : 1];
  int r_base = 0;
  for (int t = c.t0; t < c.t1; ++t) {
    const int m = t / P.n_tiles, n = t - m * P.n_tiles;
    if (m != m_cur) {
      m_cur = m;
      const int row = m * BM + q * 32 + lane;
      row_ok = row < g.M;
      if (row_ok && !DENSE) {
        int cq = row / g.c_d;
        if (g.row_map) { const int mq = __ldg(g.row_map + cq); row_ok = mq >= 0; cq = row_ok ? mq : 0; }
        c_row = g.C + (int64_t)cq * g.c_s1 + (int64_t)(row % g.c_d) * g.c_s2;
        if (EPI == 3) h_row = g.H + (int64_t)row * g.h_ld;
      }
      if (EPI == 1) {
#pragma unroll
        for (int v = 0; v < (EPI == 1 ? 32 : 1); ++v)
          aux[v] = (row_ok && v < (g.aux_cols > 0 ? g.aux_cols : g.V)) ? __ldg(g.aux + (int64_t)(row / g.aux_d) * g.aux_ld + v) : 0.f;
      }
      // DENSE: rows of this warp's quarter tile handled by this lane after the transposition
      r_base = m * BM + q * 32 + t_row;
    }
    const int n0 = n * BN + cb0;                  // first column of this warp's half tile
    float racc[MULTI ? HB : 1];
    if (MULTI) {
#pragma unroll
      for (int i = 0; i < (MULTI ? HB : 1); ++i) racc[i] = 0.f;
    }
    float red[EPI == 1 ? HB / 16 : 1];            // epilogue 1: the reduced outputs of this half tile
    for (int seg = 0; seg < (MULTI ? n_seg : 1); ++seg, ++acc_it) {
      const uint32_t buf = acc_it % NACC;
      mbar_wait(&c.acc_full[buf], (acc_it / NACC) & 1u);
      tc_fence_after();
      const bool last = !MULTI || seg == n_seg - 1;
#pragma unroll
      for (int cb = 0; cb < HB; cb += 32) {
        float v[32];
        tmem_ld32(t_lane + buf * BN + (uint32_t)cb, v);
        if (MULTI) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { racc[(MULTI ? cb : 0) + (MULTI ? i : 0)] += v[i]; v[i] = racc[(MULTI ? cb : 0) + (MULTI ? i : 0)]; }
        }
        const int nb = n0 + cb;
        if (!last || nb >= g.N || (P.dbg & 8)) continue;
        if (DENSE && EPI != 3 && P.tma_c) {
          // thread = row: its 32 values go into the swizzled staging tile, one tensor store writes the block
          // (rows / columns beyond M / N are clipped by the tensor map)
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous block has left
          __syncwarp();
          const uint32_t srow = smem_addr(c.stg) + (uint32_t)lane * 128u;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            sts128(srow + (((uint32_t)i ^ (uint32_t)(lane & 7)) << 4),
                   make_float4(epi_apply<EPI>(alpha * v[4 * i], 0.f, cst), epi_apply<EPI>(alpha * v[4 * i + 1], 0.f, cst),
                               epi_apply<EPI>(alpha * v[4 * i + 2], 0.f, cst), epi_apply<EPI>(alpha * v[4 * i + 3], 0.f, cst)));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) tma_store_2d(c.tmap_c, smem_addr(c.stg), nb, m * BM + q * 32);
        } else if (DENSE) {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(c.stg + lane * 36 + 4 * i) =
                make_float4(alpha * v[4 * i], alpha * v[4 * i + 1], alpha * v[4 * i + 2], alpha * v[4 * i + 3]);
          __syncwarp();
          const int col = nb + t_c4;
          if (col < g.N) {
            const float* hp = EPI == 3 ? g.H + (int64_t)r_base * g.h_ld + col : nullptr;
            float4 hv[EPI == 3 ? 8 : 1];
            if (EPI == 3) {                       // all eight loads in flight before the first use
#pragma unroll
              for (int i = 0; i < (EPI == 3 ? 8 : 1); ++i)
                hv[i] = r_base + 4 * i < g.M ? __ldg(reinterpret_cast<const float4*>(hp + (int64_t)(4 * i) * g.h_ld))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (r_base + 4 * i < g.M) {
                float4 o = *reinterpret_cast<const float4*>(c.stg + (t_row + 4 * i) * 36 + t_c4);
                const float4 h = hv[EPI == 3 ? i : 0];
                o.x = epi_apply<EPI>(o.x, h.x, cst); o.y = epi_apply<EPI>(o.y, h.y, cst);
                o.z = epi_apply<EPI>(o.z, h.z, cst); o.w = epi_apply<EPI>(o.w, h.w, cst);
                const int r = r_base + 4 * i;            // affine row addressing (irreps blocks written in place)
                const int rq = (int)(((uint64_t)(uint32_t)r * P.c_mul) >> 40);
                int aq = rq;                                // grouped rows: virtual -> actual row group, < 0 = padding
                if (g.row_map) aq = __ldg(g.row_map + rq);
                if (aq >= 0) {
                  float4* d4 = reinterpret_cast<float4*>(g.C + (int64_t)aq * g.c_s1 + (int64_t)(r - rq * g.c_d) * g.c_s2 + col);
                  if (accumulate) { const float4 old = *d4; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                  *d4 = o;
                }
              }
            }
          }
        } else if (EPI == 1) {
          // weighted reduction over groups of V accumulator columns (self-connection)
          if (g.V == 16) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int t2 = 0; t2 < 16; ++t2) { s0 = fmaf(aux[EPI == 1 ? t2 : 0], v[t2], s0); s1 = fmaf(aux[EPI == 1 ? t2 : 0], v[16 + t2], s1); }
            red[EPI == 1 ? cb / 16 : 0] = alpha * s0;
            red[EPI == 1 ? cb / 16 + 1 : 0] = alpha * s1;
          } else {
            float s0 = 0.f;
#pragma unroll
            for (int t2 = 0; t2 < 32; ++t2) s0 = fmaf(aux[EPI == 1 ? t2 : 0], v[t2], s0);
            red[EPI == 1 ? cb / 32 : 0] = alpha * s0;
          }
        } else {
          // generic (strided) output: one scalar store per element
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < g.N) {
                const float o = epi_apply<EPI>(alpha * v[i], EPI == 3 ? __ldg(h_row + nb + i) : 0.f, cst);
                float* cp = c_row + (int64_t)(nb + i) * g.c_s3;
                *cp = o + (accumulate ? *cp : 0.f);
              }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&c.acc_empty[buf]);
      if (EPI == 1 && last && row_ok && n0 < g.N && !(P.dbg & 8)) {
        // outputs of this half tile: columns n0 / V .. of the row; vector stores when the row is contiguous
        const int per = HB / g.V, oc0 = n0 / g.V, n_out = g.N / g.V;
        float* c0 = c_row + (int64_t)oc0 * g.c_s3;
        const bool vec = g.c_s3 == 1 && oc0 + per <= n_out;
        if (vec && per == 4 && (reinterpret_cast<uintptr_t>(c0) & 15) == 0) {
          float4 o = make_float4(red[0], red[EPI == 1 && HB >= 32 ? 1 : 0], red[EPI == 1 && HB >= 64 ? 2 : 0], red[EPI == 1 && HB >= 64 ? 3 : 0]);
          float4* d4 = reinterpret_cast<float4*>(c0);
          if (accumulate) { const float4 old = *d4; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
          *d4 = o;
        } else if (vec && per == 2 && (reinterpret_cast<uintptr_t>(c0) & 7) == 0) {
          float2 o = make_float2(red[0], red[EPI == 1 && HB >= 32 ? 1 : 0]);
          float2* d2 = reinterpret_cast<float2*>(c0);
          if (accumulate) { const float2 old = *d2; o.x += old.x; o.y += old.y; }
          *d2 = o;
        } else {
#pragma unroll
          for (int i = 0; i < (EPI == 1 ? HB / 16 : 1); ++i)
            if (i < per && oc0 + i < n_out) {
              float* cp = c0 + (int64_t)i * g.c_s3;
              *cp = red[EPI == 1 ? i : 0] + (accumulate ? *cp : 0.f);
            }
        }
      }
    }
  }
  // tensor stores of this warp must have completed before the CTA's shared memory goes away
  if (DENSE && EPI != 3 && P.tma_c && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// every role repeats the same walk over this CTA's tile range [t0, t1)
// TMEM map (512 columns): [NACC accumulators of BN columns][SA stages of A: 32 hi | 32 lo columns]
template <int BN, bool MULTI, int NACC, int SA, int PRAW, int SB>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tf32x3_kernel(const __grid_constant__ Batch batch,
                                                                  const __grid_constant__ TmapBatch tmaps) {
  using L = Smem<BN, PRAW, SB>;
  static_assert(NACC * BN + SA * A_COLS <= 512 && PRAW >= 4 && PRAW % 2 == 0, "TMEM / ring configuration");
  static_assert(!MULTI || ((size_t)PRAW * L::RAW_STAGE * 4 >= (size_t)RS * RAW_SLOT_BYTES && (SB * L::B_STAGE * 4) % 1024 == 0),
                "TMA mode: RS swizzled 16 KB slots, 1024-byte aligned, inside the raw ring");
  extern __shared__ unsigned char smem_dyn[];
  float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  float* sB = smem;                                   // [SB][hi|lo][c(8)][row(BN)][4]
  float* sRaw = sB + SB * L::B_STAGE;                 // [PRAW][row(128)][36]
  float* sEpi = sRaw + PRAW * L::RAW_STAGE;           // [8 warps][32 rows][36]
  __shared__ uint64_t a_full[SA], a_empty[SA], b_full[SB], b_empty[SB], acc_full[NACC], acc_empty[NACC];
  __shared__ uint64_t raw_full[RS], raw_empty[RS];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // which problem of the group does this CTA work on
  int gi = 0;
#pragma unroll 1
  for (int i = 1; i < batch.n; ++i)
    if ((int)blockIdx.x >= batch.pr[i].cta_begin) gi = i;
  const Problem& P = batch.pr[gi];
  const e3b_gemm_problem& g = P.p;
  const int local = (int)blockIdx.x - P.cta_begin;
  const int k_chunks = P.k_chunks;
  const bool resident = k_chunks <= SA;
  const bool tma = MULTI && P.tma != 0;               // host: only with a streaming A ring (!resident)
  // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as every CTA of this one has
  // passed this point (it waits for our completion before touching memory), and this kernel's own set-up below -- tensor
  // memory allocation, barrier initialisation -- runs while the previous kernel drains; nothing a predecessor wrote is
  // read before griddepcontrol.wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 17) {  // TMEM allocation is warp-collective; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {   // arrival counts: one per converter warp of the filling group / per epilogue warp
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], NCVT / 2 / 32); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < NACC; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], NEPI / 32); }
    for (int s = 0; s < RS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], NCVT / 2 / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_a0 = tmem_base + NACC * BN;     // first A stage
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything launched before this kernel has completed
  int m_tiles = P.m_tiles;
  if (g.n_blocks) {                                   // grouped rows: only the blocks of 128 row groups in use (a_d tiles each)
    const int used = __ldg(g.n_blocks) * g.a_d;
    m_tiles = used < m_tiles ? used : m_tiles;
  }
  const int n_all = m_tiles * P.n_tiles;
  const int t0 = (int)((int64_t)local * n_all / P.n_ctas), t1 = (int)((int64_t)(local + 1) * n_all / P.n_ctas);

  if (t1 > t0) {
    if (warp >= 8 && warp < 16) {
      // =============================== A converter ===============================
      // Two groups of 4 warps take alternate jobs (K chunks), so two chunks are in flight through the
      // load -> split -> tcgen05.st chain at any time.  Within a group:
      // load role : thread -> one 16-byte column `col` of the chunk and the 8 rows row0 + 16 i (a warp
      //             instruction reads 64 contiguous bytes of 8 rows: whole sectors);
      // split role: thread -> ITS accumulator row (TMEM lane 32 * (warp % 4) + lane), all 32 columns;
      //             the padded raw rows make the row-per-lane reads conflict-free.
      constexpr int NG = 2;
      const int ct = tid - CVT0;                 // 0..255
      const int grp = ct >> 7, w = (ct >> 5) & 3;
      const int col = (w & 1) * 4 + (lane >> 3);
      const int row0 = 8 * (w >> 1) + (lane & 7);
      const int my_row = 32 * w + lane;
      const uint32_t t_mine = tmem_a0 + ((uint32_t)(32 * w) << 16);
      if (tma) {
        // TMA mode: the producer warp fetches chunk j into slot j % RS (rows of 128 B, 16-byte pieces XOR-swizzled
        // with the row index mod 8, so the row-per-lane reads below are conflict-free without padding)
        const int total = (t1 - t0) * k_chunks;
        const uint32_t row_addr = smem_addr(sRaw) + (uint32_t)my_row * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        const bool skip_a = (P.dbg & 1) != 0;
#pragma unroll 1
        for (int j = grp; j < total; j += NG) {
          const int st = j % SA, rs = j % RS;
          mbar_wait(&raw_full[rs], ((uint32_t)j / RS) & 1u);
          float hi[32], lo[32];
          if (!skip_a) {
            const uint32_t base = row_addr + (uint32_t)rs * RAW_SLOT_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 h, l;
              split4_fast(lds128(base + (((uint32_t)i ^ sw) << 4)), &h, &l);
              hi[4 * i] = h.x; hi[4 * i + 1] = h.y; hi[4 * i + 2] = h.z; hi[4 * i + 3] = h.w;
              lo[4 * i] = l.x; lo[4 * i + 1] = l.y; lo[4 * i + 2] = l.z; lo[4 * i + 3] = l.w;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&raw_empty[rs]);       // the slot is in registers: the producer may refill it
          mbar_wait(&a_empty[st], (((uint32_t)j / SA) & 1u) ^ 1u);
          if (!skip_a) {
            tc_fence_after();
            const uint32_t ta = t_mine + (uint32_t)st * A_COLS;
            tmem_st32(ta, hi);
            tmem_st32(ta + BK, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[st]);
        }
      } else {
      float* raw_ring = sRaw + (size_t)grp * (PRAW / NG) * L::RAW_STAGE;
      constexpr int PR = PRAW / NG;              // raw slots of one group
      static_assert(PR >= 2, "raw ring");
      // jobs: one per K chunk of every tile (streaming) or of every distinct row tile (resident A)
      const int total = (resident ? (t1 - 1) / P.n_tiles - t0 / P.n_tiles + 1 : t1 - t0) * k_chunks;
      // issue stream state (runs PR - 1 of this group's jobs ahead of the conversion).  Kept lean: this is the
      // converter's own instruction stream (ncu: 180 of its 376 instructions per job went into issuing 8 copies when the
      // tile coordinates were re-derived by division and the addresses rebuilt from 64-bit offsets per job).
      const float* rp[8];                        // row pointers of the current row tile (this thread's 8 rows)
      uint32_t rmask = 0;                        // bit i: row i exists (inside M, not a padding slot)
      const int Kdim = g.K, Mdim = g.M, n_tiles = P.n_tiles;
      const float* const Abase = g.A;
      const bool skip_a = (P.dbg & 1) != 0;
      auto load_rows = [&](int m_tile) {
        rmask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = m_tile * BM + row0 + 16 * i;
          const int q = (int)(((uint64_t)(uint32_t)r * P.a_mul) >> 40);
          int aq = q;                                     // grouped rows: virtual -> actual row group, < 0 = padding
          if (g.row_map) aq = r < Mdim ? __ldg(g.row_map + q) : -1;
          const bool ok = r < Mdim && aq >= 0;
          rp[i] = Abase + (ok ? (int64_t)aq * g.a_s1 + (int64_t)(r - q * g.a_d) * g.a_s2 : 0) + col * 4;
          rmask |= ok ? (1u << i) : 0u;
        }
      };
      // coordinates of the next global job to issue: row tile i_m, column tile i_n, chunk i_kc
      int i_m = t0 / n_tiles, i_n = t0 - i_m * n_tiles, i_kc = 0, i_slot = 0, i_job = 0;
      int m_loaded = -1;
      auto advance = [&]() {                              // to the next global job
        ++i_job;
        if (++i_kc == k_chunks) {
          i_kc = 0;
          if (resident) ++i_m;                            // one block of jobs per distinct row tile
          else if (++i_n == n_tiles) { i_n = 0; ++i_m; }
        }
      };
      for (int s = 0; s < grp; ++s) advance();            // first job of this group
      const uint32_t dst0 = smem_addr(raw_ring + row0 * L::RAW_ROW + col * 4);
      auto issue = [&]() {
        if (i_job < total && !skip_a) {
          if (i_m != m_loaded) { load_rows(i_m); m_loaded = i_m; }
          const int k = i_kc * BK;
          const uint32_t dst = dst0 + (uint32_t)i_slot * (uint32_t)(L::RAW_STAGE * 4);
          if (rmask == 0xffu && k + col * 4 < Kdim) {     // the common case: whole rows, inside K
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(i * 16 * L::RAW_ROW * 4)), "l"(rp[i] + k) : "memory");
          } else {                                        // 16-byte asynchronous copies; src bytes = 0 zero-fills
            const bool kok = k + col * 4 < Kdim;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(i * 16 * L::RAW_ROW * 4)),
                           "l"(kok ? rp[i] + k : rp[i] - col * 4), "r"((kok && ((rmask >> i) & 1u)) ? 16u : 0u) : "memory");
          }
        }
#pragma unroll 1
        for (int s = 0; s < NG; ++s) advance();
        if (++i_slot == PR) i_slot = 0;
        cp_async_commit();
      };
#pragma unroll 1
      for (int j = 0; j < PR - 1; ++j) issue();
      int slot = 0;
#pragma unroll 1
      for (int j = grp; j < total; j += NG) {
        const int st = j % SA;
        const uint32_t par = (((uint32_t)j / SA) & 1u) ^ 1u;      // first pass through the ring is free
        cp_async_wait<PR - 2>();                 // this thread's copies of job j have landed ...
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // ... and the group's; the group is done
        else asm volatile("bar.sync 2, 128;" ::: "memory");            // reading its previous job
        issue();                                 // the group's job PR - 1 ahead reuses the slot just released
        mbar_wait(&a_empty[st], par);
        if (!skip_a) {
          tc_fence_after();
          const float4* raw = reinterpret_cast<const float4*>(raw_ring + (size_t)slot * L::RAW_STAGE + my_row * L::RAW_ROW);
          float hi[32], lo[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 h, l;
            split4_fast(raw[i], &h, &l);
            hi[4 * i] = h.x; hi[4 * i + 1] = h.y; hi[4 * i + 2] = h.z; hi[4 * i + 3] = h.w;
            lo[4 * i] = l.x; lo[4 * i + 1] = l.y; lo[4 * i + 2] = l.z; lo[4 * i + 3] = l.w;
          }
          const uint32_t ta = t_mine + (uint32_t)st * A_COLS;
          tmem_st32(ta, hi);
          tmem_st32(ta + BK, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[st]);  // 128 arrivals per stage cost more than the conversion itself
        if (++slot == PR) slot = 0;
      }
      cp_async_wait<0>();
      }
    } else if (warp == 16) {
      // =============================== B producer (TMA) ===============================
      if (lane == 0) {
        uint32_t it = 0;
        constexpr uint32_t bytes = (uint32_t)L::B_STAGE * 4u;
        const float* b_set = g.B_packed;
        int m_sel = -1;
        for (int t = t0; t < t1; ++t) {
          const int m = t / P.n_tiles, n = t - m * P.n_tiles;
          if (g.b_sel && m != m_sel) {                    // a row tile never straddles a block of 128 row groups
            m_sel = m;
            const int q0 = (int)(((uint64_t)(uint32_t)(m * BM) * P.a_mul) >> 40);
            b_set = g.B_packed + (int64_t)__ldg(g.b_sel + (q0 >> 7)) * g.b_set_stride;
          }
            for (int kc = 0; kc < k_chunks; ++kc, ++it) {
              if (tma) {                                    // the A chunk of this job
                const uint32_t rs = it % RS;
                mbar_wait(&raw_empty[rs], ((it / RS) & 1u) ^ 1u);
                if (P.dbg & 1) mbar_arrive(&raw_full[rs]);
                else {
                  mbar_expect_tx(&raw_full[rs], RAW_SLOT_BYTES);
                  tma_load_2d(smem_addr(sRaw) + rs * RAW_SLOT_BYTES, &tmaps.a[gi], kc * BK, m * BM, &raw_full[rs]);
                }
              }
              const int st = it % SB;
              mbar_wait(&b_empty[st], ((it / SB) & 1u) ^ 1u);
              if (P.dbg & 4) { mbar_arrive(&b_full[st]); continue; }
              mbar_expect_tx(&b_full[st], bytes);
              bulk_g2s(sB + (size_t)st * L::B_STAGE, b_set + ((size_t)n * k_chunks + kc) * L::B_STAGE, bytes, &b_full[st]);
            }
        }
      }
    } else if (warp == 17 || warp == 18) {
      // =============================== MMA issuer ===============================
      const bool dual = MULTI && !resident;       // every stage is consumed by exactly one chain -> chains can alternate
      const uint32_t mi = warp - 17;
      if (mi == 0 || dual) {
      // The whole warp walks the tiles (uniform control flow, so addresses and descriptors live in
      // uniform registers); one elected lane issues the MMAs of a chunk and the commits.
      const uint32_t idesc = umma_idesc(BN);
      const uint32_t sB_u = smem_addr(sB);
      // descriptor template: LBO / SBO / version bits fixed, 14-bit start address added per use
      const uint64_t tmplB = umma_desc(0, BN * 16, 128);
      uint32_t a_it = 0, b_it = 0, acc_it = 0;
      for (int t = t0; t < t1; ++t) {
        const int n = t % P.n_tiles;
        const bool first_m = t == t0 || n == 0, last_m = t == t1 - 1 || n == P.n_tiles - 1;
        {
          uint32_t buf = 0;
          for (int kc = 0; kc < k_chunks; ++kc) {
            const bool seg_first = (kc % KSEG) == 0;
            const bool seg_last = (kc % KSEG) == KSEG - 1 || kc == k_chunks - 1;
            if (dual && (acc_it & 1u) != mi) {      // the other issuer's chain
              if (seg_last) ++acc_it;
              ++b_it;
              ++a_it;
              continue;
            }
            if (seg_first) {
              buf = acc_it % NACC;
              mbar_wait(&acc_empty[buf], ((acc_it / NACC) & 1u) ^ 1u);
            }
            const uint32_t a_idx = resident ? a_it + (uint32_t)kc : a_it;
            const uint32_t sa = a_idx % SA, sb = b_it % SB;
            if (!resident || first_m) mbar_wait(&a_full[sa], (a_idx / SA) & 1u);
            mbar_wait(&b_full[sb], (b_it / SB) & 1u);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t tAh = tmem_a0 + sa * (uint32_t)A_COLS, tAl = tAh + BK;
              const uint64_t dBh = tmplB | (uint64_t)(((sB_u + sb * (uint32_t)(L::B_STAGE * 4)) >> 4) & 0x3FFFu);
              const uint64_t dBl = dBh + (uint64_t)((BN * BK * 4) >> 4);
              const uint32_t d = tmem_base + buf * BN;
#pragma unroll
              for (int j = 0; j < ((P.dbg & 2) ? 0 : BK / 8); ++j) {   // one MMA covers K = 8 tf32: 8 TMEM columns of A,
                const uint64_t oB = (uint64_t)((2 * j * BN * 16) >> 4); // two 16-byte columns of B
                umma_tf32_ts(d, tAl + 8 * j, dBh + oB, idesc, (seg_first && j == 0) ? 0u : 1u);
                umma_tf32_ts(d, tAh + 8 * j, dBl + oB, idesc, 1u);
                umma_tf32_ts(d, tAh + 8 * j, dBh + oB, idesc, 1u);
              }
              umma_commit(&b_empty[sb]);
              if (!resident || last_m) umma_commit(&a_empty[sa]);
              if (seg_last) umma_commit(&acc_full[buf]);
            }
            __syncwarp();
            if (seg_last) ++acc_it;
            ++b_it;
            if (!resident) ++a_it;
          }
        }
        if (resident && last_m) a_it += (uint32_t)k_chunks;
      }
      }
    } else if (warp < 8) {
      // =============================== epilogue ===============================
      const bool dense = g.c_s3 == 1 && (g.c_s1 & 3) == 0 && (g.c_s2 & 3) == 0 && (g.N & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && (g.epilogue != 3 || (g.h_ld & 3) == 0);
      EpiCtx c{&P, &tmaps.c[gi], acc_full, acc_empty, sEpi + warp * L::EPI_STAGE, tmem_base, t0, t1, k_chunks};
      if (g.epilogue == 1) epilogue_role<BN, MULTI, NACC, 1, false>(c);
      else if (g.epilogue == 0) { if (dense) epilogue_role<BN, MULTI, NACC, 0, true>(c); else epilogue_role<BN, MULTI, NACC, 0, false>(c); }
      else if (g.epilogue == 2) { if (dense) epilogue_role<BN, MULTI, NACC, 2, true>(c); else epilogue_role<BN, MULTI, NACC, 2, false>(c); }
      else { if (dense) epilogue_role<BN, MULTI, NACC, 3, true>(c); else epilogue_role<BN, MULTI, NACC, 3, false>(c); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- packing: weight B[n, k] (n = n1 * d + n2 at src + n1 * s1 + n2 * s2 + k * sk) -> tiles
// [n_tile][k_chunk][hi | lo][c (8)][row (BN)][4], zero padded
struct PackBatch {
  e3b_gemm_pack_desc d[MAXG];
  int32_t bn[MAXG];
  int64_t begin[MAXG + 1];   // first float4 slot of each descriptor in the flat index space
  int32_t n;
};

__global__ void gemm_pack_kernel(const __grid_constant__ PackBatch pb) {
  const int64_t total = pb.begin[pb.n];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int gi = 0;
    for (int i = 1; i < pb.n; ++i)
      if (idx >= pb.begin[i]) gi = i;
    const e3b_gemm_pack_desc& d = pb.d[gi];
    const int BN = pb.bn[gi];
    int64_t q = idx - pb.begin[gi];                // float4 slot: [set][n_tile][k_chunk][c][row]
    const int64_t per_set = (pb.begin[gi + 1] - pb.begin[gi]) / (d.n_sets > 1 ? d.n_sets : 1);
    const int64_t set = q / per_set;
    q -= set * per_set;
    const float* src = d.src + set * d.set_stride;
    const int k_chunks = (d.K + BK - 1) / BK;
    const int row = (int)(q % BN);
    const int c = (int)((q / BN) % 8);
    const int kc = (int)((q / (BN * 8)) % k_chunks);
    const int nt = (int)(q / ((int64_t)BN * 8 * k_chunks));
    const int n = nt * BN + row;
    float x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kc * BK + c * 4 + e;
      x[e] = (n < d.N && k < d.K && (d.n2_valid <= 0 || (n % d.d) < d.n2_valid)) ? __ldg(src + (int64_t)(n / d.d) * d.s1 + (int64_t)(n % d.d) * d.s2 + (int64_t)k * d.sk) : 0.f;
    }
    float4 hi, lo;
    split4(make_float4(x[0], x[1], x[2], x[3]), &hi, &lo);
    float4* tile = reinterpret_cast<float4*>(d.dst) + set * per_set * 2 + ((int64_t)nt * k_chunks + kc) * (2 * BN * 8);
    tile[c * BN + row] = hi;
    tile[BN * 8 + c * BN + row] = lo;
  }
}

template <int BN, bool MULTI, int NACC, int SA, int PRAW, int SB>
cudaError_t launch(const Batch& b, const TmapBatch& tm, int ctas, cudaStream_t st) {
  using L = Smem<BN, PRAW, SB>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, MULTI, NACC, SA, PRAW, SB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  static const int pdl = [] { const char* v = getenv("E3B_PDL"); return v ? atoi(v) : 1; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = L::BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<BN, MULTI, NACC, SA, PRAW, SB>, b, tm);
}

int tile_n(int N, int K) { return (K <= KSEG * BK && N > 64) ? 128 : 64; }

// cuTensorMapEncodeTiled through the runtime (no link dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// a row-major fp32 matrix [rows, cols] with a row pitch of `pitch` floats as a {cols, rows} tensor accessed in boxes of
// {32, box_rows} with the 128-byte swizzle (out-of-range parts read as zero / are not written)
bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t pitch, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" int e3b_gemm_tile_n(int32_t N, int32_t K) { return tile_n(N, K); }

extern "C" int64_t e3b_gemm_packed_floats(int32_t N, int32_t K) {
  const int bn = tile_n(N, K);
  return (int64_t)((N + bn - 1) / bn) * ((K + BK - 1) / BK) * 2 * bn * BK;
}

extern "C" int e3b_gemm_pack(const e3b_gemm_pack_desc* descs, int32_t n, void* stream) {
  if (n <= 0) return E3B_OK;
  if (!descs || n > MAXG) return e3b_fail(E3B_ERR_INVALID, "gemm_pack: 1..%d descriptors", MAXG);
  PackBatch pb;
  pb.n = n;
  pb.begin[0] = 0;
  for (int i = 0; i < n; ++i) {
    const e3b_gemm_pack_desc& d = descs[i];
    if (!d.src || !d.dst || d.N <= 0 || d.K <= 0 || d.d <= 0) return e3b_fail(E3B_ERR_INVALID, "gemm_pack: bad descriptor %d", i);
    if (reinterpret_cast<uintptr_t>(d.dst) & 127) return e3b_fail(E3B_ERR_INVALID, "gemm_pack: dst must be 128-byte aligned");
    pb.d[i] = d;
    pb.bn[i] = tile_n(d.N, d.K);
    pb.begin[i + 1] = pb.begin[i] + (d.n_sets > 1 ? d.n_sets : 1) * (e3b_gemm_packed_floats(d.N, d.K) / 8);   // one float4 slot feeds hi and lo
  }
  const int64_t total = pb.begin[n];
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  gemm_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pb);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e3b_fail(E3B_ERR_CUDA, "gemm_pack: %s", cudaGetErrorString(e));
  return E3B_OK;
}

extern "C" int e3b_gemm_run(const e3b_gemm_problem* problems, int32_t n, void* stream) {
  if (n <= 0) return E3B_OK;
  if (!problems || n > MAXG) return e3b_fail(E3B_ERR_INVALID, "gemm_run: 1..%d problems per launch", MAXG);
  Batch b;
  static TmapBatch tm;            // entries are (re)written for the problems that use them; the launch copies the struct
  b.n = 0;
  int bn = 0;
  bool multi = false;
  double work[MAXG], total_work = 0;
  for (int i = 0; i < n; ++i) {
    const e3b_gemm_problem& p = problems[i];
    if (p.M == 0 || p.N == 0) continue;
    if (!p.A || !p.B_packed || !p.C || p.M < 0 || p.N < 0 || p.K <= 0 || p.a_d <= 0 || p.c_d <= 0)
      return e3b_fail(E3B_ERR_INVALID, "gemm_run: bad argument in problem %d", i);
    if ((p.K & 3) || (p.a_s1 & 3) || (p.a_s2 & 3) || (reinterpret_cast<uintptr_t>(p.A) & 15) ||
        (reinterpret_cast<uintptr_t>(p.B_packed) & 127))
      return e3b_fail(E3B_ERR_UNSUPPORTED, "gemm_run: K, the row strides and the base of A must be multiples of 4 floats");
    if (p.epilogue < 0 || p.epilogue > 3) return e3b_fail(E3B_ERR_INVALID, "gemm_run: unknown epilogue %d", p.epilogue);
    if (p.epilogue == 1 && (!p.aux || (p.V != 16 && p.V != 32) || p.aux_d <= 0 || (p.N % p.V) != 0))
      return e3b_fail(E3B_ERR_INVALID, "gemm_run: reduce epilogue needs aux, V in {16, 32} and N %% V == 0");
    if (p.epilogue == 3 && !p.H) return e3b_fail(E3B_ERR_INVALID, "gemm_run: epilogue 3 needs H");
    if ((p.row_map || p.b_sel) && (p.a_d != p.c_d || p.epilogue == 3 || p.epilogue == 1))
      return e3b_fail(E3B_ERR_UNSUPPORTED, "gemm_run: grouped rows need a_d == c_d and epilogue 0 or 2");
    if (p.b_sel && (!p.row_map || p.b_set_stride <= 0 || (p.b_set_stride & 31)))
      return e3b_fail(E3B_ERR_INVALID, "gemm_run: b_sel needs row_map and a set stride that keeps the sets 128-byte aligned");
    const int t = tile_n(p.N, p.K);
    const bool mu = p.K > KSEG * BK;
    if (b.n == 0) { bn = t; multi = mu; }
    else if (bn != t || multi != mu)
      return e3b_fail(E3B_ERR_UNSUPPORTED, "gemm_run: the problems of one launch must share the tile shape "
                      "(K <= 64 or not; N <= 64 or not)");
    Problem& P = b.pr[b.n];
    P.p = p;
    P.m_tiles = (p.M + BM - 1) / BM;
    P.n_tiles = (p.N + t - 1) / t;
    P.k_chunks = (p.K + BK - 1) / BK;
    if (p.a_d >= 512) return e3b_fail(E3B_ERR_UNSUPPORTED, "gemm_run: a_d must be < 512");
    { static const int dbg = [] { const char* v = getenv("E3B_GEMM_DEBUG"); return v ? atoi(v) : 0; }(); P.dbg = dbg; }
    P.a_mul = ((1ull << 40) + (uint64_t)p.a_d - 1) / (uint64_t)p.a_d;
    if (p.c_d >= 512) return e3b_fail(E3B_ERR_UNSUPPORTED, "gemm_run: c_d must be < 512");
    P.c_mul = ((1ull << 40) + (uint64_t)p.c_d - 1) / (uint64_t)p.c_d;
    {
      // TMA for A: K-long (the A ring streams), plain rows of a row-major matrix, no row map
      static const int tma_env = [] { const char* v = getenv("E3B_GEMM_TMA"); return v ? atoi(v) : 1; }();
      P.tma = (tma_env && mu && P.k_chunks > 4 && p.a_d == 1 && !p.row_map && p.a_s1 >= p.K && p.a_s1 < (1ll << 38) &&
               make_map(&tm.a[b.n], p.A, p.M, p.K, p.a_s1, BM)) ? 1 : 0;
      // TMA for C: plain rows of a row-major matrix written by epilogue 0 / 2, large enough to matter
      const bool dense = p.c_s3 == 1 && (p.c_s1 & 3) == 0 && (p.N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0;
      P.tma_c = (tma_env && dense && p.c_d == 1 && !p.row_map && !p.accumulate && (p.epilogue == 0 || p.epilogue == 2) &&
                 p.c_s1 >= p.N && p.c_s1 < (1ll << 38) && (int64_t)p.M * p.N >= (1 << 22) &&
                 make_map(&tm.c[b.n], p.C, p.M, p.N, p.c_s1, 32)) ? 1 : 0;
    }
    work[b.n] = (double)P.m_tiles * P.n_tiles * (P.k_chunks + 2);
    total_work += work[b.n];
    ++b.n;
  }
  if (b.n == 0) return E3B_OK;
  // one persistent CTA per SM in total: CTAs per problem proportional to its work (>= 1, <= its tiles)
  const int target = 148;
  int want[MAXG], sum = 0;
  for (int i = 0; i < b.n; ++i) {
    const int64_t tiles = (int64_t)b.pr[i].m_tiles * b.pr[i].n_tiles;
    int64_t w = (int64_t)(target * work[i] / total_work);
    if (w < 1) w = 1;
    if (w > tiles) w = tiles;
    want[i] = (int)w;
    sum += want[i];
  }
  for (bool moved = true; moved && sum != target;) {   // hand out / take back the rounding remainder
    moved = false;
    int best = -1;
    double best_v = 0;
    for (int i = 0; i < b.n; ++i) {
      const int64_t tiles = (int64_t)b.pr[i].m_tiles * b.pr[i].n_tiles;
      if (sum < target && want[i] < tiles) {
        const double v = work[i] / want[i];                 // most loaded CTAs get help first
        if (best < 0 || v > best_v) { best = i; best_v = v; }
      } else if (sum > target && want[i] > 1) {
        const double v = -work[i] / (want[i] - 1);          // cheapest to shrink
        if (best < 0 || v > best_v) { best = i; best_v = v; }
      }
    }
    if (best >= 0) { want[best] += sum < target ? 1 : -1; sum += sum < target ? 1 : -1; moved = true; }
  }
  int ctas = 0;
  for (int i = 0; i < b.n; ++i) {
    b.pr[i].n_ctas = want[i];
    b.pr[i].cta_begin = ctas;
    ctas += want[i];
  }
  cudaError_t e;
  if (multi) e = launch<64, true, 4, 4, 6, 4>(b, tm, ctas, (cudaStream_t)stream);
  else if (bn == 64) e = launch<64, false, 4, 4, 6, 4>(b, tm, ctas, (cudaStream_t)stream);
  else e = launch<128, false, 2, 4, 4, 3>(b, tm, ctas, (cudaStream_t)stream);
  if (e != cudaSuccess) return e3b_fail(E3B_ERR_CUDA, "gemm_run: %s", cudaGetErrorString(e));
  return E3B_OK;
}

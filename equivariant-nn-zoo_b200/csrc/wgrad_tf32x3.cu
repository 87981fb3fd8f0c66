// Weight gradients on the 5th-generation tensor cores (tcgen05, sm_100a): reductions over ALL rows with a small
// output,
//
//   C[m, n] (+)= alpha * sum_{r < R} A'[r, m] * B[r, n]          A, B: activations / gradients in HBM (fp32)
//
// with A'[r, m] = A[r, m], or, for the self-connection weights W[u, v, w] (m = v * K1 + u),
// A'[r, v * K1 + u] = A[r, u] * aux[r / aux_d, v] -- the (feature x attribute) outer product is formed on the
// fly by the loader and never materialised.
//
// The reduction index r (the MMA's K) is the ROW of the row-major activations, m / n are contiguous, i.e. both
// operands are "MN-major" in memory.  (tcgen05 has MN-major TF32 operand modes, instruction-descriptor bits 15 / 16;
// with the no-swizzle layout they returned zeros on the B200 tried, so the operands are made K-major instead:) loader
// threads take 4 x 4 blocks -- 16 bytes of 4 consecutive rows -- transpose them in registers (free: a renaming), split them into TF32 hi / lo (3xTF32: lo*hi + hi*lo + hi*hi, fp32
// accumulate) and store the canonical no-swizzle K-major core matrices (8 columns x 16 bytes).  HBM latency is hidden by
// a raw ring that copier warps fill with cp.async RAW chunks ahead of the conversion (completion counted by mbarriers);
// the 4 x 4 blocks are read from that ring.  Copying and converting are separate warps on purpose: the proxy fence that
// publishes a converter's shared-memory stores to the tensor core waits for that thread's own loads in flight
// (measured: register- or cp.async-prefetching converters ran at one HBM latency per chunk), and per-row TMA bulk copies
// are issued one elected lane at a time (measured: 1.3 us per 64 copies).
//
// Split-K: the (M tile, N tile) pairs of a problem are shared by several CTAs, each reducing a contiguous
// range of rows; partial tiles go to a workspace and a second kernel sums them in a FIXED order
// (deterministic: no atomics), applies alpha and writes C through its strides.
//
// One CTA = 17 warps, warp-specialised:
//   warps 0-3   accumulate: TMEM -> fp32 partial tile in shared memory every CHAIN chunks (the tensor core accumulates
//               with truncation, so chains are cut), final store of the partial tile
//   warps 4-11  converters: raw ring -> transpose / split -> operand stages
//   warps 12-15 copiers: global -(cp.async)-> raw ring
//   warp  16    MMA issuer (one elected lane), TMEM allocation
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/e3b200.h"

int e3b_fail(int code, const char* fmt, ...);  // e3b200.cu

namespace {

constexpr int BM = 128;            // UMMA M (output rows per tile)
constexpr int BN = 64;             // UMMA N (output columns per tile)
constexpr int BK = 32;             // rows of the reduction per ring stage (4 MMA K-steps of 8)
constexpr int CHAIN = 8;           // stages per accumulation chain (256 rows)
constexpr int STAGES = 2;          // operand (canonical, split) stages
constexpr int RAW = 3;             // raw (row-major, as in HBM) stages filled by cp.async
constexpr int NTHREADS = 544;
constexpr int NCONV = 256;         // converter threads (warps 4-11)
constexpr int NCOPY = 128;         // copier threads (warps 12-15)
constexpr int MMA_WARP = 16;
constexpr int MAXG = E3B_WGRAD_MAX_GROUP;

struct Problem {
  e3b_wgrad_problem p;
  int32_t m_tiles, n_tiles, n_split;
  int32_t cta_begin;               // first CTA of this problem; CTA = (tile, split), split fastest
  int64_t ws_off;                  // floats: partial tiles [tile][split][BM][BN]
  int32_t M;                       // K1 * max(V, 1)
};
struct Batch {
  Problem pr[MAXG];
  int32_t n;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
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
// long waits (the accumulating warps wait a whole chain): back off so that the polling does not take issue slots from
// the converters
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(256);
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// shared-memory descriptor, SWIZZLE_NONE, Blackwell version bits
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// round-to-nearest TF32 by integer arithmetic (see gemm_tf32x3.cu); lo = x - hi is exact in fp32
__device__ __forceinline__ float rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split4(const float4 x, float4* hi, float4* lo) {
  hi->x = rn_tf32(x.x); hi->y = rn_tf32(x.y); hi->z = rn_tf32(x.z); hi->w = rn_tf32(x.w);
  lo->x = x.x - hi->x; lo->y = x.y - hi->y; lo->z = x.z - hi->z; lo->w = x.w - hi->w;
}

// Shared memory.  One operand tile of a stage = MN columns x BK rows, hi then lo, each as K-major no-swizzle core
// matrices (8 columns x 16 bytes = 4 rows of the reduction per column, 128 contiguous bytes); 8-column groups are GRP
// = 144 bytes apart (SBO; the 16 bytes of padding make the transposing stores of a quarter warp hit 8 different bank
// groups), row quads RQ floats apart (LBO):  element (mn, k) at float (k / 4) * RQ + (mn / 8) * 36 + (mn % 8) * 4 + k % 4.
constexpr int GRP = 36;                          // floats between 8-column groups
constexpr int RQ_A = (BM / 8) * GRP, RQ_B = (BN / 8) * GRP;
constexpr int A_HALF = (BK / 4) * RQ_A;          // floats of the hi (or lo) part of A
constexpr int B_HALF = (BK / 4) * RQ_B;
constexpr int STAGE = 2 * A_HALF + 2 * B_HALF;
constexpr int RAW_STAGE = (BM + BN) * BK;        // floats: [BK rows][BM columns of A] then [BK rows][BN columns of B]
constexpr int ACC_LD = BN + 1;                   // fp32 partial tile in shared memory, padded rows
constexpr size_t RING_BYTES = (size_t)(STAGES * STAGE + RAW * RAW_STAGE + BM * ACC_LD) * 4 + 128;

// 16-byte asynchronous copy; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread have landed (the barrier's count
// must include these arrivals: .noinc)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// transposes the 4 x 4 block (rows v[0..3], 4 columns), splits, stores the 4 column pieces (16 bytes each)
__device__ __forceinline__ void unit_store(const float4 (&v)[4], float* hi_dst, int half) {
  const float4 t4[4] = {make_float4(v[0].x, v[1].x, v[2].x, v[3].x), make_float4(v[0].y, v[1].y, v[2].y, v[3].y),
                        make_float4(v[0].z, v[1].z, v[2].z, v[3].z), make_float4(v[0].w, v[1].w, v[2].w, v[3].w)};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 hi, lo;
    split4(t4[j], &hi, &lo);
    *reinterpret_cast<float4*>(hi_dst + 4 * j) = hi;
    *reinterpret_cast<float4*>(hi_dst + half + 4 * j) = lo;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) wgrad_tf32x3_kernel(const __grid_constant__ Batch batch, float* __restrict__ ws) {
  extern __shared__ unsigned char smem_dyn[];
  float* ring = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[STAGES], empty[STAGES], raw_full[RAW], raw_empty[RAW], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int gi = 0;
#pragma unroll 1
  for (int i = 1; i < batch.n; ++i)
    if ((int)blockIdx.x >= batch.pr[i].cta_begin) gi = i;
  const Problem& P = batch.pr[gi];
  const e3b_wgrad_problem& g = P.p;
  const int local = (int)blockIdx.x - P.cta_begin;
  const int tile = local / P.n_split, split = local - tile * P.n_split;
  const int mt = tile / P.n_tiles, nt = tile - mt * P.n_tiles;
  // rows of this split, in whole chunks
  const int64_t chunks_all = (g.R + BK - 1) / BK;
  const int64_t c0 = chunks_all * split / P.n_split, c1 = chunks_all * (split + 1) / P.n_split;
  const int n_chunks = (int)(c1 - c0);

  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_smem)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NCONV / 32); mbar_init(&empty[s], 1); }
    for (int s = 0; s < RAW; ++s) { mbar_init(&raw_full[s], NCOPY); mbar_init(&raw_empty[s], NCONV / 32); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // raw tile geometry: which columns of A / B rows a chunk needs (see the producer)
  const int u_lo = g.V > 0 ? (g.K1 >= BM ? (mt * BM) % g.K1 : 0) : mt * BM;
  const int a_w = g.V > 0 ? (g.K1 >= BM ? BM : g.K1) : (g.K1 - mt * BM < BM ? g.K1 - mt * BM : BM);     // floats per raw A row
  const int b_w = g.K2 - nt * BN < BN ? g.K2 - nt * BN : BN;
  float* raw = ring + STAGES * STAGE;
  float* acc_s = raw + RAW * RAW_STAGE;                         // [BM][ACC_LD] fp32 partial sums of this CTA
  if (warp >= 12 && warp < 16) {
    // =============================== copiers ===============================
    // cp.async copies the chunk's rows as they are in HBM -- 16-byte pieces, consecutive lanes = consecutive pieces of
    // a row -- into the raw ring, RAW chunks ahead of the converters; each thread's completion is counted by the slot's
    // mbarrier (cp.async.mbarrier.arrive).  Out-of-range rows / columns are zero-filled (src-size 0).  Lane l keeps the
    // address of row l of the chunk for both operands (quotient / remainder of the affine row addressing advanced
    // incrementally, no division in the loop); a piece fetches its row's offset with a shuffle.
    const int t = tid - 384;                                    // 0..127
    const int w = t >> 5;
    const int acq = lane, bcq = lane & 15;
    const bool a_col_ok = 4 * acq < a_w, b_col_ok = 4 * bcq < b_w;
    const uint32_t ad = (uint32_t)g.a_d, bd = (uint32_t)g.b_d;
    int64_t r = c0 * BK + lane;                                  // this lane's row
    uint32_t qa = (uint32_t)r / ad, ra = (uint32_t)r - qa * ad, qb = (uint32_t)r / bd, rb = (uint32_t)r - qb * bd;
    const uint32_t a_dq = BK / ad, a_dr = BK % ad, b_dq = BK / bd, b_dr = BK % bd;
    const float* a_base = g.A + u_lo + 4 * acq;
    const float* b_base = g.B + nt * BN + 4 * bcq;
    for (int c = 0; c < n_chunks; ++c) {
      const int slot = c % RAW;
      const int64_t offA = (int64_t)qa * g.a_s1 + (int64_t)ra * g.a_s2, offB = (int64_t)qb * g.b_s1 + (int64_t)rb * g.b_s2;
      const int row_ok = r < g.R;
      mbar_wait(&raw_empty[slot], (((uint32_t)c / RAW) & 1u) ^ 1u);
      float* rs = raw + (size_t)slot * RAW_STAGE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = w + 4 * i;
        const int64_t off = __shfl_sync(0xffffffffu, offA, row);
        const int rok = __shfl_sync(0xffffffffu, row_ok, row);     // (every lane takes part in the shuffle)
        const bool ok = a_col_ok && rok;
        cp_async16(rs + row * BM + 4 * acq, ok ? a_base + off : g.A, ok ? 16u : 0u);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = 2 * w + (lane >> 4) + 8 * i;
        const int64_t off = __shfl_sync(0xffffffffu, offB, row);
        const int rok = __shfl_sync(0xffffffffu, row_ok, row);
        const bool ok = b_col_ok && rok;
        cp_async16(rs + BK * BM + row * BN + 4 * bcq, ok ? b_base + off : g.B, ok ? 16u : 0u);
      }
      cp_async_arrive(&raw_full[slot]);
      r += BK;
      qa += a_dq; ra += a_dr; if (ra >= ad) { ra -= ad; ++qa; }
      qb += b_dq; rb += b_dr; if (rb >= bd) { rb -= bd; ++qb; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp >= 4 && warp < 12) {
    // =============================== converters ===============================
    // The tensor core wants K-major operands (the reduction index contiguous in 16-byte pieces) while the activations
    // have the reduction index as the ROW: a thread reads a 4 x 4 block of the raw tile -- 16 bytes of 4 consecutive
    // rows -- transposes it in registers, splits hi / lo and stores the 4 column pieces into the operand stage.  Units
    // (row quad, column quad): the A tile has 256 (128 when it has <= 64 valid columns), the B tile 128; every thread
    // owns at most one of each.  These threads have no asynchronous copies of their own in flight.
    const int t = tid - 128;                                    // 0..255
    const int m_valid = P.M - mt * BM;                          // > 0
    const bool wideA = m_valid > 64;
    int a_rq, a_cq, b_rq, b_cq;
    bool a_on, b_on;
    if (wideA) { a_rq = t >> 5; a_cq = t & 31; a_on = true; b_rq = (t >> 4) & 7; b_cq = t & 15; b_on = t < 128; }
    else { a_rq = (t >> 4) & 7; a_cq = t & 15; a_on = t < 128; b_rq = (t >> 4) & 7; b_cq = t & 15; b_on = t >= 128; }
    const int m0 = mt * BM + 4 * a_cq;
    const bool a_ok = m0 < P.M;                                 // else the raw columns are zeros already
    const int a_v = g.V > 0 && a_ok ? m0 / g.K1 : 0;
    const int a_col = a_ok ? (g.V > 0 ? m0 - a_v * g.K1 : m0) - u_lo : 0;      // column of the raw A row
    const bool a_aux = g.V > 0 && a_ok && a_on;
    const int a_src = (4 * a_rq) * BM + a_col, b_src = BK * BM + (4 * b_rq) * BN + 4 * b_cq;
    const int a_dst = a_rq * RQ_A + (a_cq >> 1) * GRP + (a_cq & 1) * 16;
    const int b_dst = 2 * A_HALF + b_rq * RQ_B + (b_cq >> 1) * GRP + (b_cq & 1) * 16;
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int64_t r0 = (c0 + c) * BK;
      float ax[4] = {1.f, 1.f, 1.f, 1.f};
      if (a_aux) {                                              // attribute factors of this thread's rows (L1 / L2 hits)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t r = r0 + 4 * a_rq + i;
          ax[i] = r < g.R ? __ldg(g.aux + (int64_t)((uint32_t)r / (uint32_t)g.aux_d) * g.aux_ld + a_v) : 0.f;
        }
      }
      const int slot = c % RAW, s = c % STAGES;
      mbar_wait(&raw_full[slot], ((uint32_t)c / RAW) & 1u);
      mbar_wait(&empty[s], (((uint32_t)c / STAGES) & 1u) ^ 1u);
      const float* rs = raw + (size_t)slot * RAW_STAGE;
      float* st = ring + (size_t)s * STAGE;
      if (a_on) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = *reinterpret_cast<const float4*>(rs + a_src + i * BM);
          v[i].x *= ax[i]; v[i].y *= ax[i]; v[i].z *= ax[i]; v[i].w *= ax[i];
        }
        unit_store(v, st + a_dst, A_HALF);
      }
      if (b_on) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(rs + b_src + i * BN);
        unit_store(v, st + b_dst, B_HALF);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) { mbar_arrive(&full[s]); mbar_arrive(&raw_empty[slot]); }
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = umma_idesc(BN);
    const uint32_t ring_u = smem_addr(ring);
    uint32_t acc_it = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % STAGES;
      const bool first = (c % CHAIN) == 0, last = (c % CHAIN) == CHAIN - 1 || c == n_chunks - 1;
      const uint32_t buf = acc_it & 1u;
      if (first) mbar_wait(&acc_empty[buf], ((acc_it >> 1) & 1u) ^ 1u);
      mbar_wait(&full[s], ((uint32_t)c / STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = ring_u + (uint32_t)s * (uint32_t)(STAGE * 4);
        const uint32_t sb = sa + 2u * A_HALF * 4u;
        const uint32_t d = tmem_base + buf * BN;
        // K-major, no swizzle: core matrix = 8 columns x 16 bytes (4 rows of the reduction); LBO = bytes between the two
        // row quads of one MMA (K = 8), SBO = bytes between consecutive 8-column groups
        const uint64_t dA = umma_desc(sa, RQ_A * 4, GRP * 4), dB = umma_desc(sb, RQ_B * 4, GRP * 4);
#pragma unroll
        for (int kg = 0; kg < BK / 8; ++kg) {
          const uint64_t oA = (uint64_t)((2 * kg * RQ_A * 4) >> 4), oB = (uint64_t)((2 * kg * RQ_B * 4) >> 4);
          const uint64_t dAh = dA + oA, dAl = dAh + (uint64_t)((A_HALF * 4) >> 4);
          const uint64_t dBh = dB + oB, dBl = dBh + (uint64_t)((B_HALF * 4) >> 4);
          umma_tf32_ss(d, dAl, dBh, idesc, (first && kg == 0) ? 0u : 1u);
          umma_tf32_ss(d, dAh, dBl, idesc, 1u);
          umma_tf32_ss(d, dAh, dBh, idesc, 1u);
        }
        umma_commit(&empty[s]);
        if (last) umma_commit(&acc_full[buf]);
      }
      __syncwarp();
      if (last) ++acc_it;
    }
  } else if (warp < 4) {
    // =============================== accumulate + store ===============================
    // thread = one row of the tile; its partial sums live in a padded shared-memory row (conflict-free, private)
    const int q = warp;
    float* my = acc_s + (q * 32 + lane) * ACC_LD;
    const int n_chains = (n_chunks + CHAIN - 1) / CHAIN;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int ch = 0; ch < n_chains; ++ch) {
      const uint32_t buf = (uint32_t)ch & 1u;
      mbar_wait_sleep(&acc_full[buf], ((uint32_t)ch >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < BN; cb += 16) {
        float v[16];
        tmem_ld16(t_lane + buf * BN + (uint32_t)cb, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) my[cb + i] = ch == 0 ? v[i] : my[cb + i] + v[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    // partial tile [BM][BN] of (tile, split): a warp writes its 32 rows, 128 contiguous bytes per instruction and row pair
    __syncwarp();
    float* dst = ws + P.ws_off + ((int64_t)(tile * P.n_split + split) * BM + q * 32) * BN;
    const float* src = acc_s + (q * 32) * ACC_LD;
    for (int i = lane; i < 32 * BN; i += 32) {
      const int row = i / BN, col = i - row * BN;
      dst[i] = n_chains > 0 ? src[row * ACC_LD + col] : 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// sums the partial tiles in split order, scales, writes C through its strides
__global__ void wgrad_reduce_kernel(const __grid_constant__ Batch batch, const float* __restrict__ ws, int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    // flat index over problems: [problem][m][n padded to n_tiles * BN]
    int64_t rem = idx;
    int gi = 0;
    for (; gi < batch.n; ++gi) {
      const int64_t sz = (int64_t)batch.pr[gi].M * batch.pr[gi].n_tiles * BN;
      if (rem < sz) break;
      rem -= sz;
    }
    const Problem& P = batch.pr[gi];
    const e3b_wgrad_problem& g = P.p;
    const int ncols = P.n_tiles * BN;
    const int m = (int)(rem / ncols), n = (int)(rem - (int64_t)m * ncols);
    if (n >= g.K2) continue;
    const int mt = m / BM, nt = n / BN;
    const int tile = mt * P.n_tiles + nt;
    const float* src = ws + P.ws_off + ((int64_t)tile * P.n_split * BM + (m - mt * BM)) * BN + (n - nt * BN);
    float acc = 0.f;
    for (int s = 0; s < P.n_split; ++s) acc += src[(int64_t)s * BM * BN];
    float* c = g.C + (int64_t)(m / g.c_d) * g.c_s1 + (int64_t)(m % g.c_d) * g.c_s2 + (int64_t)n * g.c_s3;
    const float o = g.alpha * acc;
    *c = g.accumulate ? *c + o : o;
  }
}

// fills the launch plan; returns the workspace floats needed
int64_t plan(const e3b_wgrad_problem* problems, int32_t n, Batch* b, int* ctas_out, int64_t* reduce_total) {
  b->n = 0;
  double work[MAXG], total_work = 0;
  for (int i = 0; i < n; ++i) {
    const e3b_wgrad_problem& p = problems[i];
    if (p.R <= 0 || p.K1 <= 0 || p.K2 <= 0) continue;
    Problem& P = b->pr[b->n];
    P.p = p;
    P.M = p.K1 * (p.V > 0 ? p.V : 1);
    P.m_tiles = (P.M + BM - 1) / BM;
    P.n_tiles = (p.K2 + BN - 1) / BN;
    work[b->n] = (double)P.m_tiles * P.n_tiles * (double)((p.R + BK - 1) / BK);
    total_work += work[b->n];
    ++b->n;
  }
  if (b->n == 0) { *ctas_out = 0; *reduce_total = 0; return 0; }
  const int target = 148;
  int ctas = 0;
  int64_t ws = 0, rt = 0;
  for (int i = 0; i < b->n; ++i) {
    Problem& P = b->pr[i];
    const int tiles = P.m_tiles * P.n_tiles;
    const int64_t chunks = (P.p.R + BK - 1) / BK;
    int64_t want = (int64_t)(target * work[i] / total_work) / tiles;      // CTAs per tile
    if (want < 1) want = 1;
    if (want > (chunks + 3) / 4) want = (chunks + 3) / 4;                  // >= 4 chunks (128 rows) per CTA
    if (want < 1) want = 1;
    P.n_split = (int)want;
    P.cta_begin = ctas;
    P.ws_off = ws;
    ctas += tiles * P.n_split;
    ws += (int64_t)tiles * P.n_split * BM * BN;
    rt += (int64_t)P.M * P.n_tiles * BN;
  }
  *ctas_out = ctas;
  *reduce_total = rt;
  return ws;
}

bool valid(const e3b_wgrad_problem& p) {
  if (p.R < 0 || p.K1 < 0 || p.K2 < 0) return false;
  if (p.R == 0 || p.K1 == 0 || p.K2 == 0) return true;
  if (!p.A || !p.B || !p.C || p.a_d <= 0 || p.b_d <= 0 || p.c_d <= 0 || p.R >= (1ll << 31)) return false;
  if ((p.K1 & 3) || (p.K2 & 3) || (p.a_s1 & 3) || (p.a_s2 & 3) || (p.b_s1 & 3) || (p.b_s2 & 3) ||
      (reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.B) & 15)) return false;
  if (p.V > 0 && (!p.aux || p.aux_d <= 0 || (p.K1 > BM && p.K1 % BM))) return false;
  return true;
}

}  // namespace

extern "C" int64_t e3b_wgrad_workspace_floats(const e3b_wgrad_problem* problems, int32_t n) {
  if (n <= 0) return 0;
  if (!problems || n > MAXG) return -1;
  for (int i = 0; i < n; ++i)
    if (!valid(problems[i])) return -1;
  Batch b;
  int ctas;
  int64_t rt;
  return plan(problems, n, &b, &ctas, &rt);
}

extern "C" int e3b_wgrad_run(const e3b_wgrad_problem* problems, int32_t n, float* workspace, void* stream) {
  if (n <= 0) return E3B_OK;
  if (!problems || n > MAXG) return e3b_fail(E3B_ERR_INVALID, "wgrad_run: 1..%d problems per launch", MAXG);
  for (int i = 0; i < n; ++i)
    if (!valid(problems[i]))
      return e3b_fail(E3B_ERR_UNSUPPORTED, "wgrad_run: problem %d: K1, K2, the row strides and the bases of A / B must be "
                      "multiples of 4 floats, R < 2^31", i);
  Batch b;
  int ctas;
  int64_t rt;
  plan(problems, n, &b, &ctas, &rt);
  if (ctas == 0) return E3B_OK;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return e3b_fail(E3B_ERR_INVALID, "wgrad_run: workspace of e3b_wgrad_workspace_floats() floats, 16-byte aligned");
  static bool attr_set = false;
  cudaStream_t st = (cudaStream_t)stream;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RING_BYTES);
    if (e != cudaSuccess) return e3b_fail(E3B_ERR_CUDA, "wgrad_run: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  wgrad_tf32x3_kernel<<<ctas, NTHREADS, RING_BYTES, st>>>(b, workspace);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e3b_fail(E3B_ERR_CUDA, "wgrad_run: %s", cudaGetErrorString(e));
  const int blocks = (int)((rt + 255) / 256 < 148 * 8 ? (rt + 255) / 256 : 148 * 8);
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(b, workspace, rt);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e3b_fail(E3B_ERR_CUDA, "wgrad_run: %s", cudaGetErrorString(e));
  return E3B_OK;
}

// fp32-faithful dense contractions on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   C[r, n] = epilogue( sum_k A[r, k] * B[n, k] )          A, B, C fp32 in HBM
//
// Each fp32 operand is split in registers into a TF32 "hi" part (top 19 bits) and a TF32 "lo"
// part (the exactly representable remainder, truncated), written to shared memory in the UMMA
// canonical K-major / no-swizzle layout, and three tcgen05.mma.kind::tf32 products
// (hi*hi + hi*lo + lo*hi) accumulate in TMEM in fp32 -- "3xTF32", relative error ~1e-6, which
// the 1e-5 parity budget of the interaction block needs (plain TF32 gives ~1e-3).
//
// One CTA = 128 threads = one 128-row tile; it walks its column tiles (BN columns each) and the
// K chunks (32 floats) synchronously: cooperative global->register->(split)->shared stores,
// fence.proxy.async, one elected thread issues the MMAs and commits to an mbarrier, then the
// four warps drain their 32 TMEM lanes with tcgen05.ld and run the epilogue.  Latency is hidden
// by co-resident CTAs (64 KB shared memory and <= 128 TMEM columns per CTA -> 3 per SM).
//
// Row addressing of A and C is affine in (r / d, r % d) so that the irreps layouts of the
// interaction block ([node][component][channel] rows) are read and written in place.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/e3b200.h"

int e3b_fail(int code, const char* fmt, ...);  // e3b200.cu

namespace {

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  const float* aux;
  int64_t a_s1, a_s2;
  int64_t ldb;
  int64_t c_s1, c_s2, c_s3;
  int64_t aux_ld;
  int32_t a_d, c_d, aux_d;
  int32_t M, N, K;
  int32_t epilogue, V;
  float alpha;
};

constexpr int BM = 128;   // UMMA M
constexpr int BK = 32;    // floats per K chunk = 4 MMA K-steps of 8
constexpr int NTHREADS = 128;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B (128 B
// contiguous); LBO = bytes between the two 16-byte K chunks of one MMA, SBO = bytes between
// consecutive 8-row groups; both encoded >> 4; bits 46-47 = 1 (Blackwell descriptor version).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init1(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

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
// exact in fp32).  Rounding (not masking) keeps the residual unbiased -- with truncation the
// error grows linearly in K.
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ void split4(const float4 x, float4* hi, float4* lo) {
  hi->x = rna_tf32(x.x); hi->y = rna_tf32(x.y); hi->z = rna_tf32(x.z); hi->w = rna_tf32(x.w);
  lo->x = rna_tf32(x.x - hi->x); lo->y = rna_tf32(x.y - hi->y);
  lo->z = rna_tf32(x.z - hi->z); lo->w = rna_tf32(x.w - hi->w);
}

// stage rows x BK fp32 (row pointers via `row_ptr(r)`, nullptr = zero row) into the canonical
// layout [16-byte chunk c (8)][row][16 B]; lanes take consecutive rows -> conflict-free STS.128
template <int ROWS, typename RowPtr>
__device__ __forceinline__ void stage_operand(float* s_hi, float* s_lo, RowPtr row_ptr, int k0, int K) {
  for (int idx = threadIdx.x; idx < ROWS * (BK / 4); idx += NTHREADS) {
    const int r = idx % ROWS, c = idx / ROWS;
    const int k = k0 + 4 * c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = row_ptr(r);
    if (p != nullptr && k < K) v = __ldg(reinterpret_cast<const float4*>(p + k));
    float4 hi, lo;
    split4(v, &hi, &lo);
    reinterpret_cast<float4*>(s_hi)[c * ROWS + r] = hi;
    reinterpret_cast<float4*>(s_lo)[c * ROWS + r] = lo;
  }
}

// MULTI: K is long -> the TMEM accumulation chain is cut every KACC floats and the partial sums
// are added in fp32 registers (round-to-nearest).  The tensor core accumulates with truncation,
// so an unbroken chain over K = 1920 drifts by ~1.5e-5 (measured); chains of 64 stay below 2e-6.
constexpr int KACC = 64;

template <int BN, bool MULTI>
__global__ void __launch_bounds__(NTHREADS) gemm_tf32x3_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* sA_hi = reinterpret_cast<float*>(smem_raw);
  float* sA_lo = sA_hi + BM * BK;
  float* sB_hi = sA_lo + BM * BK;
  float* sB_lo = sB_hi + BN * BK;
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * BM;

  if (warp == 0) {  // TMEM allocation is warp-collective; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_smem)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init1(&mma_bar);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  // this thread's output row (TMEM lane = 32 * warp + lane)
  const int row = m0 + tid;
  const bool row_ok = row < g.M;
  float* c_row = nullptr;
  if (row_ok) c_row = g.C + (int64_t)(row / g.c_d) * g.c_s1 + (int64_t)(row % g.c_d) * g.c_s2;
  float aux[32];
  if (g.epilogue == 1) {
#pragma unroll
    for (int v = 0; v < 32; ++v) aux[v] = (row_ok && v < g.V) ? __ldg(g.aux + (int64_t)(row / g.aux_d) * g.aux_ld + v) : 0.f;
  }

  const uint32_t idesc = umma_idesc(BN);
  uint32_t phase = 0;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int n_chunks = (g.K + BK - 1) / BK;

  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int tile = blockIdx.y; tile < n_tiles; tile += gridDim.y) {
    const int n0 = tile * BN;
    float racc[MULTI ? BN : 1];
    if (MULTI) {
#pragma unroll
      for (int i = 0; i < (MULTI ? BN : 1); ++i) racc[i] = 0.f;
    }
    for (int kc = 0; kc < n_chunks; ++kc) {
      const int k0 = kc * BK;
      stage_operand<BM>(sA_hi, sA_lo, [&](int r) -> const float* {
        const int rr = m0 + r;
        return rr < g.M ? g.A + (int64_t)(rr / g.a_d) * g.a_s1 + (int64_t)(rr % g.a_d) * g.a_s2 : nullptr;
      }, k0, g.K);
      stage_operand<BN>(sB_hi, sB_lo, [&](int r) -> const float* {
        const int nn = n0 + r;
        return nn < g.N ? g.B + (int64_t)nn * g.ldb : nullptr;
      }, k0, g.K);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor-core reads
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_addr(sA_hi), a_lo = smem_addr(sA_lo), b_hi = smem_addr(sB_hi), b_lo = smem_addr(sB_lo);
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {   // one MMA covers K = 8 tf32 = two 16-byte chunks
          const uint32_t offA = (uint32_t)(2 * j) * BM * 16, offB = (uint32_t)(2 * j) * BN * 16;
          const uint64_t dAh = umma_desc(a_hi + offA, BM * 16, 128), dAl = umma_desc(a_lo + offA, BM * 16, 128);
          const uint64_t dBh = umma_desc(b_hi + offB, BN * 16, 128), dBl = umma_desc(b_lo + offB, BN * 16, 128);
          const bool first = MULTI ? ((kc % (KACC / BK)) | j) == 0 : (kc | j) == 0;
          umma_tf32(tmem_base, dAl, dBh, idesc, first ? 0u : 1u);
          umma_tf32(tmem_base, dAh, dBl, idesc, 1u);
          umma_tf32(tmem_base, dAh, dBh, idesc, 1u);
        }
        umma_commit(&mma_bar);   // arrives when the MMAs above have finished reading smem / writing TMEM
      }
      mbar_wait_parity(&mma_bar, phase);
      phase ^= 1u;
      if (MULTI && ((kc % (KACC / BK)) == KACC / BK - 1 || kc == n_chunks - 1)) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int cb = 0; cb < (MULTI ? BN : 0); cb += 32) {
          float v[32];
          tmem_ld32(t_row + (uint32_t)cb, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) racc[cb + i] += v[i];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    }
    // ---- epilogue: TMEM (or the register partial sums) -> global
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int cb = 0; cb < BN; cb += 32) {
      float v[32];
      if (MULTI) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = racc[(MULTI ? cb : 0) + (MULTI ? i : 0)];
      } else {
        tmem_ld32(t_row + (uint32_t)cb, v);
      }
      if (!row_ok) continue;
      if (g.epilogue == 0) {
        const int nb = n0 + cb;
        if (g.c_s3 == 1 && nb + 32 <= g.N && ((reinterpret_cast<uintptr_t>(c_row + nb) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(c_row + nb + i) =
                make_float4(g.alpha * v[i], g.alpha * v[i + 1], g.alpha * v[i + 2], g.alpha * v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (nb + i < g.N) c_row[(int64_t)(nb + i) * g.c_s3] = g.alpha * v[i];
        }
      } else {
        // weighted reduction over groups of V accumulator columns: out col = (n0 + cb) / V (+1)
        if (g.V == 16) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int t = 0; t < 16; ++t) { s0 = fmaf(aux[t], v[t], s0); s1 = fmaf(aux[t], v[16 + t], s1); }
          const int oc = (n0 + cb) / 16;
          if (oc * 16 < g.N) c_row[(int64_t)oc * g.c_s3] = g.alpha * s0;
          if ((oc + 1) * 16 < g.N) c_row[(int64_t)(oc + 1) * g.c_s3] = g.alpha * s1;
        } else {
          float s0 = 0.f;
#pragma unroll
          for (int t = 0; t < 32; ++t) s0 = fmaf(aux[t], v[t], s0);
          const int oc = (n0 + cb) / 32;
          if (oc * 32 < g.N) c_row[(int64_t)oc * g.c_s3] = g.alpha * s0;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // all TMEM reads retired before the next tile's MMAs overwrite the accumulator
  }

  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

template <int BN, bool MULTI>
int launch(const GemmArgs& g, cudaStream_t st) {
  const size_t smem = (size_t)(2 * BM * BK + 2 * BN * BK) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const int m_tiles = (g.M + BM - 1) / BM;
  const int n_tiles = (g.N + BN - 1) / BN;
  // enough CTAs to fill the machine: split the column tiles when there are few row tiles
  int ny = 1;
  while (m_tiles * ny < 3 * 148 && ny < n_tiles) ++ny;
  dim3 grid((unsigned)m_tiles, (unsigned)ny);
  gemm_tf32x3_kernel<BN, MULTI><<<grid, NTHREADS, smem, st>>>(g);
  return 0;
}

}  // namespace

extern "C" int e3b_gemm_tf32x3(const float* A, int64_t a_s1, int64_t a_s2, int32_t a_d, const float* B, int64_t ldb,
                               float* C, int64_t c_s1, int64_t c_s2, int32_t c_d, int64_t c_s3, int32_t M, int32_t N,
                               int32_t K, float alpha, int32_t epilogue, const float* aux, int64_t aux_ld,
                               int32_t aux_d, int32_t V, void* stream) {
  if (M == 0 || N == 0) return E3B_OK;
  if (!A || !B || !C || M < 0 || N < 0 || K <= 0 || a_d <= 0 || c_d <= 0)
    return e3b_fail(E3B_ERR_INVALID, "gemm_tf32x3: bad argument");
  if ((K & 3) || (a_s1 & 3) || (a_s2 & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(B) & 15))
    return e3b_fail(E3B_ERR_UNSUPPORTED, "gemm_tf32x3: K, row strides and bases must be multiples of 4 floats (16 B)");
  if (epilogue == 1 && (!aux || (V != 16 && V != 32) || aux_d <= 0 || (N % V) != 0))
    return e3b_fail(E3B_ERR_INVALID, "gemm_tf32x3: reduce epilogue needs aux, V in {16, 32} and N %% V == 0");
  if (epilogue != 0 && epilogue != 1) return e3b_fail(E3B_ERR_INVALID, "gemm_tf32x3: unknown epilogue %d", epilogue);
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.aux = aux;
  g.a_s1 = a_s1; g.a_s2 = a_s2; g.a_d = a_d; g.ldb = ldb;
  g.c_s1 = c_s1; g.c_s2 = c_s2; g.c_d = c_d; g.c_s3 = c_s3;
  g.aux_ld = aux_ld; g.aux_d = aux_d > 0 ? aux_d : 1;
  g.M = M; g.N = N; g.K = K; g.epilogue = epilogue; g.V = V > 0 ? V : 32; g.alpha = alpha;
  if (K > KACC) launch<64, true>(g, (cudaStream_t)stream);        // long K: 64-column tiles, register partial sums
  else if (N <= 64) launch<64, false>(g, (cudaStream_t)stream);
  else launch<128, false>(g, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e3b_fail(E3B_ERR_CUDA, "gemm_tf32x3: %s", cudaGetErrorString(e));
  return E3B_OK;
}

// Hidden layers of the radial MLP (e3nn nn.FullyConnectedNet as built at nn/message_passing.py:74-79, applied
// at :93) fused into one kernel per direction:
//   h_0 = x [E, K0];   h_{i+1} = cst * ssp(alpha_i * h_i W_i),  i = 0 .. L-1   (W_0 [K0,H], W_i [H,H])
// These layers are tiny (K0 = 8 or 32, H = 32 or 64: <= 17 KFLOP per edge against 246 KFLOP for the last layer)
// and were bound by launch latency and by streaming the [E,H] activations between launches.  Here a thread owns
// one edge: its activation row stays in registers across the layers, the weights (<= 40 KB) sit in shared memory
// with the 1/sqrt(fan_in) factor folded in, and every weight fetch is a 16-byte broadcast feeding four FMAs.
// Plain fp32 FMAs on the CUDA cores: exact fp32 (no TF32 split needed) and still far below the HBM time of the
// kernels around it.  The backward runs the chain in reverse from the gradient with respect to the last hidden
// pre-activation (produced by the tcgen05 GEMM of the last layer, epilogue 3) down to d/dx.
#include "common.cuh"
#include "../../include/e3b200.h"

int e3b_fail(int code, const char* fmt, ...);

namespace {

__device__ __forceinline__ float ssp_f(float z) {
  return (z > 15.f ? z : __logf(1.f + __expf(z))) - 0.6931471805599453f;
}
// d/dz [cst * ssp(z)] through the stored output h = cst * ssp(z): sigmoid(z) = 1 - 0.5 exp(-h / cst)
__device__ __forceinline__ float dssp_from_out(float h, float cst, float inv_cst) {
  return cst * (1.f - 0.5f * __expf(-h * inv_cst));
}

struct MlpArgs {
  const float* W[E3B_MLP_MAX_HIDDEN];
  float alpha[E3B_MLP_MAX_HIDDEN];
  float* h_out[E3B_MLP_MAX_HIDDEN];          // fwd: h_{i+1} (nullable);  bwd: d/dz_i for i >= 1 (nullable), index i
  const float* h_saved[E3B_MLP_MAX_HIDDEN];  // bwd: h_i for i >= 1 (index i)
  const float* x;                            // fwd: [E, K0]
  const float* g_top;                        // bwd: d/dz_L  [E, H]
  float* g_x;                                // bwd: [E, K0] (nullable)
  int64_t ldx, n_rows;
  int32_t n_layers;
  float cst;
};

// out[j] = sum_k in[k] * sW[k][j]   (sW row-major [KIN][KOUT] in shared memory, broadcast float4 reads)
template <int KIN, int KOUT>
__device__ __forceinline__ void dense_regs(const float (&in)[KIN], float (&out)[KOUT], const float* __restrict__ sW) {
#pragma unroll
  for (int j = 0; j < KOUT; ++j) out[j] = 0.f;
#pragma unroll
  for (int k = 0; k < KIN; ++k) {
    const float a = in[k];
#pragma unroll
    for (int j = 0; j < KOUT; j += 4) {
      const float4 w = *reinterpret_cast<const float4*>(sW + k * KOUT + j);
      out[j] = fmaf(a, w.x, out[j]); out[j + 1] = fmaf(a, w.y, out[j + 1]);
      out[j + 2] = fmaf(a, w.z, out[j + 2]); out[j + 3] = fmaf(a, w.w, out[j + 3]);
    }
  }
}

template <int N>
__device__ __forceinline__ void load_row(float (&r)[N], const float* __restrict__ p) {
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p + j));
    r[j] = v.x; r[j + 1] = v.y; r[j + 2] = v.z; r[j + 3] = v.w;
  }
}
template <int N>
__device__ __forceinline__ void store_row(const float (&r)[N], float* __restrict__ p) {
#pragma unroll
  for (int j = 0; j < N; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
}

template <int K0, int H>
__global__ void __launch_bounds__(128) mlp_hidden_fwd_kernel(const __grid_constant__ MlpArgs a) {
  extern __shared__ __align__(16) float smem[];   // W_0 [K0][H] | W_1 [H][H] | ...
  for (int l = 0; l < a.n_layers; ++l) {
    const int kin = l == 0 ? K0 : H;
    float* dst = smem + (l == 0 ? 0 : K0 * H + (l - 1) * H * H);
    for (int i = threadIdx.x; i < kin * H; i += blockDim.x) dst[i] = a.W[l][i] * a.alpha[l];
  }
  __syncthreads();
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.n_rows; row += (int64_t)gridDim.x * blockDim.x) {
    float in0[K0], h[H], t[H];
    load_row<K0>(in0, a.x + row * a.ldx);
    dense_regs<K0, H>(in0, t, smem);
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = a.cst * ssp_f(t[j]);
    if (a.h_out[0]) store_row<H>(h, a.h_out[0] + row * H);
    for (int l = 1; l < a.n_layers; ++l) {
      dense_regs<H, H>(h, t, smem + K0 * H + (l - 1) * H * H);
#pragma unroll
      for (int j = 0; j < H; ++j) h[j] = a.cst * ssp_f(t[j]);
      if (a.h_out[l]) store_row<H>(h, a.h_out[l] + row * H);
    }
  }
}

// g = d/dz_L; for i = L-1 .. 1:  g <- (g W_i^T alpha_i) * act'(h_i) = d/dz_i;  finally g_x = g W_0^T alpha_0
template <int K0, int H>
__global__ void __launch_bounds__(128) mlp_hidden_bwd_kernel(const __grid_constant__ MlpArgs a) {
  extern __shared__ __align__(16) float smem[];   // transposed: W_0^T [H][K0] | W_1^T [H][H] | ...
  for (int l = 0; l < a.n_layers; ++l) {
    const int kin = l == 0 ? K0 : H;
    float* dst = smem + (l == 0 ? 0 : K0 * H + (l - 1) * H * H);
    for (int i = threadIdx.x; i < kin * H; i += blockDim.x) {
      const int k = i / H, j = i - k * H;           // W[k][j] -> WT[j][k]
      dst[j * kin + k] = a.W[l][i] * a.alpha[l];
    }
  }
  __syncthreads();
  const float inv_cst = 1.f / a.cst;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.n_rows; row += (int64_t)gridDim.x * blockDim.x) {
    float g[H], t[H];
    load_row<H>(g, a.g_top + row * H);
    for (int l = a.n_layers - 1; l >= 1; --l) {
      dense_regs<H, H>(g, t, smem + K0 * H + (l - 1) * H * H);
      float hs[H];
      load_row<H>(hs, a.h_saved[l] + row * H);
#pragma unroll
      for (int j = 0; j < H; ++j) g[j] = t[j] * dssp_from_out(hs[j], a.cst, inv_cst);
      if (a.h_out[l]) store_row<H>(g, a.h_out[l] + row * H);
    }
    if (a.g_x) {
      float gx[K0];
      dense_regs<H, K0>(g, gx, smem);
      store_row<K0>(gx, a.g_x + row * K0);
    }
  }
}

template <int K0, int H>
int launch(bool bwd, const MlpArgs& a, cudaStream_t s) {
  const size_t bytes = (size_t)(K0 * H + (a.n_layers - 1) * H * H) * sizeof(float);
  auto fwd_k = mlp_hidden_fwd_kernel<K0, H>;
  auto bwd_k = mlp_hidden_bwd_kernel<K0, H>;
  static bool attr_set = false;       // idempotent; a race only repeats the call
  if (!attr_set) {
    cudaFuncSetAttribute(fwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(bwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_set = true;
  }
  const int64_t tiles = (a.n_rows + 127) / 128;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(tiles < (int64_t)sms * 3 ? tiles : (int64_t)sms * 3);
  if (bwd) bwd_k<<<grid, 128, bytes, s>>>(a);
  else fwd_k<<<grid, 128, bytes, s>>>(a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e3b_fail(E3B_ERR_CUDA, "mlp_hidden: %s", cudaGetErrorString(e));
  return E3B_OK;
}

int dispatch(bool bwd, int k0, int h, const MlpArgs& a, cudaStream_t s) {
  if (k0 == 8 && h == 64) return launch<8, 64>(bwd, a, s);
  if (k0 == 8 && h == 32) return launch<8, 32>(bwd, a, s);
  if (k0 == 32 && h == 64) return launch<32, 64>(bwd, a, s);
  return e3b_fail(E3B_ERR_UNSUPPORTED, "mlp_hidden: no kernel for k_in %d, width %d", k0, h);
}

int fill(const e3b_mlp_hidden_desc* d, MlpArgs* a) {
  if (!d || d->n_layers < 1 || d->n_layers > E3B_MLP_MAX_HIDDEN) return e3b_fail(E3B_ERR_INVALID, "mlp_hidden: bad descriptor");
  if (!e3b_mlp_hidden_supported(d->k_in, d->width, d->n_layers))
    return e3b_fail(E3B_ERR_UNSUPPORTED, "mlp_hidden: no kernel for k_in %d, width %d", d->k_in, d->width);
  for (int l = 0; l < E3B_MLP_MAX_HIDDEN; ++l) {
    a->W[l] = l < d->n_layers ? (const float*)d->W[l] : nullptr;
    a->alpha[l] = l < d->n_layers ? d->alpha[l] : 0.f;
    a->h_out[l] = nullptr; a->h_saved[l] = nullptr;
    if (l < d->n_layers && !a->W[l]) return e3b_fail(E3B_ERR_INVALID, "mlp_hidden: null weight");
  }
  a->n_layers = d->n_layers;
  a->cst = d->act_cst;
  a->x = nullptr; a->g_top = nullptr; a->g_x = nullptr; a->ldx = 0; a->n_rows = 0;
  return E3B_OK;
}

}  // namespace

extern "C" int e3b_mlp_hidden_supported(int32_t k_in, int32_t width, int32_t n_layers) {
  if (n_layers < 1 || n_layers > E3B_MLP_MAX_HIDDEN) return 0;
  return ((k_in == 8 && (width == 64 || width == 32)) || (k_in == 32 && width == 64)) ? 1 : 0;
}

extern "C" int e3b_mlp_hidden_fwd(const e3b_mlp_hidden_desc* d, const void* x, int64_t ldx, int64_t n_rows,
                                  void* const* h_h_out, void* stream) {
  MlpArgs a;
  int rc = fill(d, &a);
  if (rc) return rc;
  if (n_rows == 0) return E3B_OK;
  if (!x || !h_h_out || ldx < d->k_in || (ldx & 3)) return e3b_fail(E3B_ERR_INVALID, "mlp_hidden_fwd: bad argument");
  for (int l = 0; l < d->n_layers; ++l) a.h_out[l] = (float*)h_h_out[l];
  a.x = (const float*)x; a.ldx = ldx; a.n_rows = n_rows;
  return dispatch(false, d->k_in, d->width, a, (cudaStream_t)stream);
}

extern "C" int e3b_mlp_hidden_bwd(const e3b_mlp_hidden_desc* d, const void* g_top, const void* const* h_h_saved,
                                  int64_t n_rows, void* const* h_gz_out, void* g_x, void* stream) {
  MlpArgs a;
  int rc = fill(d, &a);
  if (rc) return rc;
  if (n_rows == 0) return E3B_OK;
  if (!g_top || !h_h_saved) return e3b_fail(E3B_ERR_INVALID, "mlp_hidden_bwd: null argument");
  for (int l = 1; l < d->n_layers; ++l) {
    a.h_saved[l] = (const float*)h_h_saved[l];
    if (!a.h_saved[l]) return e3b_fail(E3B_ERR_INVALID, "mlp_hidden_bwd: null saved activation");
    a.h_out[l] = h_gz_out ? (float*)h_gz_out[l] : nullptr;
  }
  a.g_top = (const float*)g_top; a.g_x = (float*)g_x; a.n_rows = n_rows;
  return dispatch(true, d->k_in, d->width, a, (cudaStream_t)stream);
}

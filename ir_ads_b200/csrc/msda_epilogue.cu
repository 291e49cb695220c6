// msda_epilogue.cu -- fused residual add + LayerNorm, the epilogue of the op's callers.
//
// In the reference's encoder / decoder layers every MSDeformAttn call and every FFN is followed by
//   x = x + identity ; x = LayerNorm(x)
// (detrex/layers/transformer.py:152-192 runs "self_attn", "norm", "ffn", "norm"; the attention module adds its own
// residual, multi_scale_deform_attn.py:363; the FFN adds its identity, detrex/layers/mlp.py:127-132), i.e. an
// elementwise add kernel that writes the sum and a LayerNorm kernel that reads it back.  Here both are one pass:
//   forward   y = LN(a + b) * gamma + beta          reads a, b           writes y, mean[row], rstd[row]
//   backward  dx = dLN/d(a+b) (the gradient of BOTH inputs), dgamma, dbeta
//             reads dy, a, b (the sum is recomputed, never stored)       writes dx, per-CTA partials of dgamma / dbeta
// One warp per row, 16-byte lane accesses, two-pass mean / variance in registers (the row stays in registers), fp32
// arithmetic for float and bfloat16 tensors.  dgamma / dbeta are reduced in a fixed order (per-CTA partial sums, then
// one small kernel), so the whole backward is bit-reproducible.  HBM-bound: 3 tensor passes forward, 4 backward.
#include "msda_host.h"

namespace {
using namespace msda_host;

constexpr int kLnThreads = 256;
constexpr int kLnMaxVec = 8;   // float4 chunks per lane: C <= 32 * 4 * 8 = 1024

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld_f4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_f4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const unsigned*>(&lo);
  u.y = *reinterpret_cast<const unsigned*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
  return v;
}

// NV = float4 chunks per lane (C == 128 * NV exactly when EXACT, else C <= 128 * NV and chunks are predicated)
template <typename T, int NV>
__global__ void __launch_bounds__(kLnThreads)
add_layernorm_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ gamma,
                         const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean,
                         float* __restrict__ rstd, int64_t rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * (kLnThreads / 32);
  const float inv_c = 1.0f / (float)C;
  for (int64_t row = warp0; row < rows; row += stride) {
    const T* pa = a + row * C;
    const T* pb = b + row * C;
    float4 s[NV];
    float sum = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 4;
      s[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) {
        const float4 x = ld_f4(pa + c), r = ld_f4(pb + c);
        s[v] = make_float4(x.x + r.x, x.y + r.y, x.z + r.z, x.w + r.w);
        sum += (s[v].x + s[v].y) + (s[v].z + s[v].w);
      }
    }
    const float mu = warp_sum(sum) * inv_c;
    float var = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 4;
      if (c < C) {
        const float dx = s[v].x - mu, dy = s[v].y - mu, dz = s[v].z - mu, dw = s[v].w - mu;
        var += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rs = rsqrtf(warp_sum(var) * inv_c + eps);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 4;
      if (c < C) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), bt = __ldg(reinterpret_cast<const float4*>(beta + c));
        st_f4(y + row * C + c, make_float4((s[v].x - mu) * rs * g.x + bt.x, (s[v].y - mu) * rs * g.y + bt.y,
                                           (s[v].z - mu) * rs * g.z + bt.z, (s[v].w - mu) * rs * g.w + bt.w));
      }
    }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (a + b - mean) * rstd
// partial[blockIdx.x][0][c] = sum over the CTA's rows of dy * xhat (dgamma), [1][c] = sum of dy (dbeta)
template <typename T, int NV>
__global__ void __launch_bounds__(kLnThreads)
add_layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ a, const T* __restrict__ b,
                         const float* __restrict__ gamma, const float* __restrict__ mean,
                         const float* __restrict__ rstd, T* __restrict__ dx, float* __restrict__ partial,
                         int64_t rows, int C) {
  extern __shared__ __align__(16) float red[];          // [warps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * (kLnThreads / 32) + warp;
  const int64_t stride = (int64_t)gridDim.x * (kLnThreads / 32);
  const float inv_c = 1.0f / (float)C;
  float4 dg[NV], db[NV], gm[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * 4;
    dg[v] = db[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[v] = (c < C) ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = warp0; row < rows; row += stride) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], g[NV];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 4;
      xh[v] = g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) {
        const float4 x = ld_f4(a + row * C + c), r = ld_f4(b + row * C + c), d = ld_f4(dy + row * C + c);
        xh[v] = make_float4((x.x + r.x - mu) * rs, (x.y + r.y - mu) * rs, (x.z + r.z - mu) * rs, (x.w + r.w - mu) * rs);
        g[v] = make_float4(d.x * gm[v].x, d.y * gm[v].y, d.z * gm[v].z, d.w * gm[v].w);
        s1 += (g[v].x + g[v].y) + (g[v].z + g[v].w);
        s2 += (g[v].x * xh[v].x + g[v].y * xh[v].y) + (g[v].z * xh[v].z + g[v].w * xh[v].w);
        dg[v].x += d.x * xh[v].x; dg[v].y += d.y * xh[v].y; dg[v].z += d.z * xh[v].z; dg[v].w += d.w * xh[v].w;
        db[v].x += d.x; db[v].y += d.y; db[v].z += d.z; db[v].w += d.w;
      }
    }
    const float m1 = warp_sum(s1) * inv_c, m2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 4;
      if (c < C)
        st_f4(dx + row * C + c, make_float4(rs * (g[v].x - m1 - xh[v].x * m2), rs * (g[v].y - m1 - xh[v].y * m2),
                                            rs * (g[v].z - m1 - xh[v].z * m2), rs * (g[v].w - m1 - xh[v].w * m2)));
    }
  }
  // CTA-level reduction of dgamma / dbeta in a fixed order (warp 0..7), then one partial row per CTA
  float* mine = red + (size_t)warp * 2 * C;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * 4;
    if (c < C) {
      *reinterpret_cast<float4*>(mine + c) = dg[v];
      *reinterpret_cast<float4*>(mine + C + c) = db[v];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kLnThreads) {
    float acc = 0.0f;
#pragma unroll
    for (int w = 0; w < kLnThreads / 32; ++w) acc += red[(size_t)w * 2 * C + i];
    partial[(size_t)blockIdx.x * 2 * C + i] = acc;
  }
}

__global__ void add_layernorm_param_grad_kernel(const float* __restrict__ partial, int n_partial, int C,
                                                float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  float acc = 0.0f;
  for (int p = 0; p < n_partial; ++p) acc += partial[(size_t)p * 2 * C + i];
  if (i < C) dgamma[i] = acc;
  else dbeta[i - C] = acc;
}

int ln_grid(int64_t rows) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t need = (rows + kLnThreads / 32 - 1) / (kLnThreads / 32);
  const int64_t cap = (int64_t)sms * 8;                  // 8 CTAs of 256 threads per SM
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

int check_ln(int64_t rows, int C, int dtype) {
  if (rows < 0 || C < 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "negative size (rows=%lld C=%d)", (long long)rows, C);
  if (dtype != MSDA_F32 && dtype != MSDA_BF16) return fail(MSDA_ERR_UNSUPPORTED, "add_layernorm: float32 / bfloat16 only");
  if (C % 4 != 0 || C > 128 * kLnMaxVec)
    return fail(MSDA_ERR_UNSUPPORTED, "add_layernorm: channels must be a multiple of 4 and <= %d (got %d)", 128 * kLnMaxVec, C);
  return MSDA_OK;
}

}  // namespace

extern "C" {

size_t msda_add_layernorm_workspace_bytes(int64_t rows, int channels) {
  if (rows <= 0 || channels <= 0) return 0;
  return (size_t)ln_grid(rows) * 2 * (size_t)channels * sizeof(float);
}

int msda_add_layernorm_forward(void* stream, const void* a, const void* b, const float* gamma, const float* beta,
                               int64_t rows, int channels, float eps, void* y, float* mean, float* rstd, int dtype) {
  if (int s = check_ln(rows, channels, dtype)) return s;
  if (rows == 0 || channels == 0) return MSDA_OK;
  if (!a || !b || !gamma || !beta || !y || !mean || !rstd) return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  if (int s = check_alignment({{"a", a}, {"b", b}, {"gamma", gamma}, {"beta", beta}, {"y", y}})) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = ln_grid(rows);
  const int nv = (channels + 127) / 128;
#define LN_FWD(T_, NV_)                                                                                              \
  add_layernorm_fwd_kernel<T_, NV_><<<grid, kLnThreads, 0, st>>>((const T_*)a, (const T_*)b, gamma, beta, (T_*)y, mean, \
                                                                 rstd, rows, channels, eps)
#define LN_FWD_NV(T_)                                   \
  switch (nv) {                                         \
    case 1: LN_FWD(T_, 1); break;                       \
    case 2: LN_FWD(T_, 2); break;                       \
    case 3: case 4: LN_FWD(T_, 4); break;               \
    default: LN_FWD(T_, 8); break;                      \
  }
  if (dtype == MSDA_F32) { LN_FWD_NV(float) } else { LN_FWD_NV(__nv_bfloat16) }
#undef LN_FWD_NV
#undef LN_FWD
  count_launch();
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

int msda_add_layernorm_backward(void* stream, const void* grad_y, const void* a, const void* b, const float* gamma,
                                const float* mean, const float* rstd, int64_t rows, int channels, void* grad_x,
                                float* grad_gamma, float* grad_beta, void* workspace, size_t workspace_bytes,
                                int dtype) {
  if (int s = check_ln(rows, channels, dtype)) return s;
  if (channels == 0) return MSDA_OK;
  if (!grad_gamma || !grad_beta) return fail(MSDA_ERR_INVALID_ARGUMENT, "grad_gamma / grad_beta is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    MSDA_CUDA(cudaMemsetAsync(grad_gamma, 0, (size_t)channels * sizeof(float), st));
    MSDA_CUDA(cudaMemsetAsync(grad_beta, 0, (size_t)channels * sizeof(float), st));
    return MSDA_OK;
  }
  if (!grad_y || !a || !b || !gamma || !mean || !rstd || !grad_x) return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  if (int s = check_alignment({{"grad_y", grad_y}, {"a", a}, {"b", b}, {"gamma", gamma}, {"grad_x", grad_x},
                               {"workspace", workspace}}))
    return s;
  const size_t need = msda_add_layernorm_workspace_bytes(rows, channels);
  if (!workspace || workspace_bytes < need)
    return fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", need, workspace_bytes);
  const int grid = ln_grid(rows);
  const int nv = (channels + 127) / 128;
  const size_t smem = (size_t)(kLnThreads / 32) * 2 * channels * sizeof(float);
  float* partial = static_cast<float*>(workspace);
#define LN_BWD(T_, NV_)                                                                                     \
  do {                                                                                                      \
    auto k = add_layernorm_bwd_kernel<T_, NV_>;                                                             \
    MSDA_CUDA(ensure_smem(k, smem));                                                                        \
    k<<<grid, kLnThreads, smem, st>>>((const T_*)grad_y, (const T_*)a, (const T_*)b, gamma, mean, rstd,     \
                                      (T_*)grad_x, partial, rows, channels);                                \
  } while (0)
#define LN_BWD_NV(T_)                                   \
  switch (nv) {                                         \
    case 1: LN_BWD(T_, 1); break;                       \
    case 2: LN_BWD(T_, 2); break;                       \
    case 3: case 4: LN_BWD(T_, 4); break;               \
    default: LN_BWD(T_, 8); break;                      \
  }
  if (dtype == MSDA_F32) { LN_BWD_NV(float) } else { LN_BWD_NV(__nv_bfloat16) }
#undef LN_BWD_NV
#undef LN_BWD
  count_launch();
  MSDA_CUDA(cudaGetLastError());
  add_layernorm_param_grad_kernel<<<(2 * channels + 255) / 256, 256, 0, st>>>(partial, grid, channels, grad_gamma, grad_beta);
  count_launch();
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

}  // extern "C"

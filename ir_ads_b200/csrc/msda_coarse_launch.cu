// msda_coarse_launch.cu -- instantiates and launches msda_bwd_coarse_kernel (msda_coarse.cuh).
// A translation unit of its own so that the kernel can be rebuilt without the fast-kernel template zoo of
// msda_capi.cu; msda_capi.cu owns the policy (coarse_plan), the error reporting and the launch counter.
#include "msda_coarse.cuh"

namespace msda {

template <int D, typename VT>
static cudaError_t launch(cudaStream_t st, const void* go, const int64_t* shapes, const int64_t* lsi, const float* loc,
                          const float* w, float* gv, int B, int S, int H, int L, int Q, int P, int budget) {
  auto k = msda_bwd_coarse_kernel<D, VT>;
  // tile + the scratch row of dead records + two staging buffers
  const size_t smem = (size_t)budget + (size_t)D * 4 + 2 * (size_t)coarse_stage_bytes(D, (int)sizeof(VT), L, P);
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  long long grid = ((long long)B * Q * H + 63) / 64;   // at least 64 query rows per CTA
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  k<<<(unsigned)grid, kCoarseThreads, smem, st>>>((const VT*)go, shapes, lsi, loc, w, gv, B, S, H, L, Q, P, budget);
  return cudaGetLastError();
}

// value_is_bf16: grad_out is bfloat16 (else float).  D must be 32, 64 or 128 (cudaErrorInvalidValue otherwise).
cudaError_t launch_bwd_coarse(cudaStream_t st, bool value_is_bf16, const void* go, const int64_t* shapes,
                              const int64_t* lsi, const float* loc, const float* w, float* gv, int B, int S, int H,
                              int D, int L, int Q, int P, int budget) {
#define CALL_C(D_)                                                                                   \
  (value_is_bf16 ? launch<D_, __nv_bfloat16>(st, go, shapes, lsi, loc, w, gv, B, S, H, L, Q, P, budget) \
                 : launch<D_, float>(st, go, shapes, lsi, loc, w, gv, B, S, H, L, Q, P, budget))
  switch (D) {
    case 32: return CALL_C(32);
    case 64: return CALL_C(64);
    case 128: return CALL_C(128);
    default: return cudaErrorInvalidValue;
  }
#undef CALL_C
}

}  // namespace msda

// ref_cuda_shim.cu -- C-ABI shim around the REFERENCE's own CUDA kernels.
//
// TEST / BASELINE INFRASTRUCTURE ONLY (never imported by ir_ads_b200/).
//
// The reference's launchers ms_deformable_im2col_cuda<T> / ms_deformable_col2im_cuda<T>
// (/root/reference/detrex/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:923-954, :956-1327)
// are header templates over raw device pointers and a cudaStream_t, so they can be compiled
// for sm_100a exactly where they lie, unmodified: this file only includes the header by path
// (oracle/Makefile passes -I/root/reference/detrex/layers/csrc/MsDeformAttn) and instantiates
// float and double.  The ATen-level wrapper (ms_deform_attn_cuda.cu) is not used, which also
// sidesteps its torch-1.x `value.type()` dispatch that no longer compiles against torch 2.11.
//
// Output: oracle/_ref/libmsda_refcuda.so (git-ignored; shipped to the GPU box by gpurun).
// Used as (1) a second, independent GPU-side checker and (2) the "reference kernels on the
// same B200" baseline that bench.py reports next to the new kernels.
#include "ms_deform_im2col_cuda.cuh"

#include <cuda_runtime.h>
#include <cstdint>

extern "C" {

// dtype: 0 = float, 1 = double.  Outputs must be zero-filled by the caller exactly as the
// reference's ATen wrapper does (ms_deform_attn_cuda.cu:55,122-124).
int msda_ref_forward(int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
                     const void* loc, const void* w, int B, int S, int H, int D, int L, int Q,
                     int P, void* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == 0)
    ms_deformable_im2col_cuda<float>(st, (const float*)value, shapes, lsi, (const float*)loc,
                                     (const float*)w, B, S, H, D, L, Q, P, (float*)out);
  else
    ms_deformable_im2col_cuda<double>(st, (const double*)value, shapes, lsi, (const double*)loc,
                                      (const double*)w, B, S, H, D, L, Q, P, (double*)out);
  return (int)cudaGetLastError();
}

int msda_ref_backward(int dtype, const void* grad_out, const void* value, const int64_t* shapes,
                      const int64_t* lsi, const void* loc, const void* w, int B, int S, int H,
                      int D, int L, int Q, int P, void* grad_value, void* grad_loc, void* grad_w,
                      void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == 0)
    ms_deformable_col2im_cuda<float>(st, (const float*)grad_out, (const float*)value, shapes, lsi,
                                     (const float*)loc, (const float*)w, B, S, H, D, L, Q, P,
                                     (float*)grad_value, (float*)grad_loc, (float*)grad_w);
  else
    ms_deformable_col2im_cuda<double>(st, (const double*)grad_out, (const double*)value, shapes,
                                      lsi, (const double*)loc, (const double*)w, B, S, H, D, L, Q,
                                      P, (double*)grad_value, (double*)grad_loc, (double*)grad_w);
  return (int)cudaGetLastError();
}

}  // extern "C"

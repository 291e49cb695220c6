// ref_cuda_dcnv3_shim.cu -- C-ABI shim around the REFERENCE's own DCNv3 CUDA kernels (TEST / BASELINE ONLY).
// dcnv3_im2col_cuda<T> / dcnv3_col2im_cuda<T>
// (/root/reference/detrex/layers/csrc/DCNv3/dcnv3_im2col_cuda.cuh:841-868, :870-) are header templates over raw
// device pointers, so they compile for sm_100a where they lie, unmodified (oracle/Makefile adds the include path).
// Output: oracle/_ref/libdcnv3_refcuda.so (git-ignored, shipped to the GPU box).
#include "dcnv3_im2col_cuda.cuh"

#include <cuda_runtime.h>

extern "C" {

// float only.  grad_* must be zero-filled by the caller as the reference's ATen wrapper does.
int dcnv3_ref_forward(const float* input, const float* offset, const float* mask, float* out, int kernel_h,
                      int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w, int dilation_h, int dilation_w,
                      int group, int group_channels, int batch, int height_in, int width_in, int height_out,
                      int width_out, float offset_scale, void* stream) {
  dcnv3_im2col_cuda<float>(static_cast<cudaStream_t>(stream), input, offset, mask, out, kernel_h, kernel_w, stride_h,
                           stride_w, pad_h, pad_w, dilation_h, dilation_w, group, group_channels, batch, height_in,
                           width_in, height_out, width_out, offset_scale);
  return (int)cudaGetLastError();
}

int dcnv3_ref_backward(const float* grad_out, const float* input, const float* offset, const float* mask, int kernel_h,
                       int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w, int dilation_h, int dilation_w,
                       int group, int group_channels, int batch, int height_in, int width_in, int height_out,
                       int width_out, float offset_scale, float* grad_input, float* grad_offset, float* grad_mask,
                       void* stream) {
  dcnv3_col2im_cuda<float>(static_cast<cudaStream_t>(stream), grad_out, input, offset, mask, kernel_h, kernel_w,
                           stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group, group_channels, batch,
                           height_in, width_in, height_out, width_out, offset_scale, grad_input, grad_offset, grad_mask);
  return (int)cudaGetLastError();
}

}  // extern "C"

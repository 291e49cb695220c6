// msda_host.h -- internal host-side interface between the translation units of libmsda_b200.so.
// (The public C ABI is include/msda.h; nothing here is exported.)
//
// The library is split so that the kernel families compile in parallel:
//   msda_capi.cu     the extern "C" entry points, argument checks, generic + deterministic kernels
//   msda_fwd.cu      fast forward kernels (plain / fused module chain / DCNv3)
//   msda_bwd.cu      fast backward kernels, float reds (plain operator, all row orders)
//   msda_bwd_aux.cu  fast backward kernels: fixed-point reds, no-scatter, fused module chain, DCNv3
//   msda_fold.cu     encoder-form backward with on-SM folding of grad_value (msda_fold.cuh)
#pragma once
#include "../../include/msda.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <initializer_list>

#include "msda_fast.cuh"

namespace msda_host {

struct Dims {
  int B, S, H, D, L, Q, P;
  int64_t rows() const { return (int64_t)B * Q * H; }
  int64_t n_value() const { return (int64_t)B * S * H * D; }
  int64_t n_points() const { return rows() * L * P; }
};

// sets the calling thread's last-error message and returns `status`
int fail(int status, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int cuda_fail(cudaError_t e, const char* what);
void count_launch();

#define MSDA_CUDA(call)                                               \
  do {                                                                \
    cudaError_t e__ = (call);                                         \
    if (e__ != cudaSuccess) return ::msda_host::cuda_fail(e__, #call); \
  } while (0)

// The kernels read and write rows with 128-bit accesses (LDG.E.128, REDG...F32x4, STG.E.128): every non-null tensor
// pointer of a compute call must be 16-byte aligned (any torch allocation is; an odd storage offset is not).
// Returns MSDA_OK or MSDA_ERR_INVALID_ARGUMENT naming the first offender.
struct NamedPtr {
  const char* name;
  const void* ptr;
};
int check_alignment(std::initializer_list<NamedPtr> ptrs);

// D in {16,32,64,128}, float / bf16 value, L <= 16, 1 <= L*P <= 64, 32-bit offsets inside one image
bool fast_ok(const Dims& d, int dtype, unsigned flags);

inline int grid_for(int64_t work_items, int threads, int cap_blocks) {
  int64_t g = (work_items + threads - 1) / threads;
  if (g > cap_blocks) g = cap_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename K>
cudaError_t ensure_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// ---- msda_fwd.cu ----
int fwd_fast(cudaStream_t st, const Dims& d, int dtype, unsigned flags, const void* value, const int64_t* shapes,
             const int64_t* lsi, const void* loc, const void* w, void* out);
int fwd_fused(cudaStream_t st, const Dims& d, int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
              const void* off, const void* logits, void* out, msda::FusedArgs fa);
int fwd_dcn(cudaStream_t st, const Dims& d, int dtype, const void* input, const void* off, const void* mask, void* out,
            msda::FusedArgs fa);

// ---- msda_bwd.cu ----
int bwd_fast(cudaStream_t st, const Dims& d, int dtype, unsigned flags, const void* go, const void* value,
             const int64_t* shapes, const int64_t* lsi, const void* loc, const void* w, float* gv, void* gl, void* gw);

// ---- msda_bwd_aux.cu ----
int bwd_fast_det(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
                 const int64_t* lsi, const void* loc, const void* w, unsigned long long* acc, void* gl, void* gw,
                 const msda::DetScale* det);
int bwd_fast_noscatter(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value,
                       const int64_t* shapes, const int64_t* lsi, const void* loc, const void* w, void* gl, void* gw);
int bwd_fast_emit(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
                  const int64_t* lsi, const void* loc, const void* w, void* gl, void* gw, int* cursor,
                  const int* bin_start, void* entries, float* warp_amax, int64_t* n_warps);
int bwd_fused(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
              const int64_t* lsi, const void* off, const void* logits, float* gv, void* goff, void* glog,
              msda::FusedArgs fa);
int bwd_fused_noscatter(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value,
                        const int64_t* shapes, const int64_t* lsi, const void* off, const void* logits, void* goff,
                        void* glog, msda::FusedArgs fa);
int bwd_dcn(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* input, const void* off,
            const void* mask, float* gi, void* goff, void* gmask, msda::FusedArgs fa);

// ---- msda_fold.cu ----
// does the folding backward cover this problem (encoder form Q == S, shapes it is instantiated for, flags)?
bool fold_applies(const Dims& d, int dtype, unsigned flags);
// fa == nullptr: the plain operator (loc / w given); else the fused module chain
int bwd_fold(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
             const int64_t* lsi, const void* loc, const void* w, float* gv, void* gl, void* gw,
             const msda::FusedArgs* fa);

}  // namespace msda_host

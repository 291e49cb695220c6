// msda_bwd_aux.cu -- the other fast backward kernels (msda_fast.cuh): 64-bit fixed-point reds (deterministic),
// scatter compiled out (grad_value not wanted / computed by msda_det.cuh), fused module chain, DCNv3.
#include "msda_fast_launch.cuh"

namespace msda_host {

// deterministic accumulate: LINEAR order
int bwd_fast_det(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
                 const int64_t* lsi, const void* loc, const void* w, unsigned long long* acc, void* gl, void* gw,
                 const msda::DetScale* det) {
#define MSDA_DISPATCH_ORDER MSDA_ORDER_LINEAR
#define CALL_BWD(D_, VT_, PT_, ORD_) \
  launch_bwd_fast<D_, VT_, PT_, kBwdThreads, ORD_, unsigned long long>(st, d, go, value, shapes, lsi, loc, w, acc, gl, gw, det)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_BWD);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_BWD);
#undef CALL_BWD
#undef MSDA_DISPATCH_ORDER
}

// backward without the scatter (grad_loc / grad_w only), LINEAR order
int bwd_fast_noscatter(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value,
                       const int64_t* shapes, const int64_t* lsi, const void* loc, const void* w, void* gl, void* gw) {
#define MSDA_DISPATCH_ORDER MSDA_ORDER_LINEAR
#define CALL_BWD(D_, VT_, PT_, ORD_)                                                                          \
  launch_bwd_fast<D_, VT_, PT_, kBwdThreads, ORD_, msda::NoScatter>(st, d, go, value, shapes, lsi, loc, w,   \
                                                                    (msda::NoScatter*)nullptr, gl, gw, nullptr)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_BWD);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_BWD);
#undef CALL_BWD
#undef MSDA_DISPATCH_ORDER
}

// backward without the scatter that also files every in-range point under its bilinear cell (deterministic sorted
// path: replaces det_bin_kernel<true> + the plain no-scatter backward), LINEAR order
int bwd_fast_emit(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
                  const int64_t* lsi, const void* loc, const void* w, void* gl, void* gw, int* cursor,
                  const int* bin_start, void* entries, float* warp_amax, int64_t* n_warps) {
  // one CTA per kBwdThreads / (D / 4) rows (LINEAR order), kBwdThreads / 32 warps each: the slots of EmitArgs::warp_amax
  const int64_t rpc = kBwdThreads / (d.D / 4);
  if (n_warps) *n_warps = ((d.rows() + rpc - 1) / rpc) * (kBwdThreads / 32);
  msda::EmitArgs ea{cursor, bin_start, static_cast<int4*>(entries), warp_amax};
#define MSDA_DISPATCH_ORDER MSDA_ORDER_LINEAR
#define CALL_BWD(D_, VT_, PT_, ORD_)                                                                            \
  launch_bwd_fast<D_, VT_, PT_, kBwdThreads, ORD_, msda::EmitEntries>(st, d, go, value, shapes, lsi, loc, w,   \
                                                                      (msda::EmitEntries*)nullptr, gl, gw, nullptr, \
                                                                      msda::FusedArgs{}, ea)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_BWD);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_BWD);
#undef CALL_BWD
#undef MSDA_DISPATCH_ORDER
}

#ifdef MSDA_EXP_SLIM
int bwd_fused(cudaStream_t, const Dims&, int, const void*, const void*, const int64_t*, const int64_t*, const void*,
              const void*, float*, void*, void*, msda::FusedArgs) {
  return fail(MSDA_ERR_UNSUPPORTED, "slim build: no fused kernels");
}
int bwd_fused_noscatter(cudaStream_t, const Dims&, int, const void*, const void*, const int64_t*, const int64_t*,
                        const void*, const void*, void*, void*, msda::FusedArgs) {
  return fail(MSDA_ERR_UNSUPPORTED, "slim build: no fused kernels");
}
int bwd_dcn(cudaStream_t, const Dims&, int, const void*, const void*, const void*, const void*, float*, void*, void*,
            msda::FusedArgs) {
  return fail(MSDA_ERR_UNSUPPORTED, "slim build: no DCNv3 kernels");
}
#else
// fused pre-op chain: STRIP order
int bwd_fused(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
              const int64_t* lsi, const void* off, const void* logits, float* gv, void* goff, void* glog,
              msda::FusedArgs fa) {
#define MSDA_DISPATCH_ORDER MSDA_ORDER_STRIP
#define CALL_FB(D_, VT_, PT_, ORD_)                                                                               \
  launch_bwd_fast<D_, VT_, PT_, kBwdThreads, ORD_, float, msda::kPreFused>(st, d, go, value, shapes, lsi, off, logits, gv, \
                                                                           goff, glog, nullptr, fa)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_FB);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_FB);
#undef CALL_FB
#undef MSDA_DISPATCH_ORDER
}

// fused chain, scatter compiled out: gradients of the raw offsets / logits only (deterministic fused backward)
int bwd_fused_noscatter(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value,
                        const int64_t* shapes, const int64_t* lsi, const void* off, const void* logits, void* goff,
                        void* glog, msda::FusedArgs fa) {
#define MSDA_DISPATCH_ORDER MSDA_ORDER_STRIP
#define CALL_FB(D_, VT_, PT_, ORD_)                                                                              \
  launch_bwd_fast<D_, VT_, PT_, kBwdThreads, ORD_, msda::NoScatter, msda::kPreFused>(                            \
      st, d, go, value, shapes, lsi, off, logits, (msda::NoScatter*)nullptr, goff, glog, nullptr, fa)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_FB);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_FB);
#undef CALL_FB
#undef MSDA_DISPATCH_ORDER
}

// DCNv3: runtime point count, STRIP order
int bwd_dcn(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* input, const void* off,
            const void* mask, float* gi, void* goff, void* gmask, msda::FusedArgs fa) {
#define CALL_DB(D_, VT_)                                                                                           \
  launch_bwd_fast<D_, VT_, 0, kBwdThreads, 2, float, msda::kPreDcn>(st, d, go, input, nullptr, nullptr, off, mask, gi, \
                                                                    goff, gmask, nullptr, fa)
#define DCN_D(VT_)                                                                         \
  switch (d.D) {                                                                           \
    case 16: return CALL_DB(16, VT_);                                                      \
    case 32: return CALL_DB(32, VT_);                                                      \
    case 64: return CALL_DB(64, VT_);                                                      \
    case 128: return CALL_DB(128, VT_);                                                    \
    default: return fail(MSDA_ERR_UNSUPPORTED, "dcnv3: group_channels=%d", d.D);           \
  }
  if (dtype == MSDA_F32) DCN_D(float)
  DCN_D(__nv_bfloat16)
#undef DCN_D
#undef CALL_DB
}
#endif

}  // namespace msda_host

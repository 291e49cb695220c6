// msda_fwd.cu -- instantiations and dispatch of the fast forward kernels (msda_fast.cuh).
#include "msda_fast_launch.cuh"

namespace msda_host {

int fwd_fast(cudaStream_t st, const Dims& d, int dtype, unsigned flags, const void* value, const int64_t* shapes,
             const int64_t* lsi, const void* loc, const void* w, void* out) {
#ifdef MSDA_ALL_ORDERS
#define MSDA_DISPATCH_ORDER MSDA_ORDER_ANY
#else
#define MSDA_DISPATCH_ORDER MSDA_ORDER_LINEAR
  (void)flags;
#endif
#define CALL_FWD(D_, VT_, PT_, ORD_) launch_fwd_fast<D_, VT_, PT_, kFwdThreads, ORD_>(st, d, value, shapes, lsi, loc, w, out)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_FWD);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_FWD);
#undef CALL_FWD
#undef MSDA_DISPATCH_ORDER
}

#ifdef MSDA_EXP_SLIM
int fwd_fused(cudaStream_t, const Dims&, int, const void*, const int64_t*, const int64_t*, const void*, const void*,
              void*, msda::FusedArgs) {
  return fail(MSDA_ERR_UNSUPPORTED, "slim build: no fused kernels");
}
int fwd_dcn(cudaStream_t, const Dims&, int, const void*, const void*, const void*, void*, msda::FusedArgs) {
  return fail(MSDA_ERR_UNSUPPORTED, "slim build: no DCNv3 kernels");
}
#else
// fused pre-op chain: LINEAR order
int fwd_fused(cudaStream_t st, const Dims& d, int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
              const void* off, const void* logits, void* out, msda::FusedArgs fa) {
#define MSDA_DISPATCH_ORDER MSDA_ORDER_LINEAR
#define CALL_FF(D_, VT_, PT_, ORD_) \
  launch_fwd_fast<D_, VT_, PT_, kFwdThreads, ORD_, msda::kPreFused>(st, d, value, shapes, lsi, off, logits, out, fa)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_FF);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_FF);
#undef CALL_FF
#undef MSDA_DISPATCH_ORDER
}

// DCNv3: runtime point count (K = kernel_h * kernel_w), LINEAR order
int fwd_dcn(cudaStream_t st, const Dims& d, int dtype, const void* input, const void* off, const void* mask, void* out,
            msda::FusedArgs fa) {
#define CALL_DF(D_, VT_) \
  launch_fwd_fast<D_, VT_, 0, kFwdThreads, 0, msda::kPreDcn>(st, d, input, nullptr, nullptr, off, mask, out, fa)
#define DCN_D(VT_)                                                                         \
  switch (d.D) {                                                                           \
    case 16: return CALL_DF(16, VT_);                                                      \
    case 32: return CALL_DF(32, VT_);                                                      \
    case 64: return CALL_DF(64, VT_);                                                      \
    case 128: return CALL_DF(128, VT_);                                                    \
    default: return fail(MSDA_ERR_UNSUPPORTED, "dcnv3: group_channels=%d", d.D);           \
  }
  if (dtype == MSDA_F32) DCN_D(float)
  DCN_D(__nv_bfloat16)
#undef DCN_D
#undef CALL_DF
}
#endif

}  // namespace msda_host

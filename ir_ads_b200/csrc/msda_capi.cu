// msda_capi.cu -- host side of libmsda_b200.so: argument checks, kernel selection, launches.
// The C ABI is declared and documented in include/msda.h.
//
// Plays the role of ms_deform_attn_cuda_forward / _backward
// (/root/reference/detrex/layers/csrc/MsDeformAttn/ms_deform_attn_cuda.cu:21-154) and of the
// kernel selector ms_deformable_col2im_cuda (ms_deform_im2col_cuda.cuh:956-1327), minus ATen:
// no allocation, no im2col_step chunk loop (one launch covers the batch), errors returned.
#include "../../include/msda.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "msda_fast.cuh"
#include "msda_generic.cuh"
#include "msda_det.cuh"

namespace msda {
// msda_coarse_launch.cu (kernel: msda_coarse.cuh)
constexpr int kCoarseBatch = 32;
constexpr int coarse_stage_bytes(int D, int elem_size, int L, int P) {
  return kCoarseBatch * (D * elem_size + (L < kCoarseMaxLevels ? L : kCoarseMaxLevels) * P * 12);
}
cudaError_t launch_bwd_coarse(cudaStream_t st, bool value_is_bf16, const void* go, const int64_t* shapes,
                              const int64_t* lsi, const float* loc, const float* w, float* gv, int B, int S, int H,
                              int D, int L, int Q, int P, int budget);
}  // namespace msda

namespace {

std::atomic<uint64_t> g_launches{0};
thread_local char g_err[512] = "";

int fail(int status, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

int cuda_fail(cudaError_t e, const char* what) {
  cudaGetLastError();   // clear the runtime's sticky last-error so it cannot be blamed on a later, healthy launch
  return fail(MSDA_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

#define MSDA_CUDA(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// Runs the call on the device that owns `ptr`, restoring the caller's device afterwards
// (the reference has no guard at all: ms_deform_attn_cuda.cu:66 just takes the current stream).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t enter(const void* ptr) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) return e;
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) return cudaErrorInvalidDevicePointer;
    e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return e;
    if (attr.device != prev) {
      e = cudaSetDevice(attr.device);
      switched = (e == cudaSuccess);
    }
    return e;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// Experiment knobs, compiled in only with -DMSDA_EXPERIMENTS (tools/ablate.sh builds such variants into
// build/variants/, never the product library; results with MSDA_EXP_SKIP_COARSE_KERNEL are WRONG on purpose):
//   MSDA_EXP_BWD_SMEM_PAD=<bytes>   extra dynamic shared memory per CTA of the main backward kernel (occupancy study)
//   MSDA_EXP_BWD_CARVEOUT=<percent> preferred shared-memory carveout of the main backward kernel (L1 size study)
//   MSDA_EXP_SKIP_COARSE_KERNEL=1   plan the coarse-level split but do not launch the coarse kernel (times the rest)
#ifdef MSDA_EXPERIMENTS
int exp_env(const char* name) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : 0;
}
#else
constexpr int exp_env(const char*) { return 0; }
#endif

struct Dims {
  int B, S, H, D, L, Q, P;
  int64_t rows() const { return (int64_t)B * Q * H; }
  int64_t n_value() const { return (int64_t)B * S * H * D; }
  int64_t n_points() const { return rows() * L * P; }
};

int check_dims(const Dims& d, int dtype) {
  if (d.B < 0 || d.S < 0 || d.H < 0 || d.D < 0 || d.L < 0 || d.Q < 0 || d.P < 0)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "negative dimension (B=%d S=%d H=%d D=%d L=%d Q=%d P=%d)", d.B, d.S, d.H,
                d.D, d.L, d.Q, d.P);
  if (dtype != MSDA_F32 && dtype != MSDA_F64 && dtype != MSDA_BF16)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "unknown dtype tag %d", dtype);
  return MSDA_OK;
}

bool fast_ok(const Dims& d, int dtype, unsigned flags) {
  if (flags & MSDA_FLAG_FORCE_GENERIC) return false;
  if (dtype != MSDA_F32 && dtype != MSDA_BF16) return false;
  if (!(d.D == 16 || d.D == 32 || d.D == 64 || d.D == 128)) return false;
  if (d.L < 1 || d.L > msda::kFastMaxLevels) return false;
  if (d.L * d.P < 1 || d.L * d.P > msda::kFastMaxPoints) return false;
  // element offsets inside one image are 32-bit in the fast kernels (one spare row/pixel of slack)
  if (((int64_t)d.S + 65536) * d.H * d.D >= ((int64_t)1 << 31)) return false;
  return true;
}

int grid_for(int64_t work_items, int threads, int cap_blocks) {
  int64_t g = (work_items + threads - 1) / threads;
  if (g > cap_blocks) g = cap_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------
// fast-kernel launch helpers
// ------------------------------------------------------------------------------------------
template <typename K>
cudaError_t ensure_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Persistent (TILED) kernels run one wave: SM count x resident CTAs per SM.
template <typename K>
cudaError_t persistent_grid(K kernel, int threads, size_t smem, unsigned* grid) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  *grid = (unsigned)(sms * per_sm);
  return cudaSuccess;
}

// Row order: TILED needs query i == pixel i of the pyramid (encoder self-attention, Q == S).
bool use_tiled(const Dims& d, unsigned flags) {
  if (!(d.Q == d.S && (flags & MSDA_FLAG_ORDER_TILED) && !(flags & MSDA_FLAG_ORDER_LINEAR))) return false;
  // the persistent 1024-thread CTAs stage records + raw loc/w for 4096/D rows; the row order is only a
  // scheduling choice, so fall back to the default order when that does not fit in shared memory
  const int NP = d.L * d.P, rpc = 1024 / (d.D / 8 > 0 ? d.D / 8 : 1);   // worst case: 8 channels per lane
  const size_t words = (size_t)rpc * (size_t)(msda::bwd_row_words(NP, true) > msda::fwd_row_words(NP, true)
                                                  ? msda::bwd_row_words(NP, true)
                                                  : msda::fwd_row_words(NP, true));
  return sizeof(msda::LevelTab) + words * 4 <= 200 * 1024;
}
bool use_strip(const Dims& d, unsigned flags) {
  (void)d;
  return (flags & MSDA_FLAG_ORDER_STRIP) && !(flags & (MSDA_FLAG_ORDER_LINEAR | MSDA_FLAG_ORDER_TILED | MSDA_FLAG_ORDER_TILE2D));
}
bool use_tile2d(const Dims& d, unsigned flags) {
  return d.Q == d.S && (flags & MSDA_FLAG_ORDER_TILE2D) && !(flags & (MSDA_FLAG_ORDER_LINEAR | MSDA_FLAG_ORDER_TILED));
}
// TILE2D launches an upper bound on the tile count that needs only S and L (the level shapes stay on the
// device): sum_l ceil(H_l/TH)*ceil(W_l/TW) is ~1.03 * S/RPC for image pyramids; 25 % + 32 tiles per level of
// slack covers them, the kernel's grid-stride step covers anything else.
int64_t tile2d_bound(const Dims& d, int rpc) { return ((int64_t)d.S + rpc - 1) / rpc * 5 / 4 + 32 * (int64_t)d.L; }

// CTA sizes of the single-pass row orders (LINEAR, STRIP).  An SM re-uses a CTA's slot only when the CTA's slowest
// warp is done, so small CTAs keep more warps busy: cfg 2 forward 0.665 / 0.635 / 0.620 ms at 256 / 128 / 64
// threads, backward 1.687 / 1.672 / 1.695 ms (profiles/r01s_experiments.txt).  The tile orders keep their own.
#ifndef MSDA_FWD_THREADS
#define MSDA_FWD_THREADS 64
#endif
#ifndef MSDA_BWD_THREADS
#define MSDA_BWD_THREADS 128
#endif
constexpr int kFwdThreads = MSDA_FWD_THREADS, kBwdThreads = MSDA_BWD_THREADS;
constexpr int fwd_threads(int threads, int order) { return (order == 0 || order == 2) ? kFwdThreads : threads; }
constexpr int bwd_threads(int threads, int order) { return (order == 0 || order == 2) ? kBwdThreads : threads; }

// channels per lane: 8 for bf16 rows of 32+ channels (16-byte lane loads), else 4
template <int D, typename VT>
constexpr int cpl_of() { return (sizeof(VT) == 2 && D >= 32) ? 8 : 4; }

template <int D, typename VT, int PT, int THREADS, int TILED, int PRE = 0>
int launch_fwd_fast(cudaStream_t st, const Dims& d, const void* value, const int64_t* shapes, const int64_t* lsi,
                    const void* loc, const void* w, void* out, msda::FusedArgs fa = msda::FusedArgs{},
                    int head_major = 0) {
  constexpr int CPL = cpl_of<D, VT>();
  using G = msda::Geom<D * 4 / CPL, THREADS>;
  const int NP = d.L * d.P;
  const size_t smem = sizeof(msda::LevelTab) + (size_t)G::RPC * msda::fwd_row_words(NP, TILED == 1) * 4;
  auto k = msda::msda_fwd_fast_kernel<D, VT, PT, THREADS, TILED, PRE, CPL>;
  MSDA_CUDA(ensure_smem(k, smem));
  const int64_t rows = d.rows();
  unsigned grid = (unsigned)((rows + G::RPC - 1) / G::RPC);
  if (TILED == 1) MSDA_CUDA(persistent_grid(k, THREADS, smem, &grid));
  if (TILED == 2) grid = (unsigned)((int64_t)d.B * d.H * ((d.Q + G::RPC - 1) / G::RPC));
  if (TILED == 3) grid = (unsigned)((int64_t)d.B * d.H * tile2d_bound(d, G::RPC));
  k<<<grid, THREADS, smem, st>>>((const VT*)value, shapes, lsi, (const float*)loc, (const float*)w, (VT*)out, fa,
                                 d.B, d.S, d.H, d.L, d.Q, d.P, rows, head_major);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

template <int D, typename VT, int PT, int THREADS, int TILED, typename ACC, int PRE = 0>
int launch_bwd_fast(cudaStream_t st, const Dims& d, const void* go, const void* value, const int64_t* shapes,
                    const int64_t* lsi, const void* loc, const void* w, ACC* gv, void* gl, void* gw,
                    const msda::DetScale* det, msda::FusedArgs fa = msda::FusedArgs{}, int coarse_budget = 0,
                    int head_major = 0) {
  // The backward keeps 4 channels per lane for every type: it is bound by the grad_value reds, and those run
  // fastest as one full 128-byte line per row and instruction (8 channels per lane -> two 64-byte halves per
  // row: 1.75 -> 2.09 ms at cfg2 with bf16 value), so the faster 16-byte gather buys nothing there.
  constexpr int CPL = 4;
  using G = msda::Geom<D * 4 / CPL, THREADS>;
  const int NP = d.L * d.P;
  static const int exp_pad = exp_env("MSDA_EXP_BWD_SMEM_PAD");
  const size_t smem = sizeof(msda::LevelTab) + (size_t)G::RPC * msda::bwd_row_words(NP, TILED == 1) * 4 + (size_t)exp_pad;
  auto k = msda::msda_bwd_fast_kernel<D, VT, PT, THREADS, TILED, ACC, PRE, CPL>;
  MSDA_CUDA(ensure_smem(k, smem));
  static const int exp_carve = exp_env("MSDA_EXP_BWD_CARVEOUT");   // percent of the SM's 228 KB given to shared memory
  if (exp_carve > 0) MSDA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, exp_carve));
  const int64_t rows = d.rows();
  unsigned grid = (unsigned)((rows + G::RPC - 1) / G::RPC);
  if (TILED == 1) MSDA_CUDA(persistent_grid(k, THREADS, smem, &grid));
  if (TILED == 2) grid = (unsigned)((int64_t)d.B * d.H * ((d.Q + G::RPC - 1) / G::RPC));
  if (TILED == 3) grid = (unsigned)((int64_t)d.B * d.H * tile2d_bound(d, G::RPC));
  k<<<grid, THREADS, smem, st>>>((const VT*)go, (const VT*)value, shapes, lsi, (const float*)loc, (const float*)w,
                                 gv, (float*)gl, (float*)gw, det, fa, d.B, d.S, d.H, d.L, d.Q, d.P, rows, coarse_budget,
                                 head_major);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

#ifdef MSDA_EXP_SLIM
// Kernel-variant experiment builds (tools/build_variant.sh): only D = 32, P in {4, 8}, LINEAR / STRIP orders are instantiated,
// so a variant compiles in well under a minute.  Never defined for the product library.
#define MSDA_DISPATCH_ORDER(D_, VT_, PT_, CALL)                                      \
  do {                                                                               \
    if (use_strip(d, flags)) return CALL(D_, VT_, PT_, 256, 2);                      \
    return CALL(D_, VT_, PT_, 256, 0);                                               \
  } while (0)
#define MSDA_DISPATCH_PT(D_, VT_, CALL)                      \
  do {                                                       \
    if (d.P == 4) MSDA_DISPATCH_ORDER(D_, VT_, 4, CALL);     \
    if (d.P == 8) MSDA_DISPATCH_ORDER(D_, VT_, 8, CALL);     \
    return fail(MSDA_ERR_UNSUPPORTED, "slim build: P=%d", d.P); \
  } while (0)
#define MSDA_DISPATCH_D(VT_, CALL)                           \
  do {                                                       \
    switch (d.D) {                                           \
      case 32: MSDA_DISPATCH_PT(32, VT_, CALL);              \
      default: return fail(MSDA_ERR_UNSUPPORTED, "slim build: D=%d", d.D); \
    }                                                        \
  } while (0)
#else
#define MSDA_DISPATCH_ORDER(D_, VT_, PT_, CALL)                                      \
  do {                                                                               \
    if (use_tile2d(d, flags)) return CALL(D_, VT_, PT_, 256, 3);                     \
    if (use_strip(d, flags)) return CALL(D_, VT_, PT_, 256, 2);                      \
    if (!use_tiled(d, flags)) return CALL(D_, VT_, PT_, 256, 0);                     \
    return CALL(D_, VT_, PT_, 1024, 1);                                              \
  } while (0)

#define MSDA_DISPATCH_PT(D_, VT_, CALL)                      \
  do {                                                       \
    if (d.P == 4) MSDA_DISPATCH_ORDER(D_, VT_, 4, CALL);     \
    if (d.P == 8) MSDA_DISPATCH_ORDER(D_, VT_, 8, CALL);     \
    MSDA_DISPATCH_ORDER(D_, VT_, 0, CALL);                   \
  } while (0)

#define MSDA_DISPATCH_D(VT_, CALL)                           \
  do {                                                       \
    switch (d.D) {                                           \
      case 16: MSDA_DISPATCH_PT(16, VT_, CALL);              \
      case 32: MSDA_DISPATCH_PT(32, VT_, CALL);              \
      case 64: MSDA_DISPATCH_PT(64, VT_, CALL);              \
      case 128: MSDA_DISPATCH_PT(128, VT_, CALL);            \
      default: return fail(MSDA_ERR_UNSUPPORTED, "fast path: D=%d", d.D); \
    }                                                        \
  } while (0)
#endif

int fwd_fast(cudaStream_t st, const Dims& d, int dtype, unsigned flags, const void* value, const int64_t* shapes,
             const int64_t* lsi, const void* loc, const void* w, void* out) {
#define CALL_FWD(D_, VT_, PT_, TH_, TL_)                                                              \
  launch_fwd_fast<D_, VT_, PT_, fwd_threads(TH_, TL_), TL_>(st, d, value, shapes, lsi, loc, w, out, msda::FusedArgs{}, \
                                          (flags & MSDA_FLAG_STRIP_HEAD_MAJOR) ? 1 : 0)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_FWD);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_FWD);
#undef CALL_FWD
}

int bwd_fast(cudaStream_t st, const Dims& d, int dtype, unsigned flags, const void* go, const void* value,
             const int64_t* shapes, const int64_t* lsi, const void* loc, const void* w, float* gv, void* gl,
             void* gw, int coarse_budget) {
  // default row order of the backward: STRIP (measured 3 % faster than LINEAR at cfg 2: fewer L1 misses
  // on the crossbar-bound kernel); the forward keeps LINEAR
  if (!(flags & (MSDA_FLAG_ORDER_LINEAR | MSDA_FLAG_ORDER_TILED | MSDA_FLAG_ORDER_TILE2D))) flags |= MSDA_FLAG_ORDER_STRIP;
#define CALL_BWD(D_, VT_, PT_, TH_, TL_) \
  launch_bwd_fast<D_, VT_, PT_, bwd_threads(TH_, TL_), TL_, float>(st, d, go, value, shapes, lsi, loc, w, gv, gl, gw, nullptr, \
                                                 msda::FusedArgs{}, coarse_budget,                                \
                                                 (flags & MSDA_FLAG_STRIP_HEAD_MAJOR) ? 1 : 0)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_BWD);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_BWD);
#undef CALL_BWD
}

// ---- shared-memory accumulation of the coarse levels (msda_coarse.cuh) ----
constexpr int kCoarseBudget = 200 * 1024;   // most shared memory the resident grad_value tile may take

// What the coarse-level accumulation gets for this problem.  budget == 0: not used.  Which levels are resident
// is decided on the device from the level shapes and the budget; when the caller supplied a HOST copy of the
// shapes the budget is exactly the resident levels' bytes (so the kernel leaves the rest of the SM's shared
// memory to the main backward kernel and the two can share every SM), otherwise it is the upper bound and the
// two kernels run one after the other.
struct CoarsePlan {
  int budget = 0;
  bool exact = false;
};
CoarsePlan coarse_plan(const Dims& d, int dtype, unsigned flags, const int64_t* host_shapes) {
  CoarsePlan p;
  if (flags & (MSDA_FLAG_COARSE_OFF | MSDA_FLAG_DETERMINISTIC)) return p;
  if (!fast_ok(d, dtype, flags)) return p;
  if (!(d.D == 32 || d.D == 64 || d.D == 128)) return p;
  // Opt-in: measured SLOWER than the all-reds backward at every benchmark shape: a
  // shared-memory read-modify-write costs the SM's load/store pipe two instructions per 128-byte row where a vector
  // red costs one, and the resident tile takes the L1 capacity the main kernel's gathers live on.
  if (!(flags & MSDA_FLAG_COARSE_ON)) return p;
  const int es = dtype == MSDA_BF16 ? 2 : 4;
  int cap = 227 * 1024 - 1024 - d.D * 4 - 2 * msda::coarse_stage_bytes(d.D, es, d.L, d.P);
  if (cap > kCoarseBudget) cap = kCoarseBudget;
  if (cap < d.D * 4) return p;
  if (host_shapes) {
    int Hs[msda::kFastMaxLevels], Ws[msda::kFastMaxLevels];
    for (int l = 0; l < d.L; ++l) {
      Hs[l] = (int)host_shapes[2 * l];
      Ws[l] = (int)host_shapes[2 * l + 1];
    }
    const int lc = msda::coarse_first_level(Hs, Ws, d.L, d.D, cap);
    int64_t bytes = 0;
    for (int l = lc; l < d.L; ++l) bytes += (int64_t)Hs[l] * Ws[l] * d.D * 4;
    if (bytes == 0) return p;                       // nothing fits
    p.budget = (int)bytes;
    p.exact = true;
    return p;
  }
  const int64_t whole = (int64_t)d.S * d.D * 4;
  p.budget = (int)(whole < cap ? whole : cap);
  return p;
}

int bwd_coarse(cudaStream_t st, const Dims& d, int dtype, const void* go, const int64_t* shapes, const int64_t* lsi,
               const void* loc, const void* w, float* gv, int budget) {
  static const int exp_skip = exp_env("MSDA_EXP_SKIP_COARSE_KERNEL");
  if (exp_skip) return MSDA_OK;
  MSDA_CUDA(msda::launch_bwd_coarse(st, dtype == MSDA_BF16, go, shapes, lsi, (const float*)loc, (const float*)w, gv, d.B,
                                    d.S, d.H, d.D, d.L, d.Q, d.P, budget));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MSDA_OK;
}

// The coarse kernel (1 CTA per SM, most of the SM's shared memory, few warps) and the fast kernel (many small
// CTAs) use different resources, so they run CONCURRENTLY: the coarse kernel goes first on a library-owned
// high-priority side stream, the fast kernel fills the rest of every SM from the caller's stream, and the
// caller's stream then waits for the side stream.  One side stream + event pair per device, created on first
// use; the fork/join is enqueued under a mutex (cudaStreamWaitEvent binds to the record that precedes it, so
// re-using the events afterwards is safe).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
constexpr int kMaxDevices = 64;
SideStream g_side[kMaxDevices];
std::mutex g_side_mutex;

cudaError_t side_stream(SideStream** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  SideStream& s = g_side[dev];
  if (!s.stream) {
    int lo = 0, hi = 0;
    e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e != cudaSuccess) return e;
    e = cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, hi);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  *out = &s;
  return cudaSuccess;
}

// fast backward + coarse accumulation, concurrent unless the caller's stream is being captured into a graph
// (or MSDA_FLAG_COARSE_SERIAL): then both run on the caller's stream, one after the other
int bwd_fast_with_coarse(cudaStream_t st, const Dims& d, int dtype, unsigned flags, const void* go, const void* value,
                         const int64_t* shapes, const int64_t* lsi, const void* loc, const void* w, float* gv, void* gl,
                         void* gw, int budget, bool exact) {
  bool serial = (flags & MSDA_FLAG_COARSE_SERIAL) != 0 || !exact;
  if (!serial) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    MSDA_CUDA(cudaStreamIsCapturing(st, &cs));
    serial = cs != cudaStreamCaptureStatusNone;
  }
  if (serial) {
    if (int s = bwd_fast(st, d, dtype, flags, go, value, shapes, lsi, loc, w, gv, gl, gw, budget)) return s;
    return bwd_coarse(st, d, dtype, go, shapes, lsi, loc, w, gv, budget);
  }
  std::lock_guard<std::mutex> lock(g_side_mutex);
  SideStream* side = nullptr;
  MSDA_CUDA(side_stream(&side));
  MSDA_CUDA(cudaEventRecord(side->fork, st));            // after the caller's zero-fill of grad_value
  MSDA_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
  int s = bwd_coarse(side->stream, d, dtype, go, shapes, lsi, loc, w, gv, budget);
  if (s == MSDA_OK) s = bwd_fast(st, d, dtype, flags, go, value, shapes, lsi, loc, w, gv, gl, gw, budget);
  // join even after a failure so the caller's stream never runs ahead of work already queued on the side stream
  cudaError_t e = cudaEventRecord(side->join, side->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(st, side->join, 0);
  if (s != MSDA_OK) return s;
  if (e != cudaSuccess) return cuda_fail(e, "joining the coarse-level side stream");
  return MSDA_OK;
}

// deterministic accumulate: LINEAR order only
int bwd_fast_det(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
                 const int64_t* lsi, const void* loc, const void* w, unsigned long long* acc, void* gl, void* gw,
                 const msda::DetScale* det) {
  const unsigned flags = MSDA_FLAG_ORDER_LINEAR;
#define CALL_BWD(D_, VT_, PT_, TH_, TL_)                                                                      \
  launch_bwd_fast<D_, VT_, PT_, kBwdThreads, 0, unsigned long long>(st, d, go, value, shapes, lsi, loc, w, acc, gl, \
                                                                gw, det)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_BWD);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_BWD);
#undef CALL_BWD
}

// sorted deterministic path: backward without the scatter (grad_loc / grad_w only), LINEAR order
int bwd_fast_noscatter(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value,
                       const int64_t* shapes, const int64_t* lsi, const void* loc, const void* w, void* gl, void* gw) {
  const unsigned flags = MSDA_FLAG_ORDER_LINEAR;
#define CALL_BWD(D_, VT_, PT_, TH_, TL_)                                                                       \
  launch_bwd_fast<D_, VT_, PT_, kBwdThreads, 0, msda::NoScatter>(st, d, go, value, shapes, lsi, loc, w,                  \
                                                         (msda::NoScatter*)nullptr, gl, gw, nullptr)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D(float, CALL_BWD);
  MSDA_DISPATCH_D(__nv_bfloat16, CALL_BWD);
#undef CALL_BWD
}

template <typename VT>
int det_gather(cudaStream_t st, const Dims& d, const void* go, const int4* entries, const int* bin_start,
               const int64_t* shapes, const int64_t* lsi, const msda::DetScale* scale, void* gv) {
  const int64_t items = (int64_t)d.B * d.S * d.H;
#define CALL_G(D_)                                                                                            \
  do {                                                                                                        \
    constexpr int RPC = 256 / (D_ / 4);                                                                       \
    msda::det_gather_kernel<D_, VT><<<(unsigned)((items + RPC - 1) / RPC), 256, 0, st>>>(                     \
        (const VT*)go, entries, bin_start, shapes, lsi, scale, (VT*)gv, d.B, d.S, d.H, d.L);                  \
  } while (0)
  switch (d.D) {
    case 16: CALL_G(16); break;
    case 32: CALL_G(32); break;
    case 64: CALL_G(64); break;
    case 128: CALL_G(128); break;
    default: return fail(MSDA_ERR_UNSUPPORTED, "det gather: D=%d", d.D);
  }
#undef CALL_G
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

// dense problems reduce per cell (one pass over the entries), sparse ones gather per pixel (msda_det.cuh)
bool det_dense(const Dims& d) { return d.n_points() >= (int64_t)msda::kDetDenseRatio * d.B * d.S * d.H; }

template <typename VT>
int det_cell_reduce(cudaStream_t st, const Dims& d, const void* go, const int4* entries, const int* bin_start,
                    int64_t n_bins, const int64_t* shapes, const int64_t* lsi, const msda::DetScale* scale,
                    unsigned long long* acc) {
  const int64_t slices = (d.n_points() + msda::kDetSlice - 1) / msda::kDetSlice;   // upper bound: gated points have no entry
#define CALL_G(D_)                                                                                            \
  do {                                                                                                        \
    constexpr int GPC = 256 / (D_ / 4);                                                                       \
    msda::det_cell_reduce_kernel<D_, VT><<<(unsigned)((slices + GPC - 1) / GPC), 256, 0, st>>>(               \
        (const VT*)go, entries, bin_start, (int)n_bins, shapes, lsi, scale, acc, d.S, d.H, d.L);              \
  } while (0)
  switch (d.D) {
    case 16: CALL_G(16); break;
    case 32: CALL_G(32); break;
    case 64: CALL_G(64); break;
    case 128: CALL_G(128); break;
    default: return fail(MSDA_ERR_UNSUPPORTED, "det cell reduce: D=%d", d.D);
  }
#undef CALL_G
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

// Workspace of the sorted deterministic path (all 16-byte aligned):
//   [bins+1 ints: counts -> starts][bins ints: cursors][scan block sums][entries: n_points x 16 B][amax x2, DetScale]
//   [fixed-point accumulators: n_value x 8 B]
struct DetLayout {
  int64_t bins;        // upper bound, B*H*(2S+2L)
  int64_t n_scan;      // bins + 1
  int64_t scan_blocks;
  size_t off_cursor, off_sums, off_entries, off_misc, off_acc, total;
};
DetLayout det_layout(const Dims& d) {
  DetLayout l;
  l.bins = (int64_t)d.B * d.H * msda::det_cells_bound(d.S, d.L);
  l.n_scan = l.bins + 1;
  l.scan_blocks = (l.n_scan + msda::kScanPerBlock - 1) / msda::kScanPerBlock;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  l.off_cursor = up((size_t)l.n_scan * 4);
  l.off_sums = l.off_cursor + up((size_t)l.bins * 4);
  l.off_entries = l.off_sums + up((size_t)l.scan_blocks * 4);
  l.off_misc = l.off_entries + up((size_t)d.n_points() * 16);
  l.off_acc = l.off_misc + 256;
  l.total = l.off_acc + (det_dense(d) ? up((size_t)d.n_value() * 8) : 0);   // accumulators: cell reduce only
  return l;
}
// the sorted path needs 32-bit bin / entry / row indices
bool det_sorted_ok(const Dims& d, int dtype, unsigned flags) {
  if (flags & MSDA_FLAG_DET_ATOMIC) return false;
  if (!fast_ok(d, dtype, flags)) return false;
  const int64_t lim = ((int64_t)1 << 31) - 4096;
  return (int64_t)d.B * d.H * msda::det_cells_bound(d.S, d.L) < lim && d.n_points() < lim && d.rows() < lim;
}

// fused pre-op chain: single-pass row orders only (LINEAR forward, STRIP backward)
#define MSDA_DISPATCH_PT_FUSED(D_, VT_, CALL)     \
  do {                                            \
    if (d.P == 4) return CALL(D_, VT_, 4);        \
    if (d.P == 8) return CALL(D_, VT_, 8);        \
    return CALL(D_, VT_, 0);                      \
  } while (0)
#ifdef MSDA_EXP_SLIM
#define MSDA_DISPATCH_D_FUSED(VT_, CALL) return fail(MSDA_ERR_UNSUPPORTED, "slim build: no fused kernels")
#else
#define MSDA_DISPATCH_D_FUSED(VT_, CALL)                     \
  do {                                                       \
    switch (d.D) {                                           \
      case 16: MSDA_DISPATCH_PT_FUSED(16, VT_, CALL);        \
      case 32: MSDA_DISPATCH_PT_FUSED(32, VT_, CALL);        \
      case 64: MSDA_DISPATCH_PT_FUSED(64, VT_, CALL);        \
      case 128: MSDA_DISPATCH_PT_FUSED(128, VT_, CALL);      \
      default: return fail(MSDA_ERR_UNSUPPORTED, "fused path: D=%d", d.D); \
    }                                                        \
  } while (0)
#endif

int fwd_fused(cudaStream_t st, const Dims& d, int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
              const void* off, const void* logits, void* out, msda::FusedArgs fa) {
#define CALL_FF(D_, VT_, PT_) launch_fwd_fast<D_, VT_, PT_, kFwdThreads, 0, msda::kPreFused>(st, d, value, shapes, lsi, off, logits, out, fa)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D_FUSED(float, CALL_FF);
  MSDA_DISPATCH_D_FUSED(__nv_bfloat16, CALL_FF);
#undef CALL_FF
}

int bwd_fused(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* value, const int64_t* shapes,
              const int64_t* lsi, const void* off, const void* logits, float* gv, void* goff, void* glog,
              msda::FusedArgs fa) {
#define CALL_FB(D_, VT_, PT_)                                                                                    \
  launch_bwd_fast<D_, VT_, PT_, kBwdThreads, 2, float, msda::kPreFused>(st, d, go, value, shapes, lsi, off, logits, gv, goff, glog, \
                                                     nullptr, fa)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D_FUSED(float, CALL_FB);
  MSDA_DISPATCH_D_FUSED(__nv_bfloat16, CALL_FB);
#undef CALL_FB
}

// DCNv3: runtime point count (K = kernel_h * kernel_w), LINEAR forward, STRIP backward
#ifdef MSDA_EXP_SLIM
#define MSDA_DISPATCH_D_DCN(VT_, CALL) return fail(MSDA_ERR_UNSUPPORTED, "slim build: no DCNv3 kernels")
#else
#define MSDA_DISPATCH_D_DCN(VT_, CALL)                       \
  do {                                                       \
    switch (d.D) {                                           \
      case 16: return CALL(16, VT_);                         \
      case 32: return CALL(32, VT_);                         \
      case 64: return CALL(64, VT_);                         \
      case 128: return CALL(128, VT_);                       \
      default: return fail(MSDA_ERR_UNSUPPORTED, "dcnv3: group_channels=%d", d.D); \
    }                                                        \
  } while (0)
#endif

int fwd_dcn(cudaStream_t st, const Dims& d, int dtype, const void* input, const void* off, const void* mask, void* out,
            msda::FusedArgs fa) {
#define CALL_DF(D_, VT_) \
  launch_fwd_fast<D_, VT_, 0, kFwdThreads, 0, msda::kPreDcn>(st, d, input, nullptr, nullptr, off, mask, out, fa)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D_DCN(float, CALL_DF);
  MSDA_DISPATCH_D_DCN(__nv_bfloat16, CALL_DF);
#undef CALL_DF
}

int bwd_dcn(cudaStream_t st, const Dims& d, int dtype, const void* go, const void* input, const void* off,
            const void* mask, float* gi, void* goff, void* gmask, msda::FusedArgs fa) {
#define CALL_DB(D_, VT_)                                                                                          \
  launch_bwd_fast<D_, VT_, 0, kBwdThreads, 2, float, msda::kPreDcn>(st, d, go, input, nullptr, nullptr, off, mask, gi, goff, \
                                                            gmask, nullptr, fa)
  if (dtype == MSDA_F32) MSDA_DISPATCH_D_DCN(float, CALL_DB);
  MSDA_DISPATCH_D_DCN(__nv_bfloat16, CALL_DB);
#undef CALL_DB
}

// ------------------------------------------------------------------------------------------
// generic launches
// ------------------------------------------------------------------------------------------
template <typename VT, typename CT>
int fwd_generic(cudaStream_t st, const Dims& d, const void* value, const int64_t* shapes, const int64_t* lsi,
                const void* loc, const void* w, void* out) {
  const int64_t total = d.rows() * d.D;
  const int grid = grid_for(total, 256, 148 * 32);
  msda::msda_fwd_generic_kernel<VT, CT><<<grid, 256, 0, st>>>((const VT*)value, shapes, lsi, (const CT*)loc,
                                                              (const CT*)w, (VT*)out, d.S, d.H, d.D, d.L, d.Q, d.P,
                                                              total);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

template <typename VT, typename CT, typename ACC>
int bwd_generic(cudaStream_t st, const Dims& d, const void* go, const void* value, const int64_t* shapes,
                const int64_t* lsi, const void* loc, const void* w, ACC* gv, void* gl, void* gw,
                const msda::DetScale* det = nullptr) {
  int threads = 32;
  while (threads < d.D && threads < 256) threads <<= 1;
  const int64_t rows = d.rows();
  const int grid = (int)(rows < 148 * 64 ? rows : 148 * 64);
  msda::msda_bwd_generic_kernel<VT, CT, ACC><<<grid, threads, 0, st>>>(
      (const VT*)go, (const VT*)value, shapes, lsi, (const CT*)loc, (const CT*)w, gv, (CT*)gl, (CT*)gw, det, d.S, d.H,
      d.D, d.L, d.Q, d.P, rows);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

size_t elem_size(int dtype) { return dtype == MSDA_F64 ? 8 : (dtype == MSDA_BF16 ? 2 : 4); }

}  // namespace

// ============================================================================================
extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

uint64_t msda_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* msda_last_error_message(void) { return g_err; }

const char* msda_status_string(int status) {
  switch (status) {
    case MSDA_OK: return "MSDA_OK";
    case MSDA_ERR_INVALID_ARGUMENT: return "MSDA_ERR_INVALID_ARGUMENT";
    case MSDA_ERR_UNSUPPORTED: return "MSDA_ERR_UNSUPPORTED";
    case MSDA_ERR_WORKSPACE: return "MSDA_ERR_WORKSPACE";
    case MSDA_ERR_CUDA: return "MSDA_ERR_CUDA";
    default: return "MSDA_ERR_UNKNOWN";
  }
}

const char* msda_dispatch_name(int channels, int num_levels, int num_point, int spatial_size, int num_heads,
                               int dtype, unsigned flags, int backward) {
  Dims d{1, spatial_size, num_heads, channels, num_levels, 1, num_point};
  static thread_local char name[64];
  const char* dt = dtype == MSDA_F64 ? "f64" : (dtype == MSDA_BF16 ? "bf16" : "f32");
  if (fast_ok(d, dtype, flags))
    snprintf(name, sizeof(name), "%s_fast_d%d_%s", backward ? "bwd" : "fwd", channels, dt);
  else
    snprintf(name, sizeof(name), "%s_generic_%s", backward ? "bwd" : "fwd", dt);
  return name;
}

int msda_forward(void* stream, const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                 const void* sampling_loc, const void* attn_weight, int batch, int spatial_size, int num_heads,
                 int channels, int num_levels, int num_query, int num_point, void* output, int dtype,
                 unsigned flags) {
  g_err[0] = 0;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int s = check_dims(d, dtype)) return s;
  if (d.rows() * d.D == 0) return MSDA_OK;  // nothing to write (the reference would launch a 0-block grid, cuh:942)
  if (!output) return fail(MSDA_ERR_INVALID_ARGUMENT, "output is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d.n_points() == 0 || d.S == 0) {  // no samples: the sum over an empty set
    DeviceGuard g;
    MSDA_CUDA(g.enter(output));
    MSDA_CUDA(cudaMemsetAsync(output, 0, (size_t)(d.rows() * d.D) * elem_size(dtype), st));
    return MSDA_OK;
  }
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  DeviceGuard guard;
  MSDA_CUDA(guard.enter(value));
  if (fast_ok(d, dtype, flags))
    return fwd_fast(st, d, dtype, flags, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output);
  switch (dtype) {
    case MSDA_F32:
      return fwd_generic<float, float>(st, d, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output);
    case MSDA_F64:
      return fwd_generic<double, double>(st, d, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output);
    default:
      return fwd_generic<__nv_bfloat16, float>(st, d, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output);
  }
}

size_t msda_backward_workspace_bytes(int batch, int spatial_size, int num_heads, int channels, int num_levels,
                                     int num_query, int num_point, int dtype, unsigned flags) {
  (void)num_levels; (void)num_query; (void)num_point;
  if (batch <= 0 || spatial_size <= 0 || num_heads <= 0 || channels <= 0) return 0;
  const size_t n_value = (size_t)batch * spatial_size * num_heads * channels;
  if (flags & MSDA_FLAG_DETERMINISTIC) {
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    if (det_sorted_ok(d, dtype, flags)) return det_layout(d).total;   // bins, cursors, scan sums, entries
    return n_value * sizeof(long long) + 64;   // 64-bit fixed-point accumulators + {amax bits x2, DetScale}
  }
  // bf16 grad_value is accumulated in float and converted at the end
  if (dtype == MSDA_BF16) return n_value * sizeof(float);
  return 0;
}

int msda_backward(void* stream, const void* grad_output, const void* value, const int64_t* spatial_shapes,
                  const int64_t* level_start_index, const void* sampling_loc, const void* attn_weight, int batch,
                  int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                  void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* workspace,
                  size_t workspace_bytes, int dtype, unsigned flags) {
  return msda_backward_hs(stream, grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                          batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, grad_value,
                          grad_sampling_loc, grad_attn_weight, workspace, workspace_bytes, dtype, flags, nullptr);
}

int msda_backward_hs(void* stream, const void* grad_output, const void* value, const int64_t* spatial_shapes,
                     const int64_t* level_start_index, const void* sampling_loc, const void* attn_weight, int batch,
                     int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                     void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* workspace,
                     size_t workspace_bytes, int dtype, unsigned flags, const int64_t* spatial_shapes_host) {
  g_err[0] = 0;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int s = check_dims(d, dtype)) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t es = elem_size(dtype);
  const size_t ls = dtype == MSDA_F64 ? 8 : 4;
  if (d.n_value() == 0 && d.n_points() == 0) return MSDA_OK;
  if ((flags & MSDA_FLAG_NO_GRAD_VALUE) && d.n_value() != 0 && d.n_points() != 0 && fast_ok(d, dtype, flags)) {
    // grad_value not wanted: the regular kernel with the scatter compiled out (no zero-fill, no workspace)
    if (!grad_output || !value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight ||
        !grad_sampling_loc || !grad_attn_weight)
      return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    DeviceGuard guard_ns;
    MSDA_CUDA(guard_ns.enter(value));
    return bwd_fast_noscatter(st, d, dtype, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                              attn_weight, grad_sampling_loc, grad_attn_weight);
  }
  const void* anchor = d.n_value() ? grad_value : grad_attn_weight;
  if (!anchor) return fail(MSDA_ERR_INVALID_ARGUMENT, "gradient output pointer is null");
  DeviceGuard guard;
  MSDA_CUDA(guard.enter(anchor));
  const bool det = (flags & MSDA_FLAG_DETERMINISTIC) != 0;
  const bool have_samples = d.n_points() != 0 && d.n_value() != 0;
  // zero-fill grad_value -- unless a later pass writes every element anyway: the deterministic finalize / gather,
  // or the float -> bf16 conversion of the workspace accumulators
  const bool overwritten = have_samples && d.D != 0 && (det || dtype == MSDA_BF16);
  if (d.n_value() && !overwritten) {
    if (!grad_value) return fail(MSDA_ERR_INVALID_ARGUMENT, "grad_value is null");
    MSDA_CUDA(cudaMemsetAsync(grad_value, 0, (size_t)d.n_value() * es, st));
  }
  if (d.n_points() == 0) return MSDA_OK;
  if (!grad_sampling_loc || !grad_attn_weight) return fail(MSDA_ERR_INVALID_ARGUMENT, "grad_loc / grad_w is null");
  if (d.n_value() == 0 || d.D == 0) {  // nothing to sample from: all gradients are zero
    MSDA_CUDA(cudaMemsetAsync(grad_sampling_loc, 0, (size_t)d.n_points() * 2 * ls, st));
    MSDA_CUDA(cudaMemsetAsync(grad_attn_weight, 0, (size_t)d.n_points() * ls, st));
    return MSDA_OK;
  }
  if (!grad_output || !value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_value)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  const size_t need = msda_backward_workspace_bytes(batch, spatial_size, num_heads, channels, num_levels, num_query,
                                                    num_point, dtype, flags);
  if (need && (!workspace || workspace_bytes < need))
    return fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", need, workspace_bytes);
  auto count = [] { g_launches.fetch_add(1, std::memory_order_relaxed); };

  if (det) {
    // ---- bit-reproducible grad_value: 64-bit fixed-point accumulation (see include/msda.h) ----
    const int64_t nv = d.n_value(), n_out = d.rows() * d.D, n_pts = d.n_points();
    const bool sorted = det_sorted_ok(d, dtype, flags);
    const DetLayout lay = sorted ? det_layout(d) : DetLayout{};
    char* ws = static_cast<char*>(workspace);
    auto* acc = reinterpret_cast<unsigned long long*>(ws + (sorted ? lay.off_acc : 0));
    auto* amax = reinterpret_cast<unsigned*>(ws + (sorted ? lay.off_misc : (size_t)nv * 8));
    auto* scale = reinterpret_cast<msda::DetScale*>(reinterpret_cast<char*>(amax) + 16);
    if (sorted) {   // zero the bins, cursors, {amax, scale} and the accumulators; entries are fully overwritten
      MSDA_CUDA(cudaMemsetAsync(ws, 0, lay.off_sums, st));
      MSDA_CUDA(cudaMemsetAsync(ws + lay.off_misc, 0, lay.total - lay.off_misc, st));
    } else {
      MSDA_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    }
    const int g_out = grid_for(n_out, 256, 148 * 8), g_pts = grid_for(n_pts, 256, 148 * 8);
    if (dtype == MSDA_F32) msda::msda_amax_kernel<float><<<g_out, 256, 0, st>>>((const float*)grad_output, n_out, amax);
    else if (dtype == MSDA_F64) msda::msda_amax_kernel<double><<<g_out, 256, 0, st>>>((const double*)grad_output, n_out, amax);
    else msda::msda_amax_kernel<__nv_bfloat16><<<g_out, 256, 0, st>>>((const __nv_bfloat16*)grad_output, n_out, amax);
    count();
    if (dtype == MSDA_F64) msda::msda_amax_kernel<double><<<g_pts, 256, 0, st>>>((const double*)attn_weight, n_pts, amax + 1);
    else msda::msda_amax_kernel<float><<<g_pts, 256, 0, st>>>((const float*)attn_weight, n_pts, amax + 1);
    count();
    msda::msda_det_scale_kernel<<<1, 1, 0, st>>>(amax, scale, dtype == MSDA_F64 ? 44 : 38);
    count();
    MSDA_CUDA(cudaGetLastError());
    if (sorted) {
      // sorted segment reduction (msda_det.cuh): count -> scan -> fill -> gather; no atomics on grad_value
      int* bins = reinterpret_cast<int*>(ws);
      int* cursor = reinterpret_cast<int*>(ws + lay.off_cursor);
      int* sums = reinterpret_cast<int*>(ws + lay.off_sums);
      int4* entries = reinterpret_cast<int4*>(ws + lay.off_entries);
      const float* loc = static_cast<const float*>(sampling_loc);
      const float* w = static_cast<const float*>(attn_weight);
      const int g_bin = grid_for(n_pts, 256, 148 * 32);
      msda::det_bin_kernel<false><<<g_bin, 256, 0, st>>>(loc, w, spatial_shapes, level_start_index, d.H, d.L, d.Q, d.P,
                                                         n_pts, bins, nullptr, nullptr);
      count();
      msda::det_scan_block_kernel<<<(unsigned)lay.scan_blocks, 256, 0, st>>>(bins, bins, sums, lay.n_scan);
      count();
      msda::det_scan_sums_kernel<<<1, 1024, 0, st>>>(sums, (int)lay.scan_blocks);
      count();
      msda::det_scan_add_kernel<<<(unsigned)lay.scan_blocks, 256, 0, st>>>(bins, sums, lay.n_scan);
      count();
      msda::det_bin_kernel<true><<<g_bin, 256, 0, st>>>(loc, w, spatial_shapes, level_start_index, d.H, d.L, d.Q, d.P,
                                                        n_pts, cursor, bins, entries);
      count();
      MSDA_CUDA(cudaGetLastError());
      if (int s2 = bwd_fast_noscatter(st, d, dtype, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                                      attn_weight, grad_sampling_loc, grad_attn_weight))
        return s2;
      if (!det_dense(d)) {
        if (dtype == MSDA_F32)
          return det_gather<float>(st, d, grad_output, entries, bins, spatial_shapes, level_start_index, scale, grad_value);
        return det_gather<__nv_bfloat16>(st, d, grad_output, entries, bins, spatial_shapes, level_start_index, scale,
                                         grad_value);
      }
      int s3;
      if (dtype == MSDA_F32)
        s3 = det_cell_reduce<float>(st, d, grad_output, entries, bins, lay.bins, spatial_shapes, level_start_index, scale, acc);
      else
        s3 = det_cell_reduce<__nv_bfloat16>(st, d, grad_output, entries, bins, lay.bins, spatial_shapes, level_start_index,
                                            scale, acc);
      if (s3 != MSDA_OK) return s3;
      const int g_fin = grid_for(nv, 256, 148 * 16);
      const long long* sacc = reinterpret_cast<const long long*>(acc);
      if (dtype == MSDA_F32) msda::msda_det_finalize_kernel<float><<<g_fin, 256, 0, st>>>(sacc, (float*)grad_value, nv, scale);
      else msda::msda_det_finalize_kernel<__nv_bfloat16><<<g_fin, 256, 0, st>>>(sacc, (__nv_bfloat16*)grad_value, nv, scale);
      count();
      MSDA_CUDA(cudaGetLastError());
      return MSDA_OK;
    }
    int s;
    if (fast_ok(d, dtype, flags))
      s = bwd_fast_det(st, d, dtype, grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                       acc, grad_sampling_loc, grad_attn_weight, scale);
    else if (dtype == MSDA_F32)
      s = bwd_generic<float, float, unsigned long long>(st, d, grad_output, value, spatial_shapes, level_start_index,
                                                        sampling_loc, attn_weight, acc, grad_sampling_loc,
                                                        grad_attn_weight, scale);
    else if (dtype == MSDA_F64)
      s = bwd_generic<double, double, unsigned long long>(st, d, grad_output, value, spatial_shapes, level_start_index,
                                                          sampling_loc, attn_weight, acc, grad_sampling_loc,
                                                          grad_attn_weight, scale);
    else
      s = bwd_generic<__nv_bfloat16, float, unsigned long long>(st, d, grad_output, value, spatial_shapes,
                                                                level_start_index, sampling_loc, attn_weight, acc,
                                                                grad_sampling_loc, grad_attn_weight, scale);
    if (s != MSDA_OK) return s;
    const int g_val = grid_for(nv, 256, 148 * 16);
    const long long* cacc = reinterpret_cast<const long long*>(acc);
    if (dtype == MSDA_F32) msda::msda_det_finalize_kernel<float><<<g_val, 256, 0, st>>>(cacc, (float*)grad_value, nv, scale);
    else if (dtype == MSDA_F64) msda::msda_det_finalize_kernel<double><<<g_val, 256, 0, st>>>(cacc, (double*)grad_value, nv, scale);
    else msda::msda_det_finalize_kernel<__nv_bfloat16><<<g_val, 256, 0, st>>>(cacc, (__nv_bfloat16*)grad_value, nv, scale);
    count();
    MSDA_CUDA(cudaGetLastError());
    return MSDA_OK;
  }

  float* gv32 = static_cast<float*>(grad_value);
  if (dtype == MSDA_BF16) {
    gv32 = static_cast<float*>(workspace);
    MSDA_CUDA(cudaMemsetAsync(gv32, 0, need, st));
  }

  int s;
  if (fast_ok(d, dtype, flags)) {
    const CoarsePlan plan = coarse_plan(d, dtype, flags, spatial_shapes_host);
    if (plan.budget > 0)
      s = bwd_fast_with_coarse(st, d, dtype, flags, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                               attn_weight, gv32, grad_sampling_loc, grad_attn_weight, plan.budget, plan.exact);
    else
      s = bwd_fast(st, d, dtype, flags, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                   attn_weight, gv32, grad_sampling_loc, grad_attn_weight, 0);
  } else if (dtype == MSDA_F32) {
    s = bwd_generic<float, float, float>(st, d, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                                         attn_weight, gv32, grad_sampling_loc, grad_attn_weight);
  } else if (dtype == MSDA_F64) {
    s = bwd_generic<double, double, double>(st, d, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                                            attn_weight, static_cast<double*>(grad_value), grad_sampling_loc,
                                            grad_attn_weight);
  } else {
    s = bwd_generic<__nv_bfloat16, float, float>(st, d, grad_output, value, spatial_shapes, level_start_index,
                                                 sampling_loc, attn_weight, gv32, grad_sampling_loc, grad_attn_weight);
  }
  if (s != MSDA_OK) return s;
  if (dtype == MSDA_BF16) {
    const int64_t n = d.n_value();
    msda::msda_cast_f32_to_bf16_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(
        gv32, static_cast<__nv_bfloat16*>(grad_value), n);
    count();
    MSDA_CUDA(cudaGetLastError());
  }
  return MSDA_OK;
}

int msda_fused_supported(int channels, int num_levels, int num_point, int spatial_size, int num_heads, int dtype,
                         unsigned flags) {
  const Dims d{1, spatial_size, num_heads, channels, num_levels, 1, num_point};
  return fast_ok(d, dtype, flags) && !(flags & MSDA_FLAG_DETERMINISTIC) ? 1 : 0;
}

int msda_fused_forward(void* stream, const void* value, const int64_t* spatial_shapes,
                       const int64_t* level_start_index, const float* sampling_offsets, const float* attn_logits,
                       const float* reference_points, int ref_dim, const uint8_t* value_padding_mask, int batch,
                       int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                       void* output, int dtype, unsigned flags) {
  g_err[0] = 0;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int s = check_dims(d, dtype)) return s;
  if (ref_dim != 2 && ref_dim != 4) return fail(MSDA_ERR_INVALID_ARGUMENT, "ref_dim must be 2 or 4, got %d", ref_dim);
  if (d.rows() * d.D == 0) return MSDA_OK;
  if (!msda_fused_supported(channels, num_levels, num_point, spatial_size, num_heads, dtype, flags) || d.S == 0)
    return fail(MSDA_ERR_UNSUPPORTED, "fused path needs D in {16,32,64,128}, float/bf16 value, L<=16, 1<=L*P<=64");
  if (!value || !spatial_shapes || !level_start_index || !sampling_offsets || !attn_logits || !reference_points ||
      !output)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  DeviceGuard guard;
  MSDA_CUDA(guard.enter(value));
  msda::FusedArgs fa{};
  fa.ref = reference_points;
  fa.ref_dim = ref_dim;
  fa.inv_P = 1.0f / (float)num_point;
  fa.value_mask = value_padding_mask;
  return fwd_fused(static_cast<cudaStream_t>(stream), d, dtype, value, spatial_shapes, level_start_index,
                   sampling_offsets, attn_logits, output, fa);
}

int msda_fused_backward(void* stream, const void* grad_output, const void* value, const int64_t* spatial_shapes,
                        const int64_t* level_start_index, const float* sampling_offsets, const float* attn_logits,
                        const float* reference_points, int ref_dim, const uint8_t* value_padding_mask, int batch,
                        int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                        void* grad_value, float* grad_offsets, float* grad_logits, void* workspace,
                        size_t workspace_bytes, int dtype, unsigned flags) {
  g_err[0] = 0;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int s = check_dims(d, dtype)) return s;
  if (ref_dim != 2 && ref_dim != 4) return fail(MSDA_ERR_INVALID_ARGUMENT, "ref_dim must be 2 or 4, got %d", ref_dim);
  if (!msda_fused_supported(channels, num_levels, num_point, spatial_size, num_heads, dtype, flags) ||
      d.n_value() == 0 || d.n_points() == 0)
    return fail(MSDA_ERR_UNSUPPORTED, "fused path needs a non-empty problem, D in {16,32,64,128}, float/bf16 value");
  if (!grad_output || !value || !spatial_shapes || !level_start_index || !sampling_offsets || !attn_logits ||
      !reference_points || !grad_value || !grad_offsets || !grad_logits)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard;
  MSDA_CUDA(guard.enter(value));
  const size_t need = msda_backward_workspace_bytes(batch, spatial_size, num_heads, channels, num_levels, num_query,
                                                    num_point, dtype, flags);
  if (need && (!workspace || workspace_bytes < need))
    return fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", need, workspace_bytes);
  float* gv32 = static_cast<float*>(grad_value);
  if (dtype == MSDA_BF16) gv32 = static_cast<float*>(workspace);
  MSDA_CUDA(cudaMemsetAsync(gv32, 0, (size_t)d.n_value() * sizeof(float), st));
  msda::FusedArgs fa{};
  fa.ref = reference_points;
  fa.ref_dim = ref_dim;
  fa.inv_P = 1.0f / (float)num_point;
  fa.value_mask = value_padding_mask;
  if (int s = bwd_fused(st, d, dtype, grad_output, value, spatial_shapes, level_start_index, sampling_offsets,
                        attn_logits, gv32, grad_offsets, grad_logits, fa))
    return s;
  if (dtype == MSDA_BF16) {
    const int64_t n = d.n_value();
    msda::msda_cast_f32_to_bf16_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(
        gv32, static_cast<__nv_bfloat16*>(grad_value), n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    MSDA_CUDA(cudaGetLastError());
  }
  return MSDA_OK;
}

namespace {
// DCNv3 as an MSDeformAttn problem: value = input [N, H_in*W_in, G, C], one level, Q = H_out*W_out, P = K.
int dcn_setup(int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w, int dilation_h,
              int dilation_w, int group, int group_channels, float offset_scale, int batch, int height_in,
              int width_in, int height_out, int width_out, int dtype, Dims* d, msda::FusedArgs* fa) {
  if (kernel_h < 1 || kernel_w < 1 || stride_h < 1 || stride_w < 1 || dilation_h < 1 || dilation_w < 1 || pad_h < 0 ||
      pad_w < 0 || batch < 0 || height_in < 0 || width_in < 0 || height_out < 0 || width_out < 0 || group < 1 ||
      group_channels < 1)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "dcnv3: bad geometry");
  *d = Dims{batch, height_in * width_in, group, group_channels, 1, height_out * width_out, kernel_h * kernel_w};
  if (dtype != MSDA_F32 && dtype != MSDA_BF16) return fail(MSDA_ERR_UNSUPPORTED, "dcnv3: float32 / bfloat16 only");
  if (d->rows() * d->D != 0 && !fast_ok(*d, dtype, 0))
    return fail(MSDA_ERR_UNSUPPORTED,
                "dcnv3: needs group_channels in {16,32,64,128} and kernel_h*kernel_w <= 64 (got %d, %d)",
                group_channels, kernel_h * kernel_w);
  msda::FusedArgs a{};
  a.kernel_h = kernel_h; a.kernel_w = kernel_w; a.stride_h = stride_h; a.stride_w = stride_w;
  a.pad_h = pad_h; a.pad_w = pad_w; a.dil_h = dilation_h; a.dil_w = dilation_w;
  a.height_in = height_in; a.width_in = width_in; a.width_out = width_out > 0 ? width_out : 1;
  a.offset_scale = offset_scale;
  *fa = a;
  return MSDA_OK;
}
}  // namespace

int msda_dcnv3_forward(void* stream, const void* input, const float* offset, const float* mask, int kernel_h,
                       int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w, int dilation_h, int dilation_w,
                       int group, int group_channels, float offset_scale, int batch, int height_in, int width_in,
                       int height_out, int width_out, void* output, int dtype, unsigned flags) {
  (void)flags;
  g_err[0] = 0;
  Dims d;
  msda::FusedArgs fa;
  if (int s = dcn_setup(kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                        group_channels, offset_scale, batch, height_in, width_in, height_out, width_out, dtype, &d, &fa))
    return s;
  if (d.rows() * d.D == 0) return MSDA_OK;
  if (!output) return fail(MSDA_ERR_INVALID_ARGUMENT, "output is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard;
  MSDA_CUDA(guard.enter(output));
  if (d.S == 0) {
    MSDA_CUDA(cudaMemsetAsync(output, 0, (size_t)(d.rows() * d.D) * elem_size(dtype), st));
    return MSDA_OK;
  }
  if (!input || !offset || !mask) return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  return fwd_dcn(st, d, dtype, input, offset, mask, output, fa);
}

int msda_dcnv3_backward(void* stream, const void* grad_output, const void* input, const float* offset,
                        const float* mask, int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                        int dilation_h, int dilation_w, int group, int group_channels, float offset_scale, int batch,
                        int height_in, int width_in, int height_out, int width_out, void* grad_input,
                        float* grad_offset, float* grad_mask, void* workspace, size_t workspace_bytes, int dtype,
                        unsigned flags) {
  (void)flags;
  g_err[0] = 0;
  Dims d;
  msda::FusedArgs fa;
  if (int s = dcn_setup(kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                        group_channels, offset_scale, batch, height_in, width_in, height_out, width_out, dtype, &d, &fa))
    return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d.n_value() == 0 && d.n_points() == 0) return MSDA_OK;
  const void* anchor = d.n_value() ? grad_input : (const void*)grad_mask;
  if (!anchor) return fail(MSDA_ERR_INVALID_ARGUMENT, "gradient output pointer is null");
  DeviceGuard guard;
  MSDA_CUDA(guard.enter(anchor));
  const size_t need = dtype == MSDA_BF16 ? (size_t)d.n_value() * sizeof(float) : 0;
  if (need && (!workspace || workspace_bytes < need))
    return fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", need, workspace_bytes);
  float* gi32 = dtype == MSDA_BF16 ? static_cast<float*>(workspace) : static_cast<float*>(grad_input);
  if (d.n_value()) {
    if (!grad_input) return fail(MSDA_ERR_INVALID_ARGUMENT, "grad_input is null");
    MSDA_CUDA(cudaMemsetAsync(gi32, 0, (size_t)d.n_value() * sizeof(float), st));
    if (dtype == MSDA_BF16 && d.n_points() == 0)
      MSDA_CUDA(cudaMemsetAsync(grad_input, 0, (size_t)d.n_value() * 2, st));
  }
  if (d.n_points() == 0) return MSDA_OK;
  if (!grad_offset || !grad_mask) return fail(MSDA_ERR_INVALID_ARGUMENT, "grad_offset / grad_mask is null");
  if (d.n_value() == 0) {
    MSDA_CUDA(cudaMemsetAsync(grad_offset, 0, (size_t)d.n_points() * 2 * sizeof(float), st));
    MSDA_CUDA(cudaMemsetAsync(grad_mask, 0, (size_t)d.n_points() * sizeof(float), st));
    return MSDA_OK;
  }
  if (!grad_output || !input || !offset || !mask) return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  if (int s = bwd_dcn(st, d, dtype, grad_output, input, offset, mask, gi32, grad_offset, grad_mask, fa)) return s;
  if (dtype == MSDA_BF16) {
    const int64_t n = d.n_value();
    msda::msda_cast_f32_to_bf16_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(
        gi32, static_cast<__nv_bfloat16*>(grad_input), n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    MSDA_CUDA(cudaGetLastError());
  }
  return MSDA_OK;
}

int msda_debug_bookkeeping(void* stream, const float* sampling_loc, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, int batch, int spatial_size, int num_heads, int channels,
                           int num_levels, int num_query, int num_point, int64_t* corner_offsets, float* frac) {
  g_err[0] = 0;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int s = check_dims(d, MSDA_F32)) return s;
  if (d.n_points() == 0) return MSDA_OK;
  if (!sampling_loc || !spatial_shapes || !level_start_index || !corner_offsets || !frac)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  DeviceGuard guard;
  MSDA_CUDA(guard.enter(sampling_loc));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  msda::msda_bookkeeping_kernel<<<grid_for(d.n_points(), 256, 148 * 16), 256, 0, st>>>(
      sampling_loc, spatial_shapes, level_start_index, d.S, d.H, d.D, d.L, d.Q, d.P, d.n_points(), corner_offsets,
      frac);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

}  // extern "C"

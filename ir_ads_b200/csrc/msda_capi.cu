// msda_capi.cu -- host side of libmsda_b200.so: argument checks, kernel selection, launches.
// The C ABI is declared and documented in include/msda.h.
//
// Plays the role of ms_deform_attn_cuda_forward / _backward
// (/root/reference/detrex/layers/csrc/MsDeformAttn/ms_deform_attn_cuda.cu:21-154) and of the
// kernel selector ms_deformable_col2im_cuda (ms_deform_im2col_cuda.cuh:956-1327), minus ATen:
// no allocation, no im2col_step chunk loop (one launch covers the batch), errors returned.
#include "msda_host.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "msda_generic.cuh"
#include "msda_det.cuh"

namespace {
std::atomic<uint64_t> g_launches{0};
thread_local char g_err[512] = "";
}  // namespace

namespace msda_host {

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

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_alignment(std::initializer_list<NamedPtr> ptrs) {
  for (const NamedPtr& p : ptrs)
    if (p.ptr && (reinterpret_cast<uintptr_t>(p.ptr) & 15u) != 0)
      return fail(MSDA_ERR_INVALID_ARGUMENT, "%s (%p) is not 16-byte aligned: the kernels use 128-bit accesses", p.name,
                  p.ptr);
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

}  // namespace msda_host

namespace {
using namespace msda_host;

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

int check_dims(const Dims& d, int dtype) {
  if (d.B < 0 || d.S < 0 || d.H < 0 || d.D < 0 || d.L < 0 || d.Q < 0 || d.P < 0)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "negative dimension (B=%d S=%d H=%d D=%d L=%d Q=%d P=%d)", d.B, d.S, d.H,
                d.D, d.L, d.Q, d.P);
  if (dtype != MSDA_F32 && dtype != MSDA_F64 && dtype != MSDA_BF16)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "unknown dtype tag %d", dtype);
  return MSDA_OK;
}

// Test hook (msda_debug_bookkeeping with MSDA_DEBUG_FAST_RECORDS): builds the SAME PointRec the fast kernels build
// (msda::make_record: locate + unclamped offset + corner-validity bits) and decodes it the way their gather / scatter
// loops do -- o00 = oc & ~15, +H*D for x+1, +W_l*H*D for y+1, corner k used iff bit k -- so the integers the fast
// kernels actually gather from / scatter to are what the test compares bit for bit.
__global__ void msda_fast_records_kernel(const float* __restrict__ loc, const int64_t* __restrict__ shapes,
                                         const int64_t* __restrict__ lsi, int S, int H, int D, int L, int Q, int P,
                                         int64_t npts, int64_t* __restrict__ offs, float* __restrict__ frac) {
  const int HD = H * D;
  for (int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pt < npts; pt += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)((pt / P) % L);
    const int64_t row = pt / P / L;
    const int h = (int)(row % H);
    const int64_t b = row / H / Q;
    const int Hl = (int)shapes[2 * l], Wl = (int)shapes[2 * l + 1];
    const msda::PointRec r = msda::make_record(loc[2 * pt], loc[2 * pt + 1], 1.0f, Hl, Wl, (int)lsi[l], H, h, D);
    const int oc = r.oc;
    const bool gated = (oc & 15) == 0;
    const int o00 = oc & ~15;
    const int o01 = o00 + HD, o10 = o00 + Wl * HD, o11 = o10 + HD;
    const int64_t img = b * (int64_t)S * HD;
    offs[4 * pt + 0] = (oc & 1) ? img + o00 : -1;
    offs[4 * pt + 1] = (oc & 2) ? img + o01 : -1;
    offs[4 * pt + 2] = (oc & 4) ? img + o10 : -1;
    offs[4 * pt + 3] = (oc & 8) ? img + o11 : -1;
    frac[2 * pt] = gated ? 0.0f : r.lw;
    frac[2 * pt + 1] = gated ? 0.0f : r.lh;
  }
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
  count_launch();
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
  count_launch();
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
  size_t off_cursor, off_sums, off_entries, off_misc, off_wamax, off_acc, total;
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
  l.off_wamax = l.off_acc + (det_dense(d) ? up((size_t)d.n_value() * 8) : 0);   // accumulators: cell reduce only
  // two floats per warp of the entry-filing backward (<= 256 threads per CTA)
  l.total = l.off_wamax + up((size_t)(d.rows() * d.D / 128 + 8) * 8);
  return l;
}
// the sorted path needs 32-bit bin / entry / row indices
bool det_sorted_ok(const Dims& d, int dtype, unsigned flags) {
  if (flags & MSDA_FLAG_DET_ATOMIC) return false;
  if (!fast_ok(d, dtype, flags)) return false;
  const int64_t lim = ((int64_t)1 << 31) - 4096;
  return (int64_t)d.B * d.H * msda::det_cells_bound(d.S, d.L) < lim && d.n_points() < lim && d.rows() < lim;
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
  count_launch();
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
  count_launch();
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

size_t elem_size(int dtype) { return dtype == MSDA_F64 ? 8 : (dtype == MSDA_BF16 ? 2 : 4); }

}  // namespace

// ============================================================================================
extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

unsigned msda_build_config(void) {
  unsigned c = 0;
#ifdef MSDA_EXPERIMENTS
  c |= MSDA_BUILD_EXPERIMENTS;
#endif
#ifdef MSDA_EXP_SLIM
  c |= MSDA_BUILD_SLIM;
#endif
  return c;
}

uint64_t msda_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

unsigned msda_debug_fastdiv(unsigned n, unsigned d) { return msda::fastdiv(n, msda::fastdiv_make(d)); }

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
  if (int s = check_alignment({{"value", value}, {"sampling_loc", sampling_loc}, {"attn_weight", attn_weight},
                               {"output", output}}))
    return s;
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
  g_err[0] = 0;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int s = check_dims(d, dtype)) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t es = elem_size(dtype);
  const size_t ls = dtype == MSDA_F64 ? 8 : 4;
  if (d.n_value() == 0 && d.n_points() == 0) return MSDA_OK;
  if (int s = check_alignment({{"grad_output", grad_output}, {"value", value}, {"sampling_loc", sampling_loc},
                               {"attn_weight", attn_weight}, {"grad_value", grad_value},
                               {"grad_sampling_loc", grad_sampling_loc}, {"grad_attn_weight", grad_attn_weight},
                               {"workspace", workspace}}))
    return s;
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
  auto count = [] { count_launch(); };

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
    // The fixed-point scale comes from max |grad_out| x max |weight|.  On the sorted route the backward kernel that
    // files the entries reads both tensors anyway and takes the two maxima itself (EmitArgs::amax); the scale is only
    // needed by the reduction that follows it.  Every other route scatters with the scale and takes them up front.
#ifndef MSDA_DET_AMAX_IN_EMIT
#define MSDA_DET_AMAX_IN_EMIT 1
#endif
    const bool amax_in_emit = MSDA_DET_AMAX_IN_EMIT && sorted && !(flags & MSDA_FLAG_DET_SEPARATE_FILL);
    const int g_out = grid_for(n_out, 256, 148 * 8), g_pts = grid_for(n_pts, 256, 148 * 8);
    if (!amax_in_emit) {
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
    }
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
      MSDA_CUDA(cudaGetLastError());
      // entries are filed by the backward kernel itself (ACC = EmitEntries): the lane that has just built a point's
      // record takes the cursor atomic and writes the 16-byte entry, so the separate fill pass (det_bin_kernel<true>:
      // re-read of every location / weight + the coordinate arithmetic a second time) is gone
      if (flags & MSDA_FLAG_DET_SEPARATE_FILL) {
        msda::det_bin_kernel<true><<<g_bin, 256, 0, st>>>(loc, w, spatial_shapes, level_start_index, d.H, d.L, d.Q, d.P,
                                                          n_pts, cursor, bins, entries);
        count();
        MSDA_CUDA(cudaGetLastError());
        if (int s2 = bwd_fast_noscatter(st, d, dtype, grad_output, value, spatial_shapes, level_start_index,
                                        sampling_loc, attn_weight, grad_sampling_loc, grad_attn_weight))
          return s2;
      } else if (amax_in_emit) {
        // the kernel leaves two maxima per warp; reducing those short arrays replaces the sweeps over grad_out / weights
        float* wamax = reinterpret_cast<float*>(ws + lay.off_wamax);
        int64_t n_warps = 0;
        if (int s2 = bwd_fast_emit(st, d, dtype, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                                   attn_weight, grad_sampling_loc, grad_attn_weight, cursor, bins, entries, wamax,
                                   &n_warps))
          return s2;
        const int g_w = grid_for(n_warps, 256, 148 * 8);
        msda::msda_amax_kernel<float><<<g_w, 256, 0, st>>>(wamax, n_warps, amax);
        count();
        msda::msda_amax_kernel<float><<<g_w, 256, 0, st>>>(wamax + n_warps, n_warps, amax + 1);
        count();
        msda::msda_det_scale_kernel<<<1, 1, 0, st>>>(amax, scale, 38);   // sorted route: float / bf16 only
        count();
        MSDA_CUDA(cudaGetLastError());
      } else if (int s2 = bwd_fast_emit(st, d, dtype, grad_output, value, spatial_shapes, level_start_index,
                                        sampling_loc, attn_weight, grad_sampling_loc, grad_attn_weight, cursor, bins,
                                        entries, nullptr, nullptr)) {
        return s2;
      }
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
    if (fold_applies(d, dtype, flags))
      s = bwd_fold(st, d, dtype, grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, gv32,
                   grad_sampling_loc, grad_attn_weight, nullptr);
    else
      s = bwd_fast(st, d, dtype, flags, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                   attn_weight, gv32, grad_sampling_loc, grad_attn_weight);
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
  if (int s = check_alignment({{"value", value}, {"sampling_offsets", sampling_offsets}, {"attn_logits", attn_logits},
                               {"reference_points", reference_points}, {"output", output}}))
    return s;
  DeviceGuard guard;
  MSDA_CUDA(guard.enter(value));
  msda::FusedArgs fa{};
  fa.ref = reference_points;
  fa.ref_dim = ref_dim;
  fa.inv_P = 1.0f / (float)num_point;
  fa.num_P = (float)num_point;
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
  if (int s = check_alignment({{"grad_output", grad_output}, {"value", value}, {"sampling_offsets", sampling_offsets},
                               {"attn_logits", attn_logits}, {"reference_points", reference_points},
                               {"grad_value", grad_value}, {"grad_offsets", grad_offsets}, {"grad_logits", grad_logits},
                               {"workspace", workspace}}))
    return s;
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
  fa.num_P = (float)num_point;
  fa.value_mask = value_padding_mask;
  int s;
  if (fold_applies(d, dtype, flags))
    s = bwd_fold(st, d, dtype, grad_output, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                 gv32, grad_offsets, grad_logits, &fa);
  else
    s = bwd_fused(st, d, dtype, grad_output, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                  gv32, grad_offsets, grad_logits, fa);
  if (s != MSDA_OK) return s;
  if (dtype == MSDA_BF16) {
    const int64_t n = d.n_value();
    msda::msda_cast_f32_to_bf16_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(
        gv32, static_cast<__nv_bfloat16*>(grad_value), n);
    count_launch();
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
  if (int s = check_alignment({{"input", input}, {"offset", offset}, {"mask", mask}, {"output", output}})) return s;
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
  if (int s = check_alignment({{"grad_output", grad_output}, {"input", input}, {"offset", offset}, {"mask", mask},
                               {"grad_input", grad_input}, {"grad_offset", grad_offset}, {"grad_mask", grad_mask},
                               {"workspace", workspace}}))
    return s;
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
    count_launch();
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
  if (fast_ok(d, MSDA_F32, 0))   // what the fast kernels index with (their own records, decoded)
    msda_fast_records_kernel<<<grid_for(d.n_points(), 256, 148 * 16), 256, 0, st>>>(
        sampling_loc, spatial_shapes, level_start_index, d.S, d.H, d.D, d.L, d.Q, d.P, d.n_points(), corner_offsets,
        frac);
  else
    msda::msda_bookkeeping_kernel<<<grid_for(d.n_points(), 256, 148 * 16), 256, 0, st>>>(
        sampling_loc, spatial_shapes, level_start_index, d.S, d.H, d.D, d.L, d.Q, d.P, d.n_points(), corner_offsets,
        frac);
  count_launch();
  MSDA_CUDA(cudaGetLastError());
  return MSDA_OK;
}

}  // extern "C"

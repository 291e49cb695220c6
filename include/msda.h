/*
 * msda.h -- C ABI of the B200-native multi-scale deformable attention library (libmsda_b200.so).
 *
 * This is the drop-in boundary for the reference's native MSDeformAttn pair.  A maintainer of
 * the reference binds these entry points where detrex binds its own kernels today
 * (INTEGRATION.md shows the ctypes / pybind stub):
 *
 *   msda_forward   replaces  ms_deformable_im2col_cuda<T>
 *                  (/root/reference/detrex/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:923-954),
 *                  i.e. the body of ms_deform_attn_cuda_forward (ms_deform_attn_cuda.cu:21-81),
 *                  exported to Python as detrex._C.ms_deform_attn_forward (csrc/vision.cpp:55).
 *   msda_backward  replaces  ms_deformable_col2im_cuda<T> (ms_deform_im2col_cuda.cuh:956-1327),
 *                  i.e. the body of ms_deform_attn_cuda_backward (ms_deform_attn_cuda.cu:84-154),
 *                  exported as detrex._C.ms_deform_attn_backward (csrc/vision.cpp:56).
 *
 * Argument order and meaning follow those two launchers: stream first, then the tensors, then
 * (batch, spatial_size, num_heads, channels, num_levels, num_query, num_point), then outputs.
 * Differences, all deliberate:
 *   - plain C: no ATen / torch types, `void*` for the stream (a cudaStream_t);
 *   - a dtype tag instead of a C++ template: float, double, or bfloat16 value with float
 *     locations / weights (the reference has float and double only);
 *   - errors are RETURNED (the reference only printf's launch failures, cuh:948-952);
 *   - no im2col_step: the whole batch is one launch (ms_deform_attn_cuda.cu:51-76 chunked it);
 *   - backward takes a caller-owned workspace (size from msda_backward_workspace_bytes) because
 *     the library never allocates device memory.
 *
 * Tensor layouts (all contiguous, all device pointers, as the reference asserts
 * ms_deform_attn_cuda.cu:29-39):
 *   value              [B, S, H, D]        dtype
 *   spatial_shapes     [L, 2]  int64  (H_l, W_l)
 *   level_start_index  [L]     int64
 *   sampling_loc       [B, Q, H, L, P, 2]  (x, y) normalised to [0,1] over the level; float
 *                                          (double when dtype == MSDA_F64)
 *   attn_weight        [B, Q, H, L, P]     same type as sampling_loc
 *   output / grad_output        [B, Q, H*D]   dtype
 *   grad_value         [B, S, H, D]        dtype        (zero-filled by the library)
 *   grad_sampling_loc, grad_attn_weight    shaped/typed like sampling_loc / attn_weight
 *                                          (fully written by the library, no pre-zeroing needed)
 * Every tensor pointer (workspaces included) must be 16-byte aligned: rows are read and written with 128-bit
 * accesses.  A torch / cudaMalloc allocation is; a view with an odd storage offset may not be -- such a call returns
 * MSDA_ERR_INVALID_ARGUMENT (the Python wrappers re-align by copying).  The reference's scalar kernels had no such
 * requirement.
 *
 * Thread safety: re-entrant, no global mutable state except a launch counter and a
 * thread-local last-error string; no library-owned streams, events or locks.  Work is enqueued on `stream` of the device that owns `value`
 * (the library switches to that device for the call and restores the previous one).
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 3

/* dtype tags */
#define MSDA_F32 0  /* value/out/grads float,  loc/w float  */
#define MSDA_F64 1  /* everything double (gradcheck path; generic kernels only) */
#define MSDA_BF16 2 /* value/out/grad_out/grad_value bfloat16, loc/w and their grads float, fp32 accumulate */

/* status codes */
#define MSDA_OK 0
#define MSDA_ERR_INVALID_ARGUMENT 1 /* null pointer, negative size, unknown dtype ...            */
#define MSDA_ERR_UNSUPPORTED 2      /* shape outside what any kernel handles                     */
#define MSDA_ERR_WORKSPACE 3        /* workspace missing or smaller than msda_backward_workspace_bytes */
#define MSDA_ERR_CUDA 4             /* a CUDA runtime call or kernel launch failed; see msda_last_error_message */

/* flags (bit set) */
/* Backward: bit-reproducible grad_value.  Contributions are accumulated as 64-bit fixed point
 * (integer atomics commute, float atomics do not) with a power-of-two scale derived from
 * max|grad_output| * max|attn_weight|, then converted once: run-to-run identical, absolute error
 * <= 2^-38 of that bound per contribution, headroom for 2^25 maximal contributions per element.
 * On the fast shapes the accumulation is a sorted segment reduction (points binned by bilinear cell, one
 * lane group per output pixel gathers its bins and sums in registers: no atomics on grad_value, no
 * zero-fill); elsewhere 64-bit integer reds.  Both produce identical bits.
 * grad_sampling_loc / grad_attn_weight are reproducible in every mode (fixed reduction order). */
#define MSDA_FLAG_DETERMINISTIC (1u << 0)
/* With MSDA_FLAG_DETERMINISTIC: use the fixed-point RED path even where the sorted segment reduction
 * (msda_det.cuh) is available (A/B testing; both give the same bits). */
#define MSDA_FLAG_DET_ATOMIC (1u << 6)
/* With MSDA_FLAG_DETERMINISTIC (sorted path): fill the bins with a separate pass over the locations instead of from
 * the backward kernel itself (A/B testing; both give the same bits, the separate pass is slower). */
#define MSDA_FLAG_DET_SEPARATE_FILL (1u << 14)
#define MSDA_FLAG_FORCE_GENERIC (1u << 1) /* bypass the D in {16,32,64,128} fast kernels             */
/* Row order = which rows a CTA works on.  A scheduling choice only: results never depend on it.  The PRODUCT library
 * carries one order per pass -- forward LINEAR, backward STRIP, the fastest pair on every measured configuration -- and
 * treats the three flags below as hints it ignores; experiment builds (msda_build_config() & MSDA_BUILD_EXPERIMENTS) and
 * the variant builds of tools/build_variant.sh instantiate all three orders for both passes and honour them. */
#define MSDA_FLAG_ORDER_LINEAR (1u << 2)  /* rows in memory order (b, q, h)                              */
/* One CTA = a strip of consecutive queries of ONE head (any Q): x-adjacent queries re-use corner lines in L1. */
#define MSDA_FLAG_ORDER_STRIP (1u << 4)
/* Encoder form only (Q == S): one CTA = a pixel tile (8x4 at D=32) of one level and ONE head. */
#define MSDA_FLAG_ORDER_TILE2D (1u << 5)
/* Bits 3 and 7-10 selected round-1 experiments (persistent TILED order, shared-memory accumulation of the coarse
 * levels on a side stream, head-major strips) that were measured slower and removed; they are reserved and
 * ignored. */

/* EXPERIMENT BUILDS ONLY (msda_build_config() & MSDA_BUILD_EXPERIMENTS; the product library ignores both bits).
 * Backward, encoder form (Q == S, query i is pixel i of the pyramid), D in {32, 64}, float accumulation:
 * grad_value contributions of an 8x8 query tile of one head are pre-added ON THE SM (pixel-keyed lists in
 * shared memory, csrc/msda_fold.cuh) and leave it as ONE vector red per distinct destination row instead of one
 * per (point, corner): 5.5x fewer reds on model-like inputs (profiles/r02a_fold_rate_cfg2.json) -- and 2-2.5x
 * SLOWER, because the bookkeeping more than doubles the instruction count (profiles/r02b_fold_experiment.txt).
 * grad_sampling_loc / grad_attn_weight are bit-identical to the other kernels'; grad_value differs by float
 * summation order only.  FOLD_OFF wins over FOLD_ON; an explicit ORDER_* flag, MSDA_FLAG_DETERMINISTIC and
 * MSDA_FLAG_FORCE_GENERIC also select the non-folding kernels. */
#define MSDA_FLAG_FOLD_ON (1u << 12)
#define MSDA_FLAG_FOLD_OFF (1u << 13)

/* Backward: the caller does not need grad_value (autograd: value.requires_grad is False, e.g. a frozen memory
 * branch).  On the shapes the fast kernels cover (msda_dispatch_name(..., backward=1) is "bwd_fast_...") the
 * grad_value scatter -- the dominant cost of the backward -- is compiled out: grad_value may be NULL and is left
 * untouched, no workspace is needed, grad_sampling_loc / grad_attn_weight are as always.  Elsewhere the flag is
 * ignored and grad_value must be a valid buffer. */
#define MSDA_FLAG_NO_GRAD_VALUE (1u << 11)

int msda_abi_version(void);
/* How this library was built: 0 for the product library. */
#define MSDA_BUILD_EXPERIMENTS 1u /* -DMSDA_EXPERIMENTS: the measured-slower experiments are compiled in   */
#define MSDA_BUILD_SLIM 2u        /* -DMSDA_EXP_SLIM: only D = 32, P in {4, 8} of the plain operator       */
unsigned msda_build_config(void);

/* Forward.  Returns MSDA_OK or an error code.  B*Q*H*D == 0 is a no-op. */
int msda_forward(void* stream, const void* value, const int64_t* spatial_shapes,
                 const int64_t* level_start_index, const void* sampling_loc, const void* attn_weight,
                 int batch, int spatial_size, int num_heads, int channels, int num_levels,
                 int num_query, int num_point, void* output, int dtype, unsigned flags);

/* Bytes of device scratch msda_backward needs for this problem (0 is possible). */
size_t msda_backward_workspace_bytes(int batch, int spatial_size, int num_heads, int channels,
                                     int num_levels, int num_query, int num_point, int dtype,
                                     unsigned flags);

/* Backward: the three gradients of the forward above w.r.t. value, sampling_loc, attn_weight. */
int msda_backward(void* stream, const void* grad_output, const void* value,
                  const int64_t* spatial_shapes, const int64_t* level_start_index,
                  const void* sampling_loc, const void* attn_weight, int batch, int spatial_size,
                  int num_heads, int channels, int num_levels, int num_query, int num_point,
                  void* grad_value, void* grad_sampling_loc, void* grad_attn_weight,
                  void* workspace, size_t workspace_bytes, int dtype, unsigned flags);

/*
 * Fused module path (an extension, the reference has no counterpart): the
 * softmax over L*P and the sampling-location arithmetic of MultiScaleDeformableAttention.forward
 * (/root/reference/detrex/layers/multi_scale_deform_attn.py:300-332) run inside the kernels, so
 * sampling_locations / attention_weights and their gradients never exist in HBM.
 *   sampling_offsets  [B, Q, H, L, P, 2] float  raw output of the module's sampling_offsets Linear
 *   attn_logits       [B, Q, H, L*P]     float  raw output of its attention_weights Linear (pre-softmax)
 *   reference_points  [B, Q, L, ref_dim] float  ref_dim 2: loc = ref + off / (W_l, H_l)   (py:320-324)
 *                                               ref_dim 4: loc = ref_xy + off / P * ref_wh * 0.5 (py:326-332)
 *   value_padding_mask [B, S] uint8, or NULL    the module's key_padding_mask (non-zero = padded pixel): its
 *                                               value.masked_fill(key_padding_mask[..., None], 0) (py:291-292)
 *                                               is folded into the kernels -- a masked pixel's value row reads
 *                                               as zeros in the output and in every gradient, and its
 *                                               grad_value is 0 -- so `value` is passed UNMASKED and the
 *                                               masked copy (one read + one write of value per layer, and the
 *                                               same again in backward) never exists
 * Backward returns grad_value plus the gradients w.r.t. the two RAW tensors (what the Linear layers'
 * backward consumes); it uses the same workspace rule as msda_backward.  Only the fast kernels have
 * a fused form: otherwise MSDA_ERR_UNSUPPORTED is returned (query with msda_fused_supported) and the
 * caller composes softmax / affine itself around msda_forward / msda_backward.
 */
int msda_fused_supported(int channels, int num_levels, int num_point, int spatial_size, int num_heads,
                         int dtype, unsigned flags);
int msda_fused_forward(void* stream, const void* value, const int64_t* spatial_shapes,
                       const int64_t* level_start_index, const float* sampling_offsets,
                       const float* attn_logits, const float* reference_points, int ref_dim,
                       const uint8_t* value_padding_mask, int batch, int spatial_size, int num_heads,
                       int channels, int num_levels, int num_query, int num_point, void* output, int dtype,
                       unsigned flags);
int msda_fused_backward(void* stream, const void* grad_output, const void* value,
                        const int64_t* spatial_shapes, const int64_t* level_start_index,
                        const float* sampling_offsets, const float* attn_logits,
                        const float* reference_points, int ref_dim, const uint8_t* value_padding_mask,
                        int batch, int spatial_size, int num_heads, int channels, int num_levels,
                        int num_query, int num_point, void* grad_value, float* grad_offsets,
                        float* grad_logits, void* workspace, size_t workspace_bytes, int dtype,
                        unsigned flags);

/*
 * DCNv3 core op: the other native op of detrex._C, same gather/scatter core.
 *   msda_dcnv3_forward   replaces dcnv3_im2col_cuda<T>
 *       (/root/reference/detrex/layers/csrc/DCNv3/dcnv3_im2col_cuda.cuh:840-868), i.e. _C.dcnv3_forward
 *       (csrc/vision.cpp:57); msda_dcnv3_backward replaces dcnv3_col2im_cuda<T> (:870-), _C.dcnv3_backward.
 * Argument order follows those launchers.  Layouts (contiguous, channels last, as dcn_v3.py:464-489):
 *   input  [N, H_in, W_in, group*group_channels]   dtype (MSDA_F32 or MSDA_BF16)
 *   offset [N, H_out, W_out, group*K*2] float, K = kernel_h*kernel_w, point index = i*kernel_h + j
 *          with i over kernel_w (the reference kernel's loop order), (x, y) per point, in pixels
 *   mask   [N, H_out, W_out, group*K]   float
 *   output / grad_output [N, H_out, W_out, group*group_channels] dtype
 * The input is NOT padded by the caller: pad_h / pad_w only shift the sampling grid and samples outside
 * the map read zeros, exactly like the reference kernel.  Supported: group_channels in {16,32,64,128},
 * K <= 64, float32 / bfloat16 input (MSDA_ERR_UNSUPPORTED otherwise).  bfloat16 backward needs a
 * workspace of N*H_in*W_in*C floats.
 */
int msda_dcnv3_forward(void* stream, const void* input, const float* offset, const float* mask,
                       int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                       int dilation_h, int dilation_w, int group, int group_channels, float offset_scale,
                       int batch, int height_in, int width_in, int height_out, int width_out,
                       void* output, int dtype, unsigned flags);
int msda_dcnv3_backward(void* stream, const void* grad_output, const void* input, const float* offset,
                        const float* mask, int kernel_h, int kernel_w, int stride_h, int stride_w,
                        int pad_h, int pad_w, int dilation_h, int dilation_w, int group,
                        int group_channels, float offset_scale, int batch, int height_in, int width_in,
                        int height_out, int width_out, void* grad_input, float* grad_offset,
                        float* grad_mask, void* workspace, size_t workspace_bytes, int dtype,
                        unsigned flags);

/*
 * Fused residual add + LayerNorm: the epilogue that follows every MSDeformAttn call and every
 * FFN in the reference's transformer layers -- x = x + identity; x = norm(x)
 * (/root/reference/detrex/layers/transformer.py:152-192; the residuals are added at multi_scale_deform_attn.py:363 and
 * detrex/layers/mlp.py:127-132) -- as one pass instead of an add kernel plus a LayerNorm kernel.
 *   a, b, y, grad_y, grad_x   [rows, channels]  dtype (MSDA_F32 or MSDA_BF16), contiguous
 *   gamma, beta, grad_gamma, grad_beta  [channels] float;   mean, rstd  [rows] float (saved by forward for backward)
 * forward:  y = (a + b - mean) * rstd * gamma + beta, statistics over the channels, fp32 arithmetic.
 * backward: grad_x = d loss / d (a + b) -- the gradient of BOTH addends -- plus grad_gamma / grad_beta, reduced in a
 *           fixed order (bit-reproducible); a + b is recomputed, never stored.  Needs a workspace of
 *           msda_add_layernorm_workspace_bytes(rows, channels).
 * channels must be a multiple of 4 and at most 1024 (MSDA_ERR_UNSUPPORTED otherwise).
 */
size_t msda_add_layernorm_workspace_bytes(int64_t rows, int channels);
int msda_add_layernorm_forward(void* stream, const void* a, const void* b, const float* gamma, const float* beta,
                               int64_t rows, int channels, float eps, void* y, float* mean, float* rstd, int dtype);
int msda_add_layernorm_backward(void* stream, const void* grad_y, const void* a, const void* b, const float* gamma,
                                const float* mean, const float* rstd, int64_t rows, int channels, void* grad_x,
                                float* grad_gamma, float* grad_beta, void* workspace, size_t workspace_bytes,
                                int dtype);

/*
 * Test hook: the integer bookkeeping the float kernels derive from every sampling point.
 *   corner_offsets [B*Q*H*L*P, 4] int64  flat element offset (channel 0) of the four bilinear
 *                  corners inside `value`, -1 for a zero-padded corner or a gated-out point;
 *   frac           [B*Q*H*L*P, 2] float  (lw, lh) fractional weights.
 * Same definition as msda_oracle_bookkeeping in oracle/msda_oracle.c; compared bit for bit.
 * On the shapes the fast kernels cover (float dispatch name "*_fast_*") the numbers are decoded from the very
 * record the fast kernels build per point (unclamped low-corner offset + four corner-validity bits), i.e. they are
 * the addresses those kernels gather from and scatter to; elsewhere they come from the generic kernels'
 * coordinate code.
 */
int msda_debug_bookkeeping(void* stream, const float* sampling_loc, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, int batch, int spatial_size,
                           int num_heads, int channels, int num_levels, int num_query,
                           int num_point, int64_t* corner_offsets, float* frac);

/* Test hook (host arithmetic only, no GPU): n / d as the deterministic path's count pass computes it -- multiply-high
 * by a constant derived from d (csrc/msda_det.cuh, FastDiv).  Exact for n < 2^31 and 1 <= d < 2^31. */
unsigned msda_debug_fastdiv(unsigned n, unsigned d);

/* Human-readable name of a status code. */
const char* msda_status_string(int status);
/* Detail of the last failure on the calling thread ("" if none). */
const char* msda_last_error_message(void);
/* Number of kernels this library has launched in this process (monotonic). */
uint64_t msda_kernel_launch_count(void);
/* Name of the kernel family the given problem dispatches to ("fast_d32_f32", "generic_f64", ...). */
const char* msda_dispatch_name(int channels, int num_levels, int num_point, int spatial_size,
                               int num_heads, int dtype, unsigned flags, int backward);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */

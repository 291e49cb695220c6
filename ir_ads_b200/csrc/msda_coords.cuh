// msda_coords.cuh -- where does a normalised sampling location land on a level?
//
// Reference semantics (/root/reference/detrex/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh):
//   :284-285  h_im = loc_h * H_l - 0.5, w_im = loc_w * W_l - 0.5
//   :288      the point contributes only if h_im > -1 && w_im > -1 && h_im < H_l && w_im < W_l
//   :39-46    cell = floor(), fractional weights lh, lw
//   :56-80    each of the four corners is zero-padded individually
//
// Float path: the reference rounds loc*size to fp32 before it subtracts 0.5, which costs up to
// ~4e-6 px at size ~170 and occasionally picks the neighbouring cell.  Here the rounding errors of
// the product (one fma: e = fma(loc,size,-p) is exact) and of the subtraction (TwoSum) are carried
// along, so cell + fraction equal the EXACT value of loc*size-0.5 to the last float bit of the
// fraction -- the same cell the fp64 grid_sample oracle picks (when the exact fraction rounds up to
// 1.0 the result is (cell, 1.0), the same sample as (cell+1, 0)).
// oracle/msda_oracle.c::split_f32_compensated restates these lines operation for operation; the
// bookkeeping test compares the two bit for bit, so keep them in lock step.
#pragma once
#include <cuda_runtime.h>

namespace msda {

// Fixed-point scale of the deterministic backward (MSDA_FLAG_DETERMINISTIC): written on the device
// by msda_det_scale_kernel into the caller's workspace, read by the backward kernels.
struct DetScale {
  float scale;      // 2^k
  float inv_scale;  // 2^-k (NaN when the inputs were not finite)
};

// One fixed-point contribution is round((w * scale) * g) -- det_contrib below: the corner weight w (bilinear weight x attention weight)
// is scaled first -- scale is a power of two, so that product is exact -- which leaves ONE multiplication per
// channel.  The three deterministic paths of the fast shapes (fixed-point reds in msda_fast.cuh, per-pixel gather and
// cell reduce in msda_det.cuh) form their contributions with exactly this expression, which is what makes their
// results bit-identical.  (The product only differs from (w * g) * scale when w * g is subnormal.)
__device__ __forceinline__ float det_weight(float w, float scale) { return __fmul_rn(w, scale); }

// The contribution itself, det_contrib(w * scale, g), as a 64-bit integer.
//   MSDA_DET_FP64 = 1 (default): the round-to-nearest-even integer of the EXACT product.  Both factors are floats, so
//     their product has at most 48 significant bits and is exact in double; ONE fused multiply-add with the addend
//     1.5 * 2^52 rounds it to an integer and leaves that integer in the low mantissa bits (|product| <= 2^38 by the
//     choice of the scale, far inside the trick's 2^51).  One rounding instead of two, and per entry of the cell reduce
//     8 F2F.F64.F32 + 16 DFMA instead of 16 FMUL + 16 F2I.S64: every float conversion issues at 16 lanes per clock and
//     SM on B200, DFMA at 64 (profiles/r02ak_cvt_rates.txt).  Measured: cfg 5 backward 4.70 -> 4.67 ms, the sparse
//     decoder problem 0.603 -> 0.587 ms -- halving the conversion-pipe cycles barely moves the cell reduce, so that pipe
//     was not what bound it (profiles/r02af_det_variants.txt, item 5).
//   MSDA_DET_FP64 = 0 (variant builds): the round-2 definition, the float product rounded once more to an integer.
#ifndef MSDA_DET_FP64
#define MSDA_DET_FP64 1
#endif
#if MSDA_DET_FP64
using det_factor = double;
__device__ __forceinline__ long long det_contrib(double ws, double g) {
  return __double_as_longlong(__fma_rn(ws, g, 6755399441055744.0)) - 0x4338000000000000LL;   // 1.5 * 2^52 and its bits
}
#else
using det_factor = float;
__device__ __forceinline__ long long det_contrib(float ws, float g) { return __float2ll_rn(ws * g); }
#endif

template <typename T>
struct AxisSplit {
  int low;     // floor(coord) (0 when !ok)
  T frac;      // coord - low  (0 when !ok)
  bool ok;     // coord > -1 && coord < size   (false for NaN / Inf)
};

__device__ __forceinline__ AxisSplit<float> split_axis(float loc, int size) {
  AxisSplit<float> s;
  const float sz = (float)size;
  const float p = __fmul_rn(loc, sz);
  const float e = __fmaf_rn(loc, sz, -p);                 // loc*sz == p + e exactly
  const float a = __fsub_rn(p, 0.5f);
  const float bb = __fsub_rn(a, p);                       // TwoSum: p - 0.5 == a + ea exactly
  const float ea = __fadd_rn(__fsub_rn(p, __fsub_rn(a, bb)), __fsub_rn(-0.5f, bb));
  float f0 = floorf(a);
  float r = __fadd_rn(__fsub_rn(a, f0), __fadd_rn(ea, e));   // a - floor(a) is exact whenever it is representable
  if (r < 0.0f) {
    f0 -= 1.0f;
    r = __fadd_rn(r, 1.0f);
  } else if (r >= 1.0f) {
    f0 += 1.0f;
    r = __fsub_rn(r, 1.0f);
  }
  s.ok = (f0 >= 0.0f || (f0 == -1.0f && r > 0.0f)) && (f0 < sz);
  s.low = s.ok ? (int)f0 : 0;
  s.frac = s.ok ? r : 0.0f;
  return s;
}

__device__ __forceinline__ AxisSplit<double> split_axis(double loc, int size) {
  AxisSplit<double> s;
  const double c = loc * (double)size - 0.5;  // the reference expression, evaluated in double
  s.ok = (c > -1.0) && (c < (double)size);
  const double fl = floor(c);
  s.low = s.ok ? (int)fl : 0;
  s.frac = s.ok ? (c - fl) : 0.0;
  return s;
}

// One sampling point on one level.
template <typename T>
struct Cell {
  int y0, x0;       // low corner (may be -1)
  T lh, lw;         // fractional parts
  unsigned valid;   // bit k set <=> corner k is inside the map; 0 when the point is gated out
                    // k: 0=(y0,x0) 1=(y0,x0+1) 2=(y0+1,x0) 3=(y0+1,x0+1)   (v1..v4 of cuh:56-80)
};

template <typename T>
__device__ __forceinline__ Cell<T> locate(T loc_x, T loc_y, int H, int W) {
  Cell<T> c;
  const AxisSplit<T> ay = split_axis(loc_y, H);
  const AxisSplit<T> ax = split_axis(loc_x, W);
  c.y0 = ay.low;
  c.x0 = ax.low;
  c.lh = ay.frac;
  c.lw = ax.frac;
  unsigned v = 0;
  if (ay.ok && ax.ok) {
    const bool y0ok = c.y0 >= 0, y1ok = c.y0 + 1 <= H - 1;
    const bool x0ok = c.x0 >= 0, x1ok = c.x0 + 1 <= W - 1;
    v = (unsigned)(y0ok && x0ok) | ((unsigned)(y0ok && x1ok) << 1) | ((unsigned)(y1ok && x0ok) << 2) |
        ((unsigned)(y1ok && x1ok) << 3);
  }
  c.valid = v;
  return c;
}

}  // namespace msda

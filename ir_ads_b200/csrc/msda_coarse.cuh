// msda_coarse.cuh -- backward: grad_value of the COARSE pyramid levels accumulated in shared memory.
//
// Why.  msda_bwd_fast_kernel adds every bilinear corner of every point into grad_value with a 128-byte
// vector red; at the encoder shape that is 91 M rows = 11.65 GB through the L2 atomic path, which sustains
// ~6.8 TB/s (profiles/r01m_microbench_ceilings.txt): 1.7 ms, the whole backward.  But the two (three)
// coarsest levels of one head are tiny -- (25x42 + 13x21) pixels x 128 B = 169 KB at DINO-R50 800x1333 --
// and receive HALF of all the rows (every query samples every level), hundreds of rows per pixel.  Those
// fit in ONE CTA's shared memory, so their rows can be pre-aggregated on chip and reach L2 as one red per
// pixel and per CTA instead of one per sample: the red traffic of the backward halves.
//
// How.  Shared memory has no native float atomic (atomicAdd on a shared float compiles to an
// ATOMS.CAST.SPIN loop), so the accumulation is made conflict-free by OWNERSHIP instead:
//   * a persistent CTA (one per SM) owns a contiguous range of (image, head, query) and keeps the resident
//     levels' grad_value tile of the current (image, head) in shared memory;
//   * inside the CTA one WARP per (resident level, pixel-parity class (x&1, y&1)) owns the pixels of that
//     class: the four corners of any bilinear cell fall into the four different classes of its level, so
//     every warp processes every point of its level but exactly ONE corner of it, and no two warps ever
//     touch the same pixel;
//   * inside the warp lane c owns channel(s) c of the pixel row, so consecutive read-modify-writes of the
//     same pixel are same-thread program order: no atomics, no barriers, no shuffles on the data path.
// Per 32 queries a warp first turns the level's points into (tile offset, corner weight) records, one query
// per lane (the coordinate arithmetic is msda_coords.cuh, the same cell the fast kernels pick), folds points
// of one query that hit the same pixel, then walks the 32 queries: broadcast the records, load the grad_out
// row (one coalesced 128-byte load), and do up to four independent LDS / FFMA / STS per query.
// When the (image, head) changes -- and at the end -- the tile is added to grad_value with vector reds and
// cleared.  The fast kernel is told (same shapes, same budget => same decision, made on the device) to skip
// the reds of the resident levels; grad_sampling_loc / grad_attn_weight stay entirely with it.
//
// The reference has no counterpart: its col2im kernels issue one scalar atomicAdd per channel and corner
// (/root/reference/detrex/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:125-152).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_coords.cuh"
#include "msda_fast.cuh"   // kCoarseMaxLevels, coarse_first_level, kFastMaxLevels

namespace msda {

// 8 warps: with two resident levels one warp per (level, parity class); with more, a warp takes a second class
// in a second pass over its queries.  Few threads and <= 64 registers on purpose: the kernel shares its SM with
// CTAs of msda_bwd_fast_kernel (see bwd_fast_with_coarse in msda_capi.cu).
constexpr int kCoarseThreads = 256;
constexpr int kCoarseWarps = kCoarseThreads / 32;

template <int N>
struct CVec {
  float v[N];
};
template <int N>
__device__ __forceinline__ CVec<N> cld(const float* p) {
  CVec<N> r;
  if constexpr (N == 1) {
    r.v[0] = __ldg(p);
  } else if constexpr (N == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  }
  return r;
}
template <int N>
__device__ __forceinline__ CVec<N> cld(const __nv_bfloat16* p) {
  CVec<N> r;
  if constexpr (N == 1) {
    r.v[0] = __uint_as_float((unsigned)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
  } else if constexpr (N == 2) {
    const unsigned u = __ldg(reinterpret_cast<const unsigned*>(p));
    r.v[0] = __uint_as_float(u << 16); r.v[1] = __uint_as_float(u & 0xffff0000u);
  } else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    r.v[0] = __uint_as_float(u.x << 16); r.v[1] = __uint_as_float(u.x & 0xffff0000u);
    r.v[2] = __uint_as_float(u.y << 16); r.v[3] = __uint_as_float(u.y & 0xffff0000u);
  }
  return r;
}
template <int N>
__device__ __forceinline__ CVec<N> lds(const float* p) {
  CVec<N> r;
  if constexpr (N == 1) {
    r.v[0] = *p;
  } else if constexpr (N == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  }
  return r;
}
template <int N>
__device__ __forceinline__ void sts(float* p, const CVec<N>& a) {
  if constexpr (N == 1) {
    *p = a.v[0];
  } else if constexpr (N == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(a.v[0], a.v[1]);
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  }
}

struct CoarseTab {
  int H[kCoarseMaxLevels], W[kCoarseMaxLevels], start[kCoarseMaxLevels];
  int tile_off[kCoarseMaxLevels + 1];   // first pixel of resident level k inside the tile; [nres] = total pixels
  int lc, nres;
};

// shared-memory bytes of ONE staging buffer (grad_out rows + raw locations / weights of the resident levels of
// kCoarseBatch queries); host and device must agree, so it is sized for the most levels that can be resident
constexpr int kCoarseBatch = 32;   // queries per staged batch = one per lane in the record phase
__host__ __device__ constexpr int coarse_stage_bytes(int D, int elem_size, int L, int P) {
  return kCoarseBatch * (D * elem_size + (L < kCoarseMaxLevels ? L : kCoarseMaxLevels) * P * 12);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int N>
__device__ __forceinline__ CVec<N> lds_row(const float* p) { return lds<N>(p); }
template <int N>
__device__ __forceinline__ CVec<N> lds_row(const __nv_bfloat16* p) {
  CVec<N> r;
  if constexpr (N == 1) {
    r.v[0] = __uint_as_float((unsigned)*reinterpret_cast<const unsigned short*>(p) << 16);
  } else if constexpr (N == 2) {
    const unsigned u = *reinterpret_cast<const unsigned*>(p);
    r.v[0] = __uint_as_float(u << 16); r.v[1] = __uint_as_float(u & 0xffff0000u);
  } else {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    r.v[0] = __uint_as_float(u.x << 16); r.v[1] = __uint_as_float(u.x & 0xffff0000u);
    r.v[2] = __uint_as_float(u.y << 16); r.v[3] = __uint_as_float(u.y & 0xffff0000u);
  }
  return r;
}

// D in {32, 64, 128}: lane c of a warp owns channels c*D/32 .. of a pixel row.
// Dynamic shared memory: [tile: budget_bytes + one scratch row][staging buffer 0][staging buffer 1].
// The grad_out rows and the raw loc / w of the NEXT batch of 32 queries are copied in with cp.async while the
// current batch is processed (one copy per CTA, read by all 8 warps), so no global-memory latency sits on the
// read-modify-write chain.
template <int D, typename VT>
__global__ void __launch_bounds__(kCoarseThreads, 4)
msda_bwd_coarse_kernel(const VT* __restrict__ grad_out, const int64_t* __restrict__ shapes,
                       const int64_t* __restrict__ lsi, const float* __restrict__ loc, const float* __restrict__ w,
                       float* __restrict__ grad_value, int B, int S, int H, int L, int Q, int P, int budget_bytes) {
  constexpr int CPL = D / 32;
  static_assert(CPL == 1 || CPL == 2 || CPL == 4, "D must be 32, 64 or 128");
  extern __shared__ __align__(16) float tile[];
  __shared__ CoarseTab tab;

  if (threadIdx.x == 0) {
    int Hs[kFastMaxLevels], Ws[kFastMaxLevels];
    for (int l = 0; l < L; ++l) {
      Hs[l] = (int)shapes[2 * l];
      Ws[l] = (int)shapes[2 * l + 1];
    }
    const int lc = coarse_first_level(Hs, Ws, L, D, budget_bytes);
    tab.lc = lc;
    tab.nres = L - lc;
    int acc = 0;
    for (int k = 0; k < L - lc; ++k) {
      tab.H[k] = Hs[lc + k];
      tab.W[k] = Ws[lc + k];
      tab.start[k] = (int)lsi[lc + k];
      tab.tile_off[k] = acc;
      acc += Hs[lc + k] * Ws[lc + k];
    }
    tab.tile_off[L - lc] = acc;
  }
  __syncthreads();
  const int nres = tab.nres;
  if (nres == 0) return;
  const int npix = tab.tile_off[nres];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_classes = 4 * nres;             // class = resident level * 4 + (y&1)*2 + (x&1)
  const int HD = H * D;
  const int NP = L * P;
  const int RP = nres * P;                    // resident points per query row (a suffix of the row's L*P)
  const int scratch = npix * D;               // a spare pixel row after the tile: where dead records point
  float* mytile = tile + lane * CPL;

  // staging buffers
  constexpr int kGoBytes = kCoarseBatch * D * (int)sizeof(VT);
  const int stage_bytes = coarse_stage_bytes(D, (int)sizeof(VT), L, P);
  unsigned char* stage0 = reinterpret_cast<unsigned char*>(tile) + budget_bytes + D * 4;

  // this CTA's slice of the (image, head, query) space
  const long long total = (long long)B * H * Q;
  long long pos = total * blockIdx.x / gridDim.x;
  const long long end = total * (blockIdx.x + 1) / gridDim.x;
  if (pos >= end) return;

  for (int i = threadIdx.x; i < (npix + 1) * (D / 4); i += kCoarseThreads)   // + the scratch row
    reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  while (pos < end) {
    const int bh = (int)(pos / Q);
    const int q0 = (int)(pos - (long long)bh * Q);
    const int n = (int)((long long)(Q - q0) < end - pos ? (long long)(Q - q0) : end - pos);
    const int b = bh / H, h = bh - b * H;
    const int n_batches = (n + kCoarseBatch - 1) / kCoarseBatch;

    // copy batch `i` of this segment into staging buffer i & 1 (all threads; one commit group per batch)
    auto stage = [&](int i) {
      unsigned char* buf = stage0 + (size_t)(i & 1) * stage_bytes;
      const int qb = q0 + i * kCoarseBatch;
      const int cnt = (q0 + n - qb) < kCoarseBatch ? (q0 + n - qb) : kCoarseBatch;
      const long long row0 = ((long long)b * Q + qb) * H + h;       // row of query qb; +H per query
      constexpr int G16 = D * (int)sizeof(VT) / 16;
      for (int idx = threadIdx.x; idx < cnt * G16; idx += kCoarseThreads) {
        const int r = idx / G16, g = idx - r * G16;
        cp_async16(buf + (size_t)r * (D * sizeof(VT)) + g * 16,
                   reinterpret_cast<const unsigned char*>(grad_out + (row0 + (long long)r * H) * D) + g * 16);
      }
      float2* s_xy = reinterpret_cast<float2*>(buf + kGoBytes);
      float* s_w = reinterpret_cast<float*>(buf + kGoBytes + kCoarseBatch * RP * 8);
      for (int idx = threadIdx.x; idx < cnt * RP; idx += kCoarseThreads) {
        const int r = idx / RP, e = idx - r * RP;
        const long long src = (row0 + (long long)r * H) * NP + (long long)tab.lc * P + e;
        cp_async8(s_xy + idx, reinterpret_cast<const float2*>(loc) + src);
        cp_async4(s_w + idx, w + src);
      }
      cp_async_commit();
    };

    stage(0);
    for (int i = 0; i < n_batches; ++i) {
      if (i + 1 < n_batches) {
        stage(i + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();   // batch i has landed for every thread's copies
      const unsigned char* buf = stage0 + (size_t)(i & 1) * stage_bytes;
      const VT* s_go = reinterpret_cast<const VT*>(buf) + lane * CPL;
      const float2* s_xy = reinterpret_cast<const float2*>(buf + kGoBytes) + lane * RP;
      const float* s_w = reinterpret_cast<const float*>(buf + kGoBytes + kCoarseBatch * RP * 8) + lane * RP;
      const int qb = q0 + i * kCoarseBatch;
      const int cnt = (q0 + n - qb) < kCoarseBatch ? (q0 + n - qb) : kCoarseBatch;
      const bool live = lane < cnt;
      for (int cls = warp; cls < n_classes; cls += kCoarseWarps) {
        const int k = cls >> 2;                        // resident level
        const int X = cls & 1, Y = (cls >> 1) & 1;     // pixel-parity class this warp owns now
        const int Hl = tab.H[k], Wl = tab.W[k];
        const int level_off = tab.tile_off[k] * D;     // floats
        for (int p0 = 0; p0 < P; p0 += 4) {
          // ---- records of up to 4 points of my query: (tile offset, weight of the corner my class owns) ----
          int off[4];
          float wt[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            off[j] = scratch;
            wt[j] = 0.0f;
            if (live && p0 + j < P) {
              const float2 xy = s_xy[k * P + p0 + j];
              const float aw = s_w[k * P + p0 + j];
              const Cell<float> c = locate<float>(xy.x, xy.y, Hl, Wl);
              const int cx = (X ^ c.x0) & 1, cy = (Y ^ c.y0) & 1;   // which corner of the cell has my parity
              if (c.valid & (1u << (cy * 2 + cx))) {
                const float fy = cy ? c.lh : 1.0f - c.lh, fx = cx ? c.lw : 1.0f - c.lw;
                wt[j] = fy * fx * aw;
                if (wt[j] != 0.0f) off[j] = level_off + ((c.y0 + cy) * Wl + (c.x0 + cx)) * D;   // as the fast kernel: 0 adds nothing
              }
            }
          }
          // points of one query that hit the same pixel: fold the weights (same grad_out row) and retire the
          // later record to the scratch row, so the four read-modify-writes below are independent and need no
          // predicates (a dead record adds 0 * g to the scratch row)
#pragma unroll
          for (int j = 1; j < 4; ++j) {
#pragma unroll
            for (int ii = 0; ii < j; ++ii) {
              if (off[j] == off[ii] && off[j] != scratch) {
                wt[ii] += wt[j];
                wt[j] = 0.0f;
                off[j] = scratch;
              }
            }
          }
          // ---- walk the queries: lane = channel ----
          for (int r = 0; r < cnt; ++r) {
            const CVec<CPL> g = lds_row<CPL>(s_go + r * D);
            int o[4];
            float c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              o[j] = __shfl_sync(0xffffffffu, off[j], r);
              c[j] = __shfl_sync(0xffffffffu, wt[j], r);
            }
            CVec<CPL> t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = lds<CPL>(mytile + o[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int e = 0; e < CPL; ++e) t[j].v[e] = fmaf(c[j], g.v[e], t[j].v[e]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) sts<CPL>(mytile + o[j], t[j]);
          }
        }
      }
      __syncthreads();   // everyone is done with buffer i & 1 before batch i + 2 overwrites it
    }
    // ---- flush: tile += into grad_value[b, start_l + pixel, h, :], then clear ----
    {
      constexpr int LPR = D / 4;   // lanes per pixel row (float4 each)
      const int sub = threadIdx.x % LPR;
      float* gimg = grad_value + (long long)b * S * HD + h * D + sub * 4;
      for (int pix = threadIdx.x / LPR; pix < npix; pix += kCoarseThreads / LPR) {
        float4* tp = reinterpret_cast<float4*>(tile + (size_t)pix * D) + sub;
        const float4 v = *tp;
        if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) {
          int kk = 0;
#pragma unroll
          for (int m = 1; m < kCoarseMaxLevels; ++m)
            if (m < nres && pix >= tab.tile_off[m]) kk = m;
          float* g = gimg + (long long)(tab.start[kk] + pix - tab.tile_off[kk]) * HD;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                       : "memory");
          *tp = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    __syncthreads();
    pos += n;
  }
}

}  // namespace msda

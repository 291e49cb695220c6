// msda_fold.cuh -- encoder-form backward (Q == S) that pre-adds grad_value contributions ON THE SM.
//
// Why.  msda_bwd_fast_kernel leaves one 128-byte vector red per (point, corner) to L2: 78 M red rows per encoder
// layer at cfg 2, which saturates the SM's crossbar port (93 %) and the L2 reduction rate.  Queries of an 8 x 8
// pixel tile of one head sample overlapping neighbourhoods: of the 4096 contributions such a tile makes, only
// ~18 % go to DISTINCT (pixel, head) rows on model-like inputs (tools/fold_rate.py, profiles/r02a_fold_rate_*),
// whereas the 16 destinations a warp holds at any one time are 91 % distinct -- a __match_any fold cannot see
// the duplicates, a CTA-wide one can.
//
// How.  One CTA = (image, head, 8 x 8 query tile of one level), 256 threads = 32 lane groups of D/4 lanes.
//   P1  per row: records (as the fast kernel), and every non-zero corner weight is filed under its destination
//       pixel: an open-addressing hash table in shared memory maps pixel -> list head (atomicCAS claim, atomicExch
//       push); the entry (coefficient, next) sits at a static index derived from (row, point, corner).  Entries
//       move 6 bytes, never a 128-byte row.
//   P3  per row: the corner gathers, dot products, grad_sampling_loc / grad_attn_weight -- the fast kernel's
//       arithmetic, bit for bit -- with no reds at all.
//   P4  after a CTA barrier: each lane group walks the lists of its share of the used slots, accumulating
//       sum_e coef_e * grad_out[row_e] in registers (the tile's 64 grad_out rows are staged in shared memory),
//       and issues ONE red.global.add.v4.f32 row per distinct destination.
// A contribution costs the data pipe ~1.25 wavefronts (entry broadcast + one grad_out row) instead of the 2
// wavefronts + 4 crossbar sectors + L2 reduction of its own red.  Summation order inside a list follows the
// order of the atomic pushes, so grad_value keeps the run-to-run rounding freedom the red path has.
//
// A probe sequence that finds no slot (only possible when a tile touches more distinct pixels than the table
// holds) falls back to a direct red by the filing lane, so any input is handled.
#pragma once
#include "msda_fast.cuh"

namespace msda {

constexpr int kFoldThreads = 256;
constexpr int kFoldLogSlots = 11;
constexpr int kFoldSlots = 1 << kFoldLogSlots;
constexpr int kFoldMaxProbe = 16;
constexpr unsigned short kFoldNil = 0xffffu;

struct FoldHeader {
  LevelTab tab;
  int n_used;
  int pad[3];
};

// shared-memory bytes of one CTA (host and device agree)
__host__ __device__ constexpr size_t fold_smem_bytes(int D, int NQ, int NP) {
  const size_t groups = (size_t)kFoldThreads / (D / 4);
  return sizeof(FoldHeader) + (size_t)kFoldSlots * (4 + 4 + 2)      // tag | head | used
         + (size_t)NQ * D * 4                                       // grad_out rows (float)
         + (size_t)NQ * NP * 4 * (4 + 2)                            // coef | next
         + groups * (size_t)bwd_row_words(NP) * 4;           // records of the row a group is working on
}

#ifndef MSDA_FOLD_MINB
#define MSDA_FOLD_MINB 2
#endif
template <int D, typename VT, int PT, int NQ, int PRE>
__global__ void __launch_bounds__(kFoldThreads, MSDA_FOLD_MINB)
msda_bwd_fold_kernel(const VT* __restrict__ grad_out, const VT* __restrict__ value,
                     const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                     const float* __restrict__ loc, const float* __restrict__ w, float* __restrict__ grad_value,
                     float* __restrict__ grad_loc, float* __restrict__ grad_w, const FusedArgs fused, int B, int S,
                     int H, int L, int Q, int P) {
  constexpr int LANES = D / 4;
  constexpr int GROUPS = kFoldThreads / LANES;
  constexpr int PASSES = NQ / GROUPS;
  constexpr int TW = 8, TH = NQ / 8;
  constexpr bool FUSED = (PRE == kPreFused);
  static_assert(PASSES >= 1 && PASSES * GROUPS == NQ, "the tile's rows must split evenly over the lane groups");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FoldHeader* hdr = reinterpret_cast<FoldHeader*>(smem_raw);
  LevelTab* tab = &hdr->tab;
  const int NP = L * P;
  const int n4 = NP * 4;
  int* s_tag = reinterpret_cast<int*>(smem_raw + sizeof(FoldHeader));
  int* s_head = s_tag + kFoldSlots;
  float* s_go = reinterpret_cast<float*>(s_head + kFoldSlots);
  float* s_coef = s_go + NQ * D;
  float* s_recs = s_coef + NQ * n4;
  const int row_words = bwd_row_words(NP);
  unsigned short* s_next = reinterpret_cast<unsigned short*>(s_recs + GROUPS * row_words);
  unsigned short* s_used = s_next + NQ * n4;

  const int sub = (threadIdx.x & 31) % LANES;
  const int grp = threadIdx.x / LANES;
  const int HD = H * D;
  float* my = s_recs + (size_t)grp * row_words;
  float4* s_cw = reinterpret_cast<float4*>(my);
  int4* s_fin = reinterpret_cast<int4*>(my + 4 * NP);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float inv_n4 = 1.0f / (float)n4;

  load_levels<TW, TH>(tab, shapes, lsi, L);
  const int BH = B * H;
  const int64_t n_items = (int64_t)BH * tab->total_tiles;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    // ---- work item -> (tile, image, head) ----
    const unsigned it = (unsigned)item, per_img = (unsigned)H * (unsigned)tab->total_tiles;   // image slowest (L2)
    const unsigned bb = it / per_img, rem = it - bb * per_img;
    const unsigned t = rem / (unsigned)H;
    const int b = (int)bb, h = (int)(rem - t * (unsigned)H);
    int lq = 0;
#pragma unroll 1
    for (int k = 1; k < L; ++k)
      if ((int)t >= tab->tile_begin[k]) lq = k;
    const int tt = (int)t - tab->tile_begin[lq];
    const int ty = tt / tab->tiles_x[lq], tx = tt - ty * tab->tiles_x[lq];

    // ---- reset the table ----
    for (int i = threadIdx.x; i < kFoldSlots; i += kFoldThreads) {
      s_tag[i] = -1;
      s_head[i] = (int)kFoldNil;
    }
    if (threadIdx.x == 0) hdr->n_used = 0;
    __syncthreads();

    const int64_t img = (int64_t)b * S * HD;
    const VT* vimg = value + img + sub * 4;
    float* gimg = grad_value + img + sub * 4;
    asm volatile("" : "+l"(vimg), "+l"(gimg));

#pragma unroll 1
    for (int pass = 0; pass < PASSES; ++pass) {
      const int rin = pass * GROUPS + grp;                 // row inside the tile: x fastest
      const int y = ty * TH + rin / TW, x = tx * TW + rin % TW;
      const int q = tab->start[lq] + y * tab->W[lq] + x;
      const bool live = (y < tab->H[lq]) && (x < tab->W[lq]) && (q < Q);
      const int64_t row = live ? ((int64_t)b * Q + q) * H + h : 0;
      const Vec<4> go = live ? ldv<4>(grad_out + row * D + sub * 4) : vzero<4>();
      *reinterpret_cast<float4*>(s_go + rin * D + sub * 4) = make_float4(go.v[0], go.v[1], go.v[2], go.v[3]);
      __syncwarp();   // the overflow fallback of P1 reads this row

      // ---- P1: records + filing ----
      const float* rp = FUSED ? fused.ref + (row / H) * (int64_t)L * fused.ref_dim : nullptr;
      {
        const float2* lp = reinterpret_cast<const float2*>(loc + row * (int64_t)NP * 2);
        const float* wp = w + row * (int64_t)NP;
        float sm_sum = 1.0f;
        if constexpr (FUSED) sm_sum = row_softmax<LANES, 4>(wp, reinterpret_cast<float*>(s_fin) + 3, NP, sub);
        for (int pt = sub; pt < NP; pt += LANES) {
          float4 cw = zero;
          int4 fin = make_int4(0, 0, 0, 0);
          if (live) {
            float2 xy = __ldg(lp + pt);
            float aw = 0.0f;
            const int l = level_of<PT>(pt, P);
            if constexpr (FUSED) {
              aw = __fdiv_rn(__int_as_float(s_fin[pt].w), sm_sum);
              xy = fused_location(xy, rp + l * fused.ref_dim, fused.ref_dim, tab->H[l], tab->W[l], fused.num_P);
            } else {
              aw = __ldg(wp + pt);
            }
            PointRec r = make_record(xy.x, xy.y, aw, tab->H[l], tab->W[l], tab->start[l], H, h, D);
            if constexpr (FUSED) {
              if (fused.value_mask) apply_value_mask(r, fused.value_mask + (int64_t)b * S, tab->W[l], S);
            }
            cw = r.cw;
            fin = make_int4(r.oc, __float_as_int(r.lw), __float_as_int(r.lh), __float_as_int(aw));
            // file the non-zero corners under their destination pixel
            const int ebase = (rin * NP + pt) * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float c = (k == 0) ? cw.x : ((k == 1) ? cw.y : ((k == 2) ? cw.z : cw.w));
              if (c == 0.0f) continue;   // padded, gated or masked corner (or a weight that happens to be 0)
              const int pix = r.pix + (k & 1) + ((k & 2) ? tab->W[l] : 0);
              unsigned slot = ((unsigned)pix * 2654435761u) >> (32 - kFoldLogSlots);
              bool filed = false;
#pragma unroll 1
              for (int probe = 0; probe < kFoldMaxProbe; ++probe) {
                const int old = atomicCAS(&s_tag[slot], -1, pix);
                if (old == -1) s_used[atomicAdd(&hdr->n_used, 1)] = (unsigned short)slot;
                if (old == -1 || old == pix) {
                  const int prev = atomicExch(&s_head[slot], ebase + k);
                  s_next[ebase + k] = (unsigned short)prev;
                  s_coef[ebase + k] = c;
                  filed = true;
                  break;
                }
                slot = (slot + 1) & (kFoldSlots - 1);
              }
              if (!filed) {   // table full around this hash: this lane adds the whole row itself
                float* g = grad_value + img + ((int64_t)pix * H + h) * D;
                const float* gr = s_go + rin * D;
                for (int ch = 0; ch < D; ch += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + ch), "f"(c * gr[ch]),
                               "f"(c * gr[ch + 1]), "f"(c * gr[ch + 2]), "f"(c * gr[ch + 3])
                               : "memory");
              }
            }
          }
          s_cw[pt] = cw;
          s_fin[pt] = fin;
        }
      }
      __syncwarp();

      // ---- P3: gathers, dot products, grad_sampling_loc / grad_attn_weight (msda_bwd_fast_kernel's arithmetic) ----
      float* glp = grad_loc + row * (int64_t)NP * 2;
      float* gwp = grad_w + row * (int64_t)NP;
      float sm_dot = 0.0f;
      for (int c0 = 0; c0 < NP; c0 += 4) {
        float d[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int pt = c0 + j;
          if (pt < NP) {
            const int oc = s_fin[pt].x;
            const int l = level_of<PT>(pt, P);
            const int o00 = oc & ~15;
            const int o10 = o00 + tab->W[l] * HD;
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
            if (oc & 1) d0 = dotv(go, ldv<4>(vimg + o00));
            if (oc & 2) d1 = dotv(go, ldv<4>(vimg + (o00 + HD)));
            if (oc & 4) d2 = dotv(go, ldv<4>(vimg + o10));
            if (oc & 8) d3 = dotv(go, ldv<4>(vimg + (o10 + HD)));
            d[4 * j + 0] = d0;
            d[4 * j + 1] = d1;
            d[4 * j + 2] = d2;
            d[4 * j + 3] = d3;
          } else {
            d[4 * j + 0] = 0.f; d[4 * j + 1] = 0.f; d[4 * j + 2] = 0.f; d[4 * j + 3] = 0.f;
          }
        }
        transpose_reduce_4x4<LANES>(d, sub);
        const int mine = c0 + sub / (LANES / 4);
        if (live && (sub % (LANES / 4)) == 0 && mine < NP) {
          const int4 r = s_fin[mine];
          const float lw = __int_as_float(r.y), lh = __int_as_float(r.z), aw = __int_as_float(r.w);
          float g_aw = 0.0f, g_x = 0.0f, g_y = 0.0f;
          const int l = level_of<PT>(mine, P);
          if (r.x & 15) {
            const float d0 = d[0], d1 = d[1], d2 = d[2], d3 = d[3];   // corners without a validity bit were not loaded
            const float hh = 1.0f - lh, hw = 1.0f - lw;
            g_aw = hh * hw * d0 + hh * lw * d1 + lh * hw * d2 + lh * lw * d3;
            g_x = (hh * (d1 - d0) + lh * (d3 - d2)) * aw;
            g_y = (hw * (d2 - d0) + lw * (d3 - d1)) * aw;
            g_x *= (float)tab->W[l];
            g_y *= (float)tab->H[l];
          }
          if constexpr (FUSED) {
            float2 g_off;
            if (fused.ref_dim == 2) {
              g_off = make_float2(__fdiv_rn(g_x, (float)tab->W[l]), __fdiv_rn(g_y, (float)tab->H[l]));
            } else {
              const float* r4 = rp + l * 4;
              g_off = make_float2(g_x * 0.5f * __ldg(r4 + 2) * fused.inv_P, g_y * 0.5f * __ldg(r4 + 3) * fused.inv_P);
            }
            *reinterpret_cast<float2*>(glp + 2 * mine) = g_off;
            s_fin[mine].y = __float_as_int(g_aw);
            sm_dot = fmaf(g_aw, aw, sm_dot);
          } else {
            gwp[mine] = g_aw;
            *reinterpret_cast<float2*>(glp + 2 * mine) = make_float2(g_x, g_y);
          }
        }
      }
      if constexpr (FUSED) {
#pragma unroll
        for (int k = LANES / 2; k > 0; k >>= 1) sm_dot += __shfl_xor_sync(0xffffffffu, sm_dot, k);
        __syncwarp();
        if (live) {
          for (int pt = sub; pt < NP; pt += LANES) {
            const int4 r = s_fin[pt];
            gwp[pt] = __int_as_float(r.w) * (__int_as_float(r.y) - sm_dot);
          }
        }
      }
      __syncwarp();   // the group's records are rewritten by its next row
    }
    __syncthreads();

    // ---- P4: one red per distinct destination ----
    {
      const int nu = hdr->n_used;
      int k = grp;
      int cur = (int)kFoldNil, slot = -1, pix = 0;
      float4 acc = zero;
      while (true) {
        if (cur == (int)kFoldNil) {
          if (slot >= 0) {
            float* g = gimg + ((int64_t)pix * H + h) * D;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(acc.x), "f"(acc.y), "f"(acc.z),
                         "f"(acc.w)
                         : "memory");
            acc = zero;
            slot = -1;
          }
          if (k < nu) {
            slot = (int)s_used[k];
            k += GROUPS;
            cur = s_head[slot];
            pix = s_tag[slot];
          }
        }
        const bool active = cur != (int)kFoldNil;
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
          const float c = s_coef[cur];
          const int r = (int)(((float)cur + 0.5f) * inv_n4);
          const float4 g4 = *reinterpret_cast<const float4*>(s_go + r * D + sub * 4);
          acc.x = fmaf(c, g4.x, acc.x);
          acc.y = fmaf(c, g4.y, acc.y);
          acc.z = fmaf(c, g4.z, acc.z);
          acc.w = fmaf(c, g4.w, acc.w);
          cur = (int)s_next[cur];
        }
      }
    }
    __syncthreads();   // the table and the staged rows are rewritten by the next work item
  }
}

}  // namespace msda

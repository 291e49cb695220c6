/*
 * msda_oracle.c -- CPU restatement of the reference's multi-scale deformable attention.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ir_ads_b200/) may link, import or
 * execute this file.  It exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg can check the CUDA kernels against an independent scalar implementation.
 *
 * What it restates (all citations relative to /root/reference):
 *   forward   detrex/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:237-299 (thread body)
 *             and :33-84 (bilinear fetch with per-corner zero padding)
 *   backward  ms_deform_im2col_cuda.cuh:87-159 (per-corner scatter, grad_w_weight / grad_h_weight,
 *             the W_l / H_l scaling of grad_sampling_loc at :157-158) driven by the loop of
 *             :320-402
 *   indexing  ms_deform_im2col_cuda.cuh:253-269: index -> (c, h, q, b);
 *             weight offset ((b*Q+q)*H+h)*L*P, loc offset 2x that,
 *             value offset b*S*H*D + (level_start + y*W_l + x)*H*D + h*D + c
 * The same function is what the reference's Python fallback computes with F.grid_sample
 * (detrex/layers/multi_scale_deform_attn.py:96-136: bilinear, padding_mode="zeros",
 * align_corners=False, grid = 2*loc-1  =>  pixel coordinate loc*size - 0.5).
 *
 * Parity pin: tests/test_oracle.py checks this file against golden vectors produced by the
 * reference's own multi_scale_deformable_attn_pytorch + autograd (tests/golden/*.npz, made by
 * oracle/make_golden.py) -- so parity is PINNED, not assumed.
 *
 * Coordinate modes for the float path (double always follows the reference expression):
 *   MSDA_COORD_REFERENCE   h_im = loc*H - 0.5 rounded the way the reference kernel rounds it
 *                          (fp32 product, then subtraction), floor(), fractional part.
 *   MSDA_COORD_COMPENSATED the product's error term is recovered with one fma so that the
 *                          integer cell and the fractional weight are those of the EXACT value
 *                          of loc*H - 0.5 (this is what the sm_100a kernels do; see
 *                          ir_ads_b200/csrc/msda_coords.cuh -- the two must stay in lock step,
 *                          the bookkeeping test compares them bit for bit).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MSDA_COORD_REFERENCE 0
#define MSDA_COORD_COMPENSATED 1

typedef struct {
  int in_range; /* 1 if the axis passes the reference gate (coord > -1 && coord < size) */
  int low;      /* floor(coord); only meaningful when in_range */
  double frac;  /* coord - low in [0,1]                                          */
} axis_t;

/* ---- coordinate split, double: reference expression (cuh:284-288, :39-46) ---- */
static axis_t split_f64(double loc, int size) {
  axis_t a;
  double c = loc * (double)size - 0.5;
  a.in_range = (c > -1.0) && (c < (double)size);
  if (!a.in_range) { a.low = 0; a.frac = 0.0; return a; }
  double fl = floor(c);
  a.low = (int)fl;
  a.frac = c - fl;
  return a;
}

/* ---- coordinate split, float, reference rounding order ---- */
static axis_t split_f32_reference(float loc, int size) {
  axis_t a;
  volatile float prod = loc * (float)size; /* fp32 product (cuh:284: scalar_t * int) */
  float c = (float)((double)prod - 0.5);    /* "- 0.5" promotes to double, result stored as float */
  a.in_range = (c > -1.0f) && (c < (float)size);
  if (!a.in_range) { a.low = 0; a.frac = 0.0; return a; }
  float fl = floorf(c);
  a.low = (int)fl;
  a.frac = (double)(float)(c - fl);
  return a;
}

/* ---- coordinate split, float, compensated (mirrors msda_coords.cuh::split_axis) ---- */
static axis_t split_f32_compensated(float loc, int size) {
  axis_t a;
  const float sz = (float)size;
  volatile float p = loc * sz;    /* rounded product                                  */
  float e = fmaf(loc, sz, -p);    /* loc*sz == p + e exactly                          */
  volatile float av = p - 0.5f;
  volatile float bb = av - p;     /* TwoSum: p - 0.5 == av + ea exactly               */
  volatile float t1 = av - bb;
  volatile float t2 = p - t1;
  volatile float t3 = -0.5f - bb;
  volatile float ea = t2 + t3;
  float f0 = floorf(av);
  volatile float r0 = av - f0;
  volatile float tt = ea + e;
  volatile float rr = r0 + tt;
  float r = rr;
  if (r < 0.0f) { f0 -= 1.0f; volatile float t = r + 1.0f; r = t; }
  else if (r >= 1.0f) { f0 += 1.0f; volatile float t = r - 1.0f; r = t; }
  /* gate: exact coord = f0 + r, r in [0,1].  coord > -1  <=>  f0 >= 0 or (f0 == -1 and r > 0);
     coord < size <=> f0 < size.  NaN/Inf fall out as "not in range". */
  a.in_range = (f0 >= 0.0f || (f0 == -1.0f && r > 0.0f)) && (f0 < sz);
  if (!a.in_range) { a.low = 0; a.frac = 0.0; return a; }
  a.low = (int)f0;
  a.frac = (double)r;
  return a;
}

static axis_t split_axis(double loc, int size, int is_f32, int coord_mode) {
  if (!is_f32) return split_f64(loc, size);
  if (coord_mode == MSDA_COORD_REFERENCE) return split_f32_reference((float)loc, size);
  return split_f32_compensated((float)loc, size);
}

/* round-through-float helper: when is_f32, every intermediate is rounded to float the way a
   float kernel would; otherwise identity. */
static inline double rf(double x, int is_f32) { return is_f32 ? (double)(float)x : x; }

typedef struct {
  int gate;        /* both axes in range                                   */
  int y0, x0;      /* low corner                                           */
  int valid[4];    /* v1=(y0,x0) v2=(y0,x1) v3=(y1,x0) v4=(y1,x1)           */
  double lh, lw;   /* fractional parts                                      */
} cell_t;

static cell_t locate(double loc_x, double loc_y, int H, int W, int is_f32, int coord_mode) {
  cell_t c;
  memset(&c, 0, sizeof(c));
  axis_t ay = split_axis(loc_y, H, is_f32, coord_mode);
  axis_t ax = split_axis(loc_x, W, is_f32, coord_mode);
  c.gate = ay.in_range && ax.in_range;
  if (!c.gate) return c;
  c.y0 = ay.low; c.x0 = ax.low; c.lh = ay.frac; c.lw = ax.frac;
  const int y1 = c.y0 + 1, x1 = c.x0 + 1;
  /* cuh:58,64,70,76 -- the gate already guarantees y0 <= H-1 and x0 <= W-1 */
  c.valid[0] = (c.y0 >= 0 && c.x0 >= 0);
  c.valid[1] = (c.y0 >= 0 && x1 <= W - 1);
  c.valid[2] = (y1 <= H - 1 && c.x0 >= 0);
  c.valid[3] = (y1 <= H - 1 && x1 <= W - 1);
  return c;
}

/* ------------------------------------------------------------------------------------------
 * Generic driver.  value/loc/w/out are passed as double arrays by the typed wrappers below so
 * that one body serves float and double; is_f32 makes every arithmetic result round through
 * float (products and sums separately: no contraction), which is how a float kernel without
 * fma contraction behaves.  Tolerance-based tests never depend on that detail; the bit-exact
 * tests only look at the integer bookkeeping.
 * ------------------------------------------------------------------------------------------ */

#define IDX_VALUE(b, s, h, c) ((((int64_t)(b) * S + (s)) * Hh + (h)) * D + (c))

static void forward_body(const double* value, const int64_t* shapes, const int64_t* lsi,
                         const double* loc, const double* w, double* out, int B, int S, int Hh,
                         int D, int L, int Q, int P, int is_f32, int coord_mode) {
  for (int b = 0; b < B; ++b)
    for (int q = 0; q < Q; ++q)
      for (int h = 0; h < Hh; ++h) {
        const int64_t row = ((int64_t)b * Q + q) * Hh + h;
        double* o = out + row * D;
        for (int c = 0; c < D; ++c) o[c] = 0.0;
        for (int l = 0; l < L; ++l) {
          const int Hl = (int)shapes[2 * l], Wl = (int)shapes[2 * l + 1];
          const int64_t start = lsi[l];
          for (int p = 0; p < P; ++p) {
            const int64_t pt = (row * L + l) * P + p;
            const double lx = loc[2 * pt], ly = loc[2 * pt + 1], aw = w[pt];
            cell_t cell = locate(lx, ly, Hl, Wl, is_f32, coord_mode);
            if (!cell.gate) continue;
            const double lh = cell.lh, lw = cell.lw;
            const double hh = rf(1.0 - lh, is_f32), hw = rf(1.0 - lw, is_f32);
            const double cw[4] = {rf(hh * hw, is_f32), rf(hh * lw, is_f32), rf(lh * hw, is_f32),
                                  rf(lh * lw, is_f32)};
            const int ys[4] = {cell.y0, cell.y0, cell.y0 + 1, cell.y0 + 1};
            const int xs[4] = {cell.x0, cell.x0 + 1, cell.x0, cell.x0 + 1};
            for (int c = 0; c < D; ++c) {
              double val = 0.0;
              for (int k = 0; k < 4; ++k) {
                double v = 0.0;
                if (cell.valid[k]) v = value[IDX_VALUE(b, start + (int64_t)ys[k] * Wl + xs[k], h, c)];
                val = rf(val + rf(cw[k] * v, is_f32), is_f32);
              }
              o[c] = rf(o[c] + rf(val * aw, is_f32), is_f32);
            }
          }
        }
      }
}

static void backward_body(const double* grad_out, const double* value, const int64_t* shapes,
                          const int64_t* lsi, const double* loc, const double* w, double* grad_value,
                          double* grad_loc, double* grad_w, int B, int S, int Hh, int D, int L,
                          int Q, int P, int is_f32, int coord_mode) {
  const int64_t nv = (int64_t)B * S * Hh * D;
  const int64_t npts = (int64_t)B * Q * Hh * L * P;
  for (int64_t i = 0; i < nv; ++i) grad_value[i] = 0.0;
  for (int64_t i = 0; i < npts; ++i) { grad_w[i] = 0.0; grad_loc[2 * i] = 0.0; grad_loc[2 * i + 1] = 0.0; }
  for (int b = 0; b < B; ++b)
    for (int q = 0; q < Q; ++q)
      for (int h = 0; h < Hh; ++h) {
        const int64_t row = ((int64_t)b * Q + q) * Hh + h;
        const double* go = grad_out + row * D;
        for (int l = 0; l < L; ++l) {
          const int Hl = (int)shapes[2 * l], Wl = (int)shapes[2 * l + 1];
          const int64_t start = lsi[l];
          for (int p = 0; p < P; ++p) {
            const int64_t pt = (row * L + l) * P + p;
            const double lx = loc[2 * pt], ly = loc[2 * pt + 1], aw = w[pt];
            cell_t cell = locate(lx, ly, Hl, Wl, is_f32, coord_mode);
            if (!cell.gate) continue; /* cuh:369: gradients of a gated point stay zero */
            const double lh = cell.lh, lw = cell.lw;
            const double hh = rf(1.0 - lh, is_f32), hw = rf(1.0 - lw, is_f32);
            const double cw[4] = {rf(hh * hw, is_f32), rf(hh * lw, is_f32), rf(lh * hw, is_f32),
                                  rf(lh * lw, is_f32)};
            /* d val / d lw and d val / d lh coefficients per corner (cuh:119-153) */
            const double cgw[4] = {-hh, hh, -lh, lh};
            const double cgh[4] = {-hw, -lw, hw, lw};
            const int ys[4] = {cell.y0, cell.y0, cell.y0 + 1, cell.y0 + 1};
            const int xs[4] = {cell.x0, cell.x0 + 1, cell.x0, cell.x0 + 1};
            double acc_w = 0.0, acc_x = 0.0, acc_y = 0.0;
            for (int c = 0; c < D; ++c) {
              const double top = go[c];
              const double tgv = rf(top * aw, is_f32); /* cuh:117 top_grad_value */
              double val = 0.0, gww = 0.0, ghw = 0.0;
              for (int k = 0; k < 4; ++k) {
                if (!cell.valid[k]) continue;
                const int64_t vi = IDX_VALUE(b, start + (int64_t)ys[k] * Wl + xs[k], h, c);
                const double v = value[vi];
                val = rf(val + rf(cw[k] * v, is_f32), is_f32);
                gww = rf(gww + rf(cgw[k] * v, is_f32), is_f32);
                ghw = rf(ghw + rf(cgh[k] * v, is_f32), is_f32);
                grad_value[vi] = rf(grad_value[vi] + rf(cw[k] * tgv, is_f32), is_f32);
              }
              acc_w += rf(top * val, is_f32);                           /* cuh:156 */
              acc_x += rf(rf((double)Wl * gww, is_f32) * tgv, is_f32);  /* cuh:157 */
              acc_y += rf(rf((double)Hl * ghw, is_f32) * tgv, is_f32);  /* cuh:158 */
            }
            grad_w[pt] = rf(acc_w, is_f32);
            grad_loc[2 * pt] = rf(acc_x, is_f32);
            grad_loc[2 * pt + 1] = rf(acc_y, is_f32);
          }
        }
      }
}

/* ---------------------------------- typed C entry points ---------------------------------- */

static double* widen(const float* x, int64_t n) {
  double* d = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) d[i] = (double)x[i];
  return d;
}
static void narrow(const double* d, float* x, int64_t n) {
  for (int64_t i = 0; i < n; ++i) x[i] = (float)d[i];
}

int msda_oracle_forward_f64(const double* value, const int64_t* shapes, const int64_t* lsi,
                            const double* loc, const double* w, double* out, int B, int S, int H,
                            int D, int L, int Q, int P) {
  forward_body(value, shapes, lsi, loc, w, out, B, S, H, D, L, Q, P, 0, 0);
  return 0;
}

int msda_oracle_backward_f64(const double* grad_out, const double* value, const int64_t* shapes,
                             const int64_t* lsi, const double* loc, const double* w,
                             double* grad_value, double* grad_loc, double* grad_w, int B, int S,
                             int H, int D, int L, int Q, int P) {
  backward_body(grad_out, value, shapes, lsi, loc, w, grad_value, grad_loc, grad_w, B, S, H, D, L,
                Q, P, 0, 0);
  return 0;
}

int msda_oracle_forward_f32(const float* value, const int64_t* shapes, const int64_t* lsi,
                            const float* loc, const float* w, float* out, int B, int S, int H,
                            int D, int L, int Q, int P, int coord_mode) {
  const int64_t nv = (int64_t)B * S * H * D, np = (int64_t)B * Q * H * L * P,
                no = (int64_t)B * Q * H * D;
  double *dv = widen(value, nv), *dl = widen(loc, 2 * np), *dw = widen(w, np);
  double* dout = (double*)malloc(sizeof(double) * (size_t)(no > 0 ? no : 1));
  forward_body(dv, shapes, lsi, dl, dw, dout, B, S, H, D, L, Q, P, 1, coord_mode);
  narrow(dout, out, no);
  free(dv); free(dl); free(dw); free(dout);
  return 0;
}

int msda_oracle_backward_f32(const float* grad_out, const float* value, const int64_t* shapes,
                             const int64_t* lsi, const float* loc, const float* w,
                             float* grad_value, float* grad_loc, float* grad_w, int B, int S, int H,
                             int D, int L, int Q, int P, int coord_mode) {
  const int64_t nv = (int64_t)B * S * H * D, np = (int64_t)B * Q * H * L * P,
                no = (int64_t)B * Q * H * D;
  double *dgo = widen(grad_out, no), *dv = widen(value, nv), *dl = widen(loc, 2 * np),
         *dw = widen(w, np);
  double* gv = (double*)malloc(sizeof(double) * (size_t)(nv > 0 ? nv : 1));
  double* gl = (double*)malloc(sizeof(double) * (size_t)(np > 0 ? 2 * np : 1));
  double* gw = (double*)malloc(sizeof(double) * (size_t)(np > 0 ? np : 1));
  backward_body(dgo, dv, shapes, lsi, dl, dw, gv, gl, gw, B, S, H, D, L, Q, P, 1, coord_mode);
  narrow(gv, grad_value, nv); narrow(gl, grad_loc, 2 * np); narrow(gw, grad_w, np);
  free(dgo); free(dv); free(dl); free(dw); free(gv); free(gl); free(gw);
  return 0;
}

/*
 * Integer bookkeeping of every sampling point, for the bit-exact test.
 *   corner_offsets [N_pts,4] int64: flat element offset of channel 0 of each corner inside
 *       `value` ( b*S*H*D + (level_start + y*W + x)*H*D + h*D ), or -1 when the corner is
 *       zero-padded or the point is gated out.  Corner order v1..v4 as in cuh:56-80.
 *   frac [N_pts,2] float: (lw, lh) fractional weights as the float path computes them
 *       (0 for gated points).
 * is_f32 = 1 uses coord_mode on float inputs; is_f32 = 0 evaluates the float inputs in double
 * (exact floor), which is what the grid_sample oracle in fp64 does.
 */
int msda_oracle_bookkeeping(const float* loc, const int64_t* shapes, const int64_t* lsi, int B,
                            int S, int H, int D, int L, int Q, int P, int is_f32, int coord_mode,
                            int64_t* corner_offsets, float* frac) {
  for (int b = 0; b < B; ++b)
    for (int q = 0; q < Q; ++q)
      for (int h = 0; h < H; ++h)
        for (int l = 0; l < L; ++l) {
          const int Hl = (int)shapes[2 * l], Wl = (int)shapes[2 * l + 1];
          const int64_t start = lsi[l];
          for (int p = 0; p < P; ++p) {
            const int64_t pt = ((((int64_t)b * Q + q) * H + h) * L + l) * P + p;
            cell_t cell = locate((double)loc[2 * pt], (double)loc[2 * pt + 1], Hl, Wl, is_f32, coord_mode);
            const int ys[4] = {cell.y0, cell.y0, cell.y0 + 1, cell.y0 + 1};
            const int xs[4] = {cell.x0, cell.x0 + 1, cell.x0, cell.x0 + 1};
            for (int k = 0; k < 4; ++k) {
              int64_t off = -1;
              if (cell.gate && cell.valid[k])
                off = ((int64_t)b * S + start + (int64_t)ys[k] * Wl + xs[k]) * H * D + (int64_t)h * D;
              corner_offsets[4 * pt + k] = off;
            }
            frac[2 * pt] = cell.gate ? (float)cell.lw : 0.0f;
            frac[2 * pt + 1] = cell.gate ? (float)cell.lh : 0.0f;
          }
        }
  return 0;
}

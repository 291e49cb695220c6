// Microbenchmark: issue rates of the instructions the deterministic (fixed-point) backward is made of, per SM:
//   0  F2I.S64.F32        cvt.rni.s64.f32   (what __float2ll_rn compiles to: 16 per entry and lane in det_cell_reduce_kernel)
//   1  F2I.S32.F32        cvt.rni.s32.f32
//   2  F2F.F64.F32        cvt.f64.f32
//   3  DFMA               fma.rn.f64
//   4  FFMA               fma.rn.f32        (reference point: one warp instruction per clock and SM sub-partition)
//   5  IMAD.WIDE          mad.wide.s32      (32 x 32 + 64 -> 64)
// 32 warps per SM issue independent chains back to back; prints warp instructions per clock and SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/cvt_rates tools/microbench/cvt_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cycles, int iters, float seed) {
  float x[8];
  double d[8];
  long long q[8];
  int n[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    x[u] = seed + threadIdx.x * 0.001f + u;
    d[u] = x[u];
    q[u] = u;
    n[u] = threadIdx.x + u;
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE <= 2) x[u] = __int_as_float(__float_as_int(x[u]) + 0x100);   // a new operand every time: ptxas merges
      if (MODE == 0) {                                                          // conversions of one value otherwise
        long long r;
        asm volatile("cvt.rni.s64.f32 %0, %1;" : "=l"(r) : "f"(x[u]));
        q[u] ^= r;
      } else if (MODE == 1) {
        int r;
        asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(r) : "f"(x[u]));
        n[u] ^= r;
      } else if (MODE == 2) {
        double r;
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(r) : "f"(x[u]));
        q[u] ^= __double_as_longlong(r);
      } else if (MODE == 3) {
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[u]) : "d"(1.0000001), "d"(0.5));
      } else if (MODE == 4) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[u]) : "f"(1.0000001f), "f"(0.5f));
      } else {
        asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(q[u]) : "r"(n[u]), "r"(it));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) acc += x[u] + (float)d[u] + (float)q[u] + (float)n[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* what, float* out, long long* cyc) {
  const int iters = 4000, threads = 1024;
  k<MODE><<<148, threads>>>(out, cyc, iters, 1.5f);
  k<MODE><<<148, threads>>>(out, cyc, iters, 1.5f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double inst_per_sm = (double)iters * 8 * (threads / 32);
  // MODE 0..2 carry an integer add and one or two XORs per conversion (ALU instructions, another pipe)
  printf("%-22s %6.3f warp instructions / clk / SM   (%5.1f lanes / clk / SM)\n", what, inst_per_sm / (double)h[0],
         32.0 * inst_per_sm / (double)h[0]);
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run<0>("F2I.S64.F32", out, cyc);
  run<1>("F2I.S32.F32", out, cyc);
  run<2>("F2F.F64.F32", out, cyc);
  run<3>("DFMA", out, cyc);
  run<4>("FFMA", out, cyc);
  run<5>("IMAD.WIDE (s32, +s64)", out, cyc);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}

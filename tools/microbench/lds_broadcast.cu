// Microbenchmark: what does a broadcast LDS.128 cost the shared-memory data path?  The forward kernel reads a
// 16-byte record per (row, point) that all 8 lanes of the row need; a warp holds 4 rows.
//   mode 0  lanes 8r..8r+7 read record r            (contiguous lane groups: the kernels' mapping)
//   mode 1  lane l reads record l % 4               (rows interleaved across the quarter-warps)
//   mode 2  all 32 lanes read the same record
//   mode 3  every lane reads its own 16 bytes       (512 distinct bytes: 4 wavefronts by construction)
//   mode 4  as mode 0 with LDS.32 x 4 instead of one LDS.128
// Prints cycles per warp-level load instruction with 8 warps per SM sub-partition issuing back to back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lds_broadcast tools/microbench/lds_broadcast.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cycles, int iters) {
  __shared__ float4 rec[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) rec[i] = make_float4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int idx;
  if (MODE == 0 || MODE == 4) idx = lane >> 3;
  else if (MODE == 1) idx = lane & 3;
  else if (MODE == 2) idx = 0;
  else idx = lane;
  idx += warp * 32;                       // every warp its own 512-byte window
  float4 acc = make_float4(0, 0, 0, 0);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = (idx + u * 32 + (it & 1) * 4) & 1023;    // distinct addresses, so no load is merged with another
      if (MODE == 4) {
        const float* p = reinterpret_cast<const float*>(rec + j);
        volatile const float* vp = p;
        acc.x += vp[0]; acc.y += vp[1]; acc.z += vp[2]; acc.w += vp[3];
      } else {
        float4 v;
        const unsigned a = (unsigned)__cvta_generic_to_shared(rec + j);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  __syncthreads();                         // the slowest warp ends the measurement (the scheduler favours old warps)
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MODE>
void run(const char* what, float* out, long long* cyc) {
  const int iters = 2000, threads = 1024;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double loads_per_sm = (double)iters * 16 * (threads / 32) * (MODE == 4 ? 4 : 1);
  printf("%-58s %6.2f cycles per warp load instruction (SM-wide)\n", what, (double)h[0] / loads_per_sm);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run<0>("LDS.128, 4 records, contiguous 8-lane groups", out, cyc);
  run<1>("LDS.128, 4 records, rows interleaved (lane % 4)", out, cyc);
  run<2>("LDS.128, one record for the whole warp", out, cyc);
  run<3>("LDS.128, 32 distinct records (512 B)", out, cyc);
  run<4>("LDS.32 x4, 4 records, contiguous 8-lane groups (per LDS.32)", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

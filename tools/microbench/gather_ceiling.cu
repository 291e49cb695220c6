// Microbenchmark: how fast can one B200 GATHER 128-byte fp32 rows (one LDG.128 per lane, 8 lanes per row --
// the access pattern of msda_fwd_fast_kernel) from an L2-resident buffer, as a function of locality?
// Each 8-lane group reads `rows_per_group` rows picked at random from a window that slides with the group id,
// accumulates them (4 FFMA per row) and writes one row.  4 independent rows in flight per group (as the kernel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gather_ceiling tools/microbench/gather_ceiling.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <initializer_list>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// MODE 0: __ldg (ld.global.nc)   1: ld.global.nc.L1::no_allocate   2: ld.global.nc.L1::evict_last   3: ld.global.cg (L2 only)
template <int MODE>
__device__ __forceinline__ float4 load_row(const float* p) {
  float4 r;
  if (MODE == 0) return __ldg(reinterpret_cast<const float4*>(p));
  if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256, 6)
k_gather(const float* __restrict__ g, float* __restrict__ out, uint32_t n_lines, uint32_t rows_per_group,
         uint32_t n_groups, uint32_t window, uint32_t seed) {
  const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const uint32_t sub = threadIdx.x & 7;
  const uint32_t base = (uint32_t)((uint64_t)gid * (n_lines - window) / n_groups);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t i = 0; i < rows_per_group; i += 4) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t line = base + hash32(gid * 977u + (i + k) * 131071u + seed) % window;
      v[k] = load_row<MODE>(g + (size_t)line * 32 + sub * 4);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc.x = fmaf(0.25f, v[k].x, acc.x); acc.y = fmaf(0.25f, v[k].y, acc.y);
      acc.z = fmaf(0.25f, v[k].z, acc.z); acc.w = fmaf(0.25f, v[k].w, acc.w);
    }
  }
  *reinterpret_cast<float4*>(out + (size_t)gid * 32 + sub * 4) = acc;
}

int main() {
  const uint32_t n_lines = 1422272;           // cfg2: B*S*H lines of 128 B (182 MB)
  const uint32_t rows_total = 91025408;       // corner rows of one forward
  const uint32_t rows_per_group = 64;         // 16 points x 4 corners
  const uint32_t n_groups = rows_total / rows_per_group;
  const int threads = 256;
  const int grid = n_groups / (threads / 8);
  float *g, *out;
  cudaMalloc(&g, (size_t)n_lines * 128);
  cudaMalloc(&out, (size_t)n_groups * 128);
  cudaMemset(g, 0, (size_t)n_lines * 128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 4; ++mode)
  for (uint32_t window : {n_lines - 1, 177784u, 16384u, 2048u, 512u, 64u}) {
    if (mode > 0 && window != 177784u && window != 2048u) continue;
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_gather<0><<<grid, threads>>>(g, out, n_lines, rows_per_group, n_groups, window, 1u + rep);
      if (mode == 1) k_gather<1><<<grid, threads>>>(g, out, n_lines, rows_per_group, n_groups, window, 1u + rep);
      if (mode == 2) k_gather<2><<<grid, threads>>>(g, out, n_lines, rows_per_group, n_groups, window, 1u + rep);
      if (mode == 3) k_gather<3><<<grid, threads>>>(g, out, n_lines, rows_per_group, n_groups, window, 1u + rep);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("mode %d window %8u lines (%7.2f MB)  %7.3f ms  %7.2f Grows/s  %6.2f TB/s gathered\n", mode, window,
           window * 128 / 1e6, best, rows_total / best / 1e6, rows_total * 128.0 / best / 1e9);
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}

// Microbenchmark: how fast can one B200 add 128-byte fp32 rows into random lines of a large buffer?
//   A  red.global.add.v4.f32 by 8-lane groups (what msda_bwd_fast_kernel does)
//   B  rows staged in shared memory, then cp.reduce.async.bulk (TMA reduce) of 128 B per row
//   C  like B but 512 B per bulk op (4 consecutive lines: upper bound for a tile-flush style scatter)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/red_vs_tma tools/microbench/red_vs_tma.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

__device__ __forceinline__ uint32_t pick_line(uint32_t gid, uint32_t i, uint32_t seed, uint32_t n_lines, uint32_t n_groups, uint32_t window) {
  // a window of `window` lines sliding over the buffer with the group id: L2-resident like the real backward
  const uint32_t base = (uint32_t)((uint64_t)gid * (n_lines - window) / n_groups);
  return base + hash32(gid * 977u + i * 131071u + seed) % window;
}

__global__ void k_red(float* g, uint32_t n_lines, uint32_t rows_per_group, uint32_t seed, uint32_t n_groups, uint32_t window) {
  const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;   // 8-lane group id
  const uint32_t sub = threadIdx.x & 7;
  for (uint32_t i = 0; i < rows_per_group; ++i) {
    const uint32_t line = pick_line(gid, i, seed, n_lines, n_groups, window);
    float* p = g + (size_t)line * 32 + sub * 4;
    const float v = (float)(i + 1);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
  }
}

template <int BYTES>
__global__ void k_tma(float* g, uint32_t n_lines, uint32_t rows_per_group, uint32_t seed, uint32_t n_groups, uint32_t window) {
  extern __shared__ __align__(128) float smem[];
  const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const uint32_t lgid = threadIdx.x >> 3;              // group inside the CTA
  const uint32_t sub = threadIdx.x & 7;
  float* my = smem + lgid * (BYTES / 4);               // BYTES of staging per group
  for (uint32_t i = 0; i < rows_per_group; ++i) {
    uint32_t line = pick_line(gid, i, seed, n_lines, n_groups, window);
    if (BYTES > 128) line = (line / (BYTES / 128)) * (BYTES / 128);
    const float v = (float)(i + 1);
#pragma unroll
    for (int k = 0; k < BYTES / 128; ++k)
      *reinterpret_cast<float4*>(my + k * 32 + sub * 4) = make_float4(v, v, v, v);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (sub == 0) {
      const uint32_t s = (uint32_t)__cvta_generic_to_shared(my);
      float* dst = g + (size_t)line * 32;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(s),
                   "n"(BYTES)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
}

int main() {
  const uint32_t n_lines = 1422272;          // cfg2: B*S*H lines of 128 B
  const uint32_t rows_total = 91025408 / 4;  // a quarter of one backward's corner contributions
  const int threads = 256, groups_per_cta = threads / 8;
  const uint32_t rows_per_group = 16;
  const uint32_t n_groups = rows_total / rows_per_group;
  const int grid = n_groups / groups_per_cta;
  float* g;
  cudaMalloc(&g, (size_t)n_lines * 128);
  cudaMemset(g, 0, (size_t)n_lines * 128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, float ms, double bytes_per_row) {
    printf("%-28s %8.3f ms  %7.2f Grows/s  %7.2f TB/s payload\n", name, ms, rows_total / ms / 1e6,
           rows_total * bytes_per_row / ms / 1e9);
  };
  for (uint32_t window : {n_lines - 1, 262144u, 16384u, 2048u})
  for (int rep = 0; rep < 2; ++rep) {
    if (rep) printf("-- window %u lines (%.1f MB)\n", window, window * 128 / 1e6);
    cudaEventRecord(e0);
    k_red<<<grid, threads>>>(g, n_lines, rows_per_group, 1u + rep, n_groups, window);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep) report("A red.v4.f32 128B/row", ms, 128);
    cudaEventRecord(e0);
    k_tma<128><<<grid, threads, groups_per_cta * 128>>>(g, n_lines, rows_per_group, 7u + rep, n_groups, window);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) report("B tma reduce 128B/row", ms, 128);
    cudaEventRecord(e0);
    k_tma<512><<<grid, threads, groups_per_cta * 512>>>(g, n_lines, rows_per_group, 13u + rep, n_groups, window);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) report("C tma reduce 512B/op", ms, 512);
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  // checksum so nothing is optimised away
  float* h = (float*)malloc(128 * 4);
  cudaMemcpy(h, g, 128 * 4, cudaMemcpyDeviceToHost);
  printf("g[0]=%f\n", h[0]);
  return err != cudaSuccess;
}

// Microbenchmark: gather ceiling for 64-byte rows (bf16 value, D=32), L2-resident window.
//   A  8 lanes x LDG.64  per row, 4 rows per warp instruction   (what the bf16 kernels do today)
//   B  4 lanes x LDG.128 per row, 8 rows per warp instruction   (8 channels per lane)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gather_bf16 tools/microbench/gather_bf16.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int LANES>   // lanes per 64-byte row: 8 (uint2 loads) or 4 (uint4 loads)
__global__ void __launch_bounds__(256, 6)
k_gather(const unsigned char* __restrict__ g, float* __restrict__ out, uint32_t n_lines, uint32_t rows_per_group,
         uint32_t n_groups, uint32_t window, uint32_t seed) {
  const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const uint32_t sub = threadIdx.x % LANES;
  if (gid >= n_groups) return;
  const uint32_t base = (uint32_t)((uint64_t)gid * (n_lines - window) / n_groups);
  float acc = 0.f;
  for (uint32_t i = 0; i < rows_per_group; i += 4) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t line = base + hash32(gid * 977u + (i + k) * 131071u + seed) % window;
      const unsigned char* p = g + (size_t)line * 64;
      if (LANES == 8) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p) + sub);
        v[k] = make_uint4(t.x, t.y, 0, 0);
      } else {
        v[k] = __ldg(reinterpret_cast<const uint4*>(p) + sub);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += __uint_as_float(v[k].x << 16) + __uint_as_float(v[k].y << 16) +
                                      __uint_as_float(v[k].z << 16) + __uint_as_float(v[k].w << 16);
  }
  out[(size_t)gid * LANES + sub] = acc;
}

int main() {
  const uint32_t n_lines = 1422272;       // B*S*H rows of 64 B (91 MB)
  const uint32_t rows_total = 91025408;
  const uint32_t rows_per_group = 64;
  const uint32_t n_groups = rows_total / rows_per_group;
  unsigned char* g; float* out;
  cudaMalloc(&g, (size_t)n_lines * 64);
  cudaMalloc(&out, (size_t)n_groups * 8 * 4);
  cudaMemset(g, 0, (size_t)n_lines * 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (uint32_t window : {177784u, 2048u, 64u}) {
    for (int variant = 0; variant < 2; ++variant) {
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (variant == 0) k_gather<8><<<n_groups / 32, 256>>>(g, out, n_lines, rows_per_group, n_groups, window, 1u + rep);
        else k_gather<4><<<n_groups / 64, 256>>>(g, out, n_lines, rows_per_group, n_groups, window, 1u + rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      printf("window %7u lines  %s  %7.3f ms  %7.2f Grows/s  %6.2f TB/s\n", window,
             variant == 0 ? "A 8 lanes x LDG.64 " : "B 4 lanes x LDG.128", best, rows_total / best / 1e6,
             rows_total * 64.0 / best / 1e9);
    }
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}

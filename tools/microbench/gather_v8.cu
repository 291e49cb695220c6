// Microbenchmark: 128-byte fp32 rows gathered with 256-bit lane loads (ld.global.nc.v8.f32, SASS LDG.E.ENL2.256,
// new on sm_100): 4 lanes per row, 8 rows per warp instruction, against the 128-bit form (8 lanes per row, 4 rows per
// instruction) the kernels use.  Same problem as gather_ceiling.cu: 91 M corner rows of one cfg2 layer, random picks
// from a sliding window, 4 independent rows in flight per lane group.
// Also: 16 lanes x 64-bit per row (2 rows per instruction) and 32 lanes x 32-bit (1 row per instruction), to see
// whether the L1 path charges per instruction, per line or per byte.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gather_v8 tools/microbench/gather_v8.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <initializer_list>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int N> struct V { float v[N]; };
template <int N> __device__ __forceinline__ V<N> ldn(const float* p);
template <> __device__ __forceinline__ V<8> ldn<8>(const float* p) {
  V<8> r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
template <> __device__ __forceinline__ V<4> ldn<4>(const float* p) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  V<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <> __device__ __forceinline__ V<2> ldn<2>(const float* p) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(p));
  V<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <> __device__ __forceinline__ V<1> ldn<1>(const float* p) { V<1> r; r.v[0] = __ldg(p); return r; }

// CPL channels per lane -> LANES = 32 / CPL lanes per 128-byte row
template <int CPL, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_gather(const float* __restrict__ g, float* __restrict__ out, uint32_t n_lines, uint32_t rows_per_group,
         uint32_t n_groups, uint32_t window, uint32_t seed) {
  constexpr int LANES = 32 / CPL;
  const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const uint32_t sub = threadIdx.x % LANES;
  if (gid >= n_groups) return;
  const uint32_t base = (uint32_t)((uint64_t)gid * (n_lines - window) / n_groups);
  V<CPL> acc;
#pragma unroll
  for (int c = 0; c < CPL; ++c) acc.v[c] = 0.f;
  for (uint32_t i = 0; i < rows_per_group; i += 4) {
    V<CPL> v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // cheap index (3 integer instructions): the benchmark must not be bound by its own address arithmetic
      const uint32_t line = base + (((gid + seed) * 2654435761u + (i + k) * 40503u) >> 7 & (window - 1));
      v[k] = ldn<CPL>(g + (size_t)line * 32 + sub * CPL);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc.v[c] = fmaf(0.25f, v[k].v[c], acc.v[c]);
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c) out[(size_t)gid * 32 + sub * CPL + c] = acc.v[c];
}

template <int CPL, int MINB>
float run(const float* g, float* out, uint32_t n_lines, uint32_t rpg, uint32_t n_groups, uint32_t window) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int gpb = 256 / (32 / CPL);
  const int grid = (n_groups + gpb - 1) / gpb;
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_gather<CPL, MINB><<<grid, 256>>>(g, out, n_lines, rpg, n_groups, window, 1u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  return best;
}

int main() {
  const uint32_t n_lines = 1422272, rows_total = 91025408, rpg = 64, n_groups = rows_total / rpg;
  float *g, *out;
  cudaMalloc(&g, (size_t)n_lines * 128);
  cudaMalloc(&out, (size_t)n_groups * 128);
  cudaMemset(g, 0, (size_t)n_lines * 128);
  for (uint32_t window : {131072u, 2048u, 64u}) {   // powers of two (the index is masked, not reduced modulo)
    const float t8 = run<8, 4>(g, out, n_lines, rpg, n_groups, window);
    const float t8b = run<8, 6>(g, out, n_lines, rpg, n_groups, window);
    const float t4 = run<4, 6>(g, out, n_lines, rpg, n_groups, window);
    const float t2 = run<2, 6>(g, out, n_lines, rpg, n_groups, window);
    const float t1 = run<1, 6>(g, out, n_lines, rpg, n_groups, window);
    printf("window %7u lines: v8 (4 lanes/row, LDG.256) %.3f ms [minb 6: %.3f] | v4 (8 lanes, LDG.128) %.3f ms | v2 (16 lanes, LDG.64) %.3f ms | v1 (32 lanes, LDG.32) %.3f ms\n",
           window, t8, t8b, t4, t2, t1);
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}

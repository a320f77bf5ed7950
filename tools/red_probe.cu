// red_probe.cu -- how fast can a B200 retire `red.global.add.v4.f32` into a small table, and does it matter how the
// lanes of a warp are arranged?  (design input for the EMA statistics scatter: N*D/4 vector reductions into K rows)
//   G lanes share a "pixel" (one table row) and add adjacent float4s: G = 1 -> 32 rows per instruction,
//   G = 2 -> 16 rows x 32 B, G = 8 -> 4 rows x 128 B (one full line per pixel per instruction)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_probe tools/red_probe.cu && tools/red_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) probe(float* table, int K, int D, int R, int G, int iters, unsigned seed) {
  const int lane = threadIdx.x & 31;
  const int grp = (blockIdx.x * 512 + threadIdx.x) / G;        // "pixel stream" shared by G adjacent lanes
  unsigned x = seed ^ (grp * 9781u + 1u);
  float* base = table + (size_t)(blockIdx.x % R) * K * D;
  const int nq = D / 4;
  for (int it = 0; it < iters * G; ++it) {                     // same number of lane-ops per thread for every G
    x = x * 1664525u + 1013904223u;
    const int k = (x >> 8) % K;
    float* row = base + (size_t)k * D;
    for (int j = lane % G; j < nq; j += G) {
      asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(row + 4 * j), "f"(1.0f) : "memory");
    }
  }
}

int main() {
  const int Ks[] = {64, 512};
  const int Ds[] = {64, 256};
  float* table;
  cudaMalloc(&table, (size_t)32 * 512 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int K : Ks)
    for (int D : Ds)
      for (int R : {1, 8, 32})
        for (int G : {1, 2, 4, 8}) {
          cudaMemset(table, 0, (size_t)R * K * D * 4);
          const int iters = 64 * 64 / D * 4;
          probe<<<148, 512, 0>>>(table, K, D, R, G, 2, 1u);
          cudaEventRecord(e0);
          probe<<<148, 512, 0>>>(table, K, D, R, G, iters, 7u);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          const double ops = 148.0 * 512 * iters * (D / 4);
          printf("K=%4d D=%3d R=%2d G=%d : %.3f ms, %.1f G red.v4 lane-ops/s, %.3f cyc/lane-op/SM @1.9GHz\n", K, D, R, G, ms,
                 ops / ms * 1e-6, ms * 1e-3 * 1.9e9 * 148 / ops);
        }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

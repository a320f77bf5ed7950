// red_probe.cu -- how fast can a B200 retire `red.global.add.v4.f32` into a small table?
// (design input for the EMA statistics scatter of vq_assign_tc: N*D/4 vector reductions into K rows of D floats)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_probe tools/red_probe.cu && tools/red_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) probe(float* table, int K, int D, int R, int iters, unsigned seed) {
  // every thread plays "pixel": picks a pseudo-random code per iteration, adds D/4 float4s (one per loop step)
  unsigned x = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u + 1u);
  float* base = table + (size_t)(blockIdx.x % R) * K * D;
  const int nq = D / 4;
  for (int it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    const int k = (x >> 8) % K;
    float* row = base + (size_t)k * D;
    for (int j = 0; j < nq; ++j) {
      asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(row + 4 * j), "f"(1.0f) : "memory");
    }
  }
}

int main() {
  const int Ks[] = {64, 512, 4096};
  const int Ds[] = {64, 256};
  const int Rs[] = {1, 4, 16};
  float* table;
  cudaMalloc(&table, (size_t)16 * 4096 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int K : Ks)
    for (int D : Ds)
      for (int R : Rs) {
        if ((size_t)R * K * D * 4 > (size_t)16 * 4096 * 256 * 4) continue;
        cudaMemset(table, 0, (size_t)R * K * D * 4);
        const int iters = 64 * 64 / D * 4;
        probe<<<148, 512, 0>>>(table, K, D, R, 4, 1u);
        cudaEventRecord(e0);
        probe<<<148, 512, 0>>>(table, K, D, R, iters, 7u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = 148.0 * 512 * iters * (D / 4);
        printf("K=%4d D=%3d R=%2d : %.3f ms, %.2f G red.v4/s, %.1f GB/s of payload, %.3f cyc/lane-op/SM @1.9GHz\n", K, D, R, ms,
               ops / ms * 1e-6, ops * 16 / ms * 1e-6, ms * 1e-3 * 1.9e9 * 148 / ops);
      }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

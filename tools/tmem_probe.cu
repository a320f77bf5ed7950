// tmem_probe2.cu -- TMEM read bandwidth of one SM by tcgen05.ld width and loads in flight per wait (generated)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ldx8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldx16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldx32(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldx64(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int W, int F>   // W columns per load, F loads in flight per wait
__device__ __forceinline__ uint32_t group(uint32_t taddr) {
  uint32_t r[W * F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    if (W == 8) ldx8(taddr + f * W, r + f * W);
    if (W == 16) ldx16(taddr + f * W, r + f * W);
    if (W == 32) ldx32(taddr + f * W, r + f * W);
    if (W == 64) ldx64(taddr + f * W, r + f * W);
  }
  ldwait();
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < W * F; ++i) x ^= r[i];
  return x;
}
template <int W, int F>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  const int nw = blockDim.x >> 5;
  const int per = 512 / (nw / 4);
  const int c0 = (warp >> 2) * per;
  uint32_t x = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    for (int c = 0; c < per; c += W * F) x ^= group<W, F>(base + c0 + c);
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (x == 0x12345678u) sink[0] = x;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}
template <int W, int F>
void run(long long* cyc, uint32_t* sink) {
  const int iters = 200;
  for (int nw : {4, 8, 16}) {
    if (512 / (nw / 4) < W * F) continue;
    for (int rep = 0; rep < 2; ++rep) probe<W, F><<<148, nw * 32>>>(iters, cyc, sink);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double bytes = (double)iters * 128 * 512 * 4;
    const double loads_per_warp = (double)iters * (512 / (nw / 4)) / W;
    printf("x%-2d, %d in flight, warps %2d : %6.1f B/clk/SM, 256 KB in %5.0f cycles, %5.0f cycles per load per warp\n", W, F, nw,
           bytes / avg, avg / iters, avg / loads_per_warp);
  }
}
int main() {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  run<8, 1>(cyc, sink); run<8, 2>(cyc, sink); run<8, 4>(cyc, sink);
  run<16, 1>(cyc, sink); run<16, 2>(cyc, sink); run<16, 4>(cyc, sink);
  run<32, 1>(cyc, sink); run<32, 2>(cyc, sink);
  run<64, 1>(cyc, sink);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

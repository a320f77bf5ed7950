// umma_rate.cu -- issue-rate probe of tcgen05.mma (debug tool, not product): cycles per M=128 x N=256 instruction when one
// thread issues a long back-to-back chain on operands that already sit in shared memory (contents irrelevant).
//   variant 0: kind::tf32, A MN-major (128B swizzle / 32B atom: the search kernels' z operand), B K-major   (K = 8)
//   variant 1: kind::tf32, A K-major, B K-major                                                              (K = 8)
//   variant 2: kind::f16 (bf16), A K-major, B K-major                                                        (K = 16)
//   variant 3: as 0 with cta_group::2 is not probed here (needs a cluster launch)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int variant, int n_mma, int n_cols, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 98304);
  uint32_t* slot = (uint32_t*)(smem + 98304 + 64);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (float)(i & 1023);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
    uint32_t idesc = 0;
    idesc |= 1u << 4;                                   // D = f32
    if (variant == 2) { idesc |= 1u << 7; idesc |= 1u << 10; }      // bf16 x bf16
    else { idesc |= 2u << 7; idesc |= 2u << 10; }       // tf32 x tf32
    if (variant == 0) idesc |= 1u << 15;                // A MN-major
    idesc |= (uint32_t)(n_cols >> 3) << 17;
    idesc |= (uint32_t)(128 >> 4) << 24;
    const uint64_t ad = variant == 0 ? make_desc(sa, 4096, 512, 1) : make_desc(sa, 16, 1024, 2);
    const uint64_t bd = make_desc(sb, 16, 1024, 2);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tmem + (uint32_t)(i & 1) * 256;     // alternate two accumulators like the kernels
      const uint32_t acc = i > 1;
      // four k-steps inside one 128-byte swizzle row, as the kernels issue them
      const uint64_t a2 = ad + (uint64_t)((variant == 0 ? (i & 3) * 1024 : (i & 3) * 32) >> 4);
      const uint64_t b2 = bd + (uint64_t)(((i & 3) * 32) >> 4);
      if (variant == 2) {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(d), "l"(a2), "l"(b2), "r"(idesc), "r"(acc) : "memory");
      } else {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(d), "l"(a2), "l"(b2), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 1024);
  const char* names[3] = {"tf32 A MN-major (z operand of the search kernels), B K-major, K=8",
                          "tf32 A K-major, B K-major, K=8", "bf16 A K-major, B K-major, K=16"};
  for (int grid : {1, 148}) {
    for (int v = 0; v < 3; ++v) {
      for (int ncols : {256, 128}) {
        const int n = 4096;
        rate_kernel<<<grid, 128, 100 * 1024 + 1024>>>(v, n, ncols, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const int kk = v == 2 ? 16 : 8;
        const double cyc = (double)mx / n;
        printf("grid %3d  N=%3d  %-68s : %7.1f cycles per MMA  (%.0f MAC/clk/SM)\n", grid, ncols, names[v], cyc,
               128.0 * ncols * kk / cyc);
      }
    }
  }
  return 0;
}

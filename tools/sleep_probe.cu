// sleep_probe.cu -- how long does nanosleep.u32 really suspend a warp on B200?  (design input for the waits of vq_assign_tc)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned ns, long long* out) {
  long long acc = 0, mn = 1 << 30, mx = 0;
  for (int i = 0; i < 64; ++i) {
    const long long t0 = clock64();
    asm volatile("nanosleep.u32 %0;" ::"r"(ns) : "memory");
    const long long d = clock64() - t0;
    acc += d; mn = d < mn ? d : mn; mx = d > mx ? d : mx;
  }
  if (threadIdx.x == 0) { out[0] = acc / 64; out[1] = mn; out[2] = mx; }
}
int main() {
  long long* d; cudaMalloc(&d, 24);
  for (unsigned ns : {0u, 16u, 32u, 64u, 128u, 256u, 512u, 1024u, 2048u, 4096u}) {
    probe<<<1, 32>>>(ns, d);
    long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("nanosleep %4u ns : avg %6lld cycles, min %6lld, max %6lld\n", ns, h[0], h[1], h[2]);
  }
  return 0;
}

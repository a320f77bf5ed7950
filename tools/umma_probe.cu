// umma_probe.cu -- stand-alone probe of tcgen05.mma kind::tf32 operand layouts (debug tool, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe <variant>
// D[128 x N] = A[128 x K] * B[N x K]^T with K = 8*ksteps, operands written to shared memory by threads.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

struct Cfg {
  int a_mn;        // 1: A stored MN-major (m contiguous), 0: K-major
  int a_layout;    // descriptor layout code for A (2 = SW128, 0 = none, 1 = SW128_BASE32B)
  int a_lbo, a_sbo;
  int a_kstep;     // descriptor start advance per k-step (bytes)
  int b_layout, b_lbo, b_sbo, b_kstep;
  int N, ksteps;
  int a_swz;       // swizzle applied by the writer for A: 0 none, 3 = 128B (xor 16B chunk with row&7)
  int b_swz;
  int b_interleave; // B written in no-swizzle interleaved core-matrix layout
};

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

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* Dout, Cfg c) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 64 KB region for A
  uint8_t* sb = smem + 65536;         // 64 KB region for B
  uint64_t* bar = (uint64_t*)(smem + 131072);
  uint32_t* slot = (uint32_t*)(smem + 131072 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = 8 * c.ksteps;
  for (int i = tid; i < 32768; i += 128) ((float*)smem)[i] = 0.f;
  __syncthreads();
  // ---- A (128 x K) --------------------------------------------------------------------------
  for (int i = tid; i < 128 * K; i += 128) {
    const int m = i % 128, k = i / 128;
    uint32_t off;
    if (c.a_mn) {
      // MN-major: chunk(k/32) -> group(m/32) -> row(k%32) of 128 B -> 16-byte column ((m%32)/4 ^ row&7)
      const int row = k & 31;
      int col = (m & 31) >> 2;
      if (c.a_swz == 3) col ^= (row & 7);
      if (c.a_swz == 2) col ^= ((row & 3) << 1);     // 128B span, 32-byte atoms: 32 B chunk index ^= row % 4
      off = (k >> 5) * 16384 + (m >> 5) * 4096 + row * 128 + (col << 4) + ((m & 3) << 2);
    } else {
      // K-major SW128: row m of 128 B (32 floats of k), 8-row atoms of 1024 B
      int col = (k & 31) >> 2;
      if (c.a_swz == 3) col ^= (m & 7);
      off = (k >> 5) * 16384 + m * 128 + (col << 4) + ((k & 3) << 2);
    }
    *(float*)(sa + off) = A[m * K + k];
  }
  // ---- B (N x K) ----------------------------------------------------------------------------
  for (int i = tid; i < c.N * K; i += 128) {
    const int n = i % c.N, k = i / c.N;
    uint32_t off;
    if (c.b_interleave) {
      // per 8 codes and 8 k: 256 B = [k-half 0: 8 rows x 16 B][k-half 1: 8 rows x 16 B]; k-steps 8 KB apart
      off = (k >> 3) * 8192 + (n >> 3) * 256 + ((k & 7) >> 2) * 128 + (n & 7) * 16 + ((k & 3) << 2);
    } else {
      int col = (k & 31) >> 2;
      if (c.b_swz == 3) col ^= (n & 7);
      off = (k >> 5) * 32768 + n * 128 + (col << 4) + ((k & 3) << 2);
    }
    *(float*)(sb + off) = B[n * K + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *slot;
  if (tid == 0) {
    uint32_t idesc = 0;
    idesc |= 1u << 4; idesc |= 2u << 7; idesc |= 2u << 10;
    idesc |= (uint32_t)(c.a_mn ? 1 : 0) << 15;
    idesc |= (uint32_t)(c.N >> 3) << 17;
    idesc |= (uint32_t)(128 >> 4) << 24;
    for (int ks = 0; ks < c.ksteps; ++ks) {
      const uint64_t ad = make_desc(smem_u32(sa) + (ks / 4) * 16384 + (ks % 4) * c.a_kstep, c.a_lbo, c.a_sbo, c.a_layout);
      const uint64_t bd = c.b_interleave
                              ? make_desc(smem_u32(sb) + ks * 8192, c.b_lbo, c.b_sbo, c.b_layout)
                              : make_desc(smem_u32(sb) + (ks / 4) * 32768 + (ks % 4) * c.b_kstep, c.b_lbo, c.b_sbo, c.b_layout);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tbase), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(smem_u32(bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tbase + ((uint32_t)(warp * 32) << 16) + c0)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) Dout[tid * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase) : "memory");
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  Cfg c{};
  c.N = 64; c.ksteps = 4;
  // B default: K-major SW128 (rows of 128 B), as DeepGEMM-style kernels use
  c.b_layout = 2; c.b_lbo = 16; c.b_sbo = 1024; c.b_kstep = 32; c.b_swz = 3; c.b_interleave = 0;
  switch (variant) {
    case 0:  // A K-major SW128 (baseline known-good form)
      c.a_mn = 0; c.a_layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_kstep = 32; c.a_swz = 3; break;
    case 1:  // A MN-major SW128: LBO = group stride 4096, SBO = k-atom stride 1024  (what vq_assign_tc uses)
      c.a_mn = 1; c.a_layout = 2; c.a_lbo = 4096; c.a_sbo = 1024; c.a_kstep = 1024; c.a_swz = 3; break;
    case 2:  // A MN-major SW128 with LBO/SBO swapped
      c.a_mn = 1; c.a_layout = 2; c.a_lbo = 1024; c.a_sbo = 4096; c.a_kstep = 1024; c.a_swz = 3; break;
    case 3:  // variant 1 + B in the no-swizzle interleaved layout (LBO = 128 k-half stride, SBO = 256 group stride)
      c.a_mn = 1; c.a_layout = 2; c.a_lbo = 4096; c.a_sbo = 1024; c.a_kstep = 1024; c.a_swz = 3;
      c.b_interleave = 1; c.b_layout = 0; c.b_lbo = 128; c.b_sbo = 256; break;
    case 4:  // variant 0 + B interleaved
      c.a_mn = 0; c.a_layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_kstep = 32; c.a_swz = 3;
      c.b_interleave = 1; c.b_layout = 0; c.b_lbo = 128; c.b_sbo = 256; break;
    case 5:  // variant 4 with LBO/SBO swapped for B
      c.a_mn = 0; c.a_layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_kstep = 32; c.a_swz = 3;
      c.b_interleave = 1; c.b_layout = 0; c.b_lbo = 256; c.b_sbo = 128; break;
    case 6:  // variant 1 with N = 256
      c.a_mn = 1; c.a_layout = 2; c.a_lbo = 4096; c.a_sbo = 1024; c.a_kstep = 1024; c.a_swz = 3; c.N = 256; break;
    case 7:  // variant 0 with N = 256
      c.a_mn = 0; c.a_layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_kstep = 32; c.a_swz = 3; c.N = 256; break;
    case 8:  // A MN-major, SWIZZLE_128B_BASE32B (layout 1), LBO = group stride, SBO = 4-row k-atom stride
      c.a_mn = 1; c.a_layout = 1; c.a_lbo = 4096; c.a_sbo = 512; c.a_kstep = 1024; c.a_swz = 2; break;
    case 9:  // same, LBO/SBO swapped
      c.a_mn = 1; c.a_layout = 1; c.a_lbo = 512; c.a_sbo = 4096; c.a_kstep = 1024; c.a_swz = 2; break;
    case 10: // variant 8, N = 256, B interleaved no-swizzle
      c.a_mn = 1; c.a_layout = 1; c.a_lbo = 4096; c.a_sbo = 512; c.a_kstep = 1024; c.a_swz = 2; c.N = 256;
      c.b_interleave = 1; c.b_layout = 0; c.b_lbo = 128; c.b_sbo = 256; break;
    case 11: // variant 8 with N = 256
      c.a_mn = 1; c.a_layout = 1; c.a_lbo = 4096; c.a_sbo = 512; c.a_kstep = 1024; c.a_swz = 2; c.N = 256; break;
    default: printf("unknown variant\n"); return 2;
  }
  const int K = 8 * c.ksteps;
  std::vector<float> A(128 * K), B(c.N * K), D(128 * c.N, -777.f);
  srand(1);
  for (auto& v : A) v = (float)((rand() % 17) - 8) * 0.25f;      // exactly representable in tf32
  for (auto& v : B) v = (float)((rand() % 13) - 6) * 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
  probe<<<1, 128, 140000>>>(dA, dB, dD, c);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d: sync -> %s\n", variant, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int nz = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < c.N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
      maxerr = fmax(maxerr, fabs(ref - D[m * c.N + n]));
      nz += D[m * c.N + n] != 0.f;
    }
  printf("variant %d: N=%d K=%d max|err| = %g, nonzero outputs = %d / %d ; D[0][0..3] = %g %g %g %g\n", variant, c.N, K,
         maxerr, nz, 128 * c.N, D[0], D[1], D[2], D[3]);
  return 0;
}

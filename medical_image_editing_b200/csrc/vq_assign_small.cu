// vq_assign_small.cu -- exact fp32 quantiser forward for SMALL codebooks in inference (no EMA statistics): K <= 16,
// K*D <= 256 -- the reference's real dictionaries (run_recon.py:27-48: K = 10, D = 16 on 512 x 512 slices).
//
// With so few codes the search is a handful of fmas per pixel, but the tensor-core kernel pays a fixed per-tile cost
// (barriers, publish, merge) for every 128 pixels.  Here: thread = pixel (a warp reads 128 contiguous bytes per
// channel), the codebook sits transposed in shared memory (every lane reads the same words: broadcast), all scores in
// registers, exactly the arithmetic of the other search kernels (dot = A + B over even / odd channel quads, ref_score,
// ties to the lowest index).  Measured at K = 10, D = 16: 0.064 ms per million pixels against 0.086 ms (tensor-core
// kernel); 16 x 512 x 512: 0.22 ms against 0.34 ms.  Training calls keep the tensor-core kernel: with thread = pixel
// the EMA sums would be 32 scattered reductions per instruction onto ten hot rows (measured 2x slower).
#include "vq_common.cuh"

namespace vqb200 {

constexpr int SM_THREADS = 256;
constexpr int SM_MAX_KD = 256;         // K * D beyond this the tensor-core kernel is faster

bool small_path_supported(int B, int D, int H, int W, int K) {
  return B > 0 && H > 0 && W > 0 && K >= 1 && K <= 16 && D >= 1 && (long long)K * D <= SM_MAX_KD;
}

template <int KMAX>
__global__ void __launch_bounds__(SM_THREADS)
vq_assign_small_kernel(const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ e2,
                       int B, int D, int H, int W, int K, int64_t* __restrict__ ids, int32_t* __restrict__ ids_nat,
                       float* __restrict__ q, double* __restrict__ loss_acc, int ids_mode) {
  extern __shared__ __align__(16) float smem_f[];
  float* e_s = smem_f;                              // [D][KMAX]  transposed codebook, 0 for k >= K
  float* e2_s = e_s + (size_t)D * KMAX;             // [KMAX]     |e|^2, +inf for k >= K
  const int tid = threadIdx.x, lane = tid & 31;
  const int HW = H * W;
  const long long N = (long long)B * HW;

  for (int i = tid; i < D * KMAX; i += SM_THREADS) {
    const int d = i / KMAX, k = i % KMAX;
    e_s[i] = k < K ? __ldg(E + (size_t)k * D + d) : 0.f;
  }
  for (int k = tid; k < KMAX; k += SM_THREADS) {
    e2_s[k] = k < K ? e2[k] : INFINITY;
  }
  __syncthreads();

  float lsum = 0.f;
  const int nq = (D + 3) >> 2;
  for (long long n0 = (long long)blockIdx.x * SM_THREADS; n0 < N; n0 += (long long)gridDim.x * SM_THREADS) {
    const long long n = n0 + tid;
    const bool valid = n < N;
    const long long b = valid ? n / HW : 0;
    const int p = valid ? (int)(n - b * HW) : 0;
    const float* zp = z + b * (long long)D * HW + p;

    // ---- search: exact scores of all codes (dot = A + B: fma chains over the even / odd channel quads) ----------
    float accA[KMAX], accB[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { accA[k] = 0.f; accB[k] = 0.f; }
    float z2A = 0.f, z2B = 0.f;                  // |z|^2 = A + B like the dot products (vq_common.cuh)
    if (valid) {
      for (int j = 0; j < nq; j += 2) {
        float zv8[8];                               // all eight loads of this pair of quads in flight before any use
#pragma unroll
        for (int u = 0; u < 8; ++u) zv8[u] = (4 * j + u < D) ? __ldg(zp + (long long)(4 * j + u) * HW) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {               // u < 4: even quad j (chain A); u >= 4: odd quad j + 1 (chain B)
          const int d = 4 * j + u;
          if (d < D) {
            const float zv = zv8[u];
            if (u < 4) z2A = __fmaf_rn(zv, zv, z2A);
            else z2B = __fmaf_rn(zv, zv, z2B);
            const float4* er = reinterpret_cast<const float4*>(e_s + (size_t)d * KMAX);
#pragma unroll
            for (int k4 = 0; k4 < KMAX / 4; ++k4) {
              const float4 e4 = er[k4];
              if (u < 4) {
                accA[4 * k4 + 0] = __fmaf_rn(zv, e4.x, accA[4 * k4 + 0]);
                accA[4 * k4 + 1] = __fmaf_rn(zv, e4.y, accA[4 * k4 + 1]);
                accA[4 * k4 + 2] = __fmaf_rn(zv, e4.z, accA[4 * k4 + 2]);
                accA[4 * k4 + 3] = __fmaf_rn(zv, e4.w, accA[4 * k4 + 3]);
              } else {
                accB[4 * k4 + 0] = __fmaf_rn(zv, e4.x, accB[4 * k4 + 0]);
                accB[4 * k4 + 1] = __fmaf_rn(zv, e4.y, accB[4 * k4 + 1]);
                accB[4 * k4 + 2] = __fmaf_rn(zv, e4.z, accB[4 * k4 + 2]);
                accB[4 * k4 + 3] = __fmaf_rn(zv, e4.w, accB[4 * k4 + 3]);
              }
            }
          }
        }
      }
    }
    const float z2 = __fadd_rn(z2A, z2B);
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {                  // ascending k, strict '>' keeps the lowest index; padding scores -inf
      const float s = ref_score(__fadd_rn(accA[k], accB[k]), e2_s[k], z2);
      if (s > best) { best = s; bi = k; }
    }
    if (!(best > -INFINITY)) bi = 0;                  // all-NaN row: the reference's topk returns index 0

    if (valid) {
      const int h = p / W, w = p - h * W;
      if (ids) store_id(ids + b * HW, p, h, w, H, bi, ids_mode);
      if (ids_nat) ids_nat[n] = bi;
    }

    // ---- outputs: q, (z - q)^2 (z is re-read through L1) ------------------------------------------
    float* qp = q ? q + b * (long long)D * HW + p : nullptr;
    for (int d0 = 0; d0 < D; d0 += 8) {
      float zv8[8];                                   // eight channels per step: the loads first
#pragma unroll
      for (int u = 0; u < 8; ++u) zv8[u] = (valid && d0 + u < D) ? __ldg(zp + (long long)(d0 + u) * HW) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int d = d0 + u;
        if (d < D) {                                  // warp-uniform
          const float zv = zv8[u];
          const float ev = e_s[(size_t)d * KMAX + bi];
          if (valid) {
            const float df = zv - ev;
            lsum = __fmaf_rn(df, df, lsum);
            if (qp) __stcs(qp + (long long)d * HW, ev);
          }
        }
      }
    }
  }

  lsum = warp_sum(lsum);
  if (lane == 0 && loss_acc && lsum != 0.f) atomicAdd(loss_acc, (double)lsum);
}

static int small_sm_count() { return device_sm_count(); }

int launch_assign_small(const FwdArgs& a, cudaStream_t s) {
  VQ_REQUIRE(small_path_supported(a.B, a.D, a.H, a.W, a.K) && a.stats == nullptr, VQ_ERR_UNSUPPORTED,
             "small-codebook path: unsupported shape or training call");
  const long long N = (long long)a.B * a.H * a.W;
  constexpr int KMAX = 16;
  const size_t smem = ((size_t)a.D * KMAX + KMAX) * sizeof(float);
  long long blocks = (N + SM_THREADS - 1) / SM_THREADS;
  const long long cap = (long long)small_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const bool prof = profile_begin(s);
  vq_assign_small_kernel<KMAX><<<(unsigned)blocks, SM_THREADS, smem, s>>>(a.z, a.embed, a.ws.e2, a.B, a.D, a.H, a.W, a.K,
                                                                          a.ids, a.ids_nat, a.q, a.ws.loss_acc, ids_mode_of(a.flags));
  if (prof) profile_end(s);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

}  // namespace vqb200

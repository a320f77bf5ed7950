// vq_embed_loss.cu -- cross-view cluster loss of EmbeddingLoss (reference: src/functions/embed_loss.py:46-66), the direct
// consumer of the quantiser's outputs in the stage-1 trainers (single_window_trainer.py:91-105).  SURVEY section 8(f) rank 1.
//
// The reference expands embed and codebook to a (b, n_features, n_clusters, n_loc) tensor (B*D*K*HW floats: 2 TB at
// config 2 -- it only runs because the real K is ~10) and multiplies by the one-hot label map.  Here: one pass over z
// (HBM-bound, 4*D bytes per pixel), every pixel measured against the ONE code its label names, warp-aggregated atomics
// into a (B, K) table, a one-CTA finalise, and a streaming backward of the same shape as vq_bwd.
#include "vq_common.cuh"

namespace vqb200 {

constexpr float EL_EPS = 1e-6f;       // EmbeddingLoss.epsilon (embed_loss.py:8)

struct ElWork {
  double* sums;      // [R][B*K]  sum of squared distances of the pixels labelled (b, k), R replicas
  int* counts;       // [R][B*K]
  int nrep;
  size_t bytes;
};
// small (B, K) tables (run_recon: K = 10) are hit by every warp: spread the atomics over replicas (CTA -> replica)
static inline int el_replicas(int B, int K) {
  long long r = 16384LL / ((long long)B * K);
  return (int)(r < 1 ? 1 : (r > 64 ? 64 : r));
}
static inline ElWork carve_el(void* base, int B, int K) {
  ElWork w;
  w.nrep = el_replicas(B, K);
  const size_t n = (size_t)B * K * w.nrep;
  w.sums = (double*)base;
  w.counts = (int*)((char*)base + align_up(n * sizeof(double), 256));
  w.bytes = align_up(n * sizeof(double), 256) + align_up(n * sizeof(int), 256);
  return w;
}
size_t embed_loss_work_bytes(int B, int K) { return carve_el(nullptr, B, K).bytes; }

// add (v, n) to table entry `key` for every lane with key >= 0; lanes with equal keys are combined first so that a
// piecewise-constant label map (the usual case: neighbouring pixels share a class) costs one atomic per warp and class
__device__ __forceinline__ void warp_group_add(int key, float v, int n, double* sums, int* counts) {
  const unsigned full = 0xffffffffu;
  const unsigned grp = __match_any_sync(full, key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(grp) - 1;
  float tot = 0.f;
  int cnt = 0;
  for (unsigned rest = grp; rest; rest &= rest - 1) {                       // fixed (lane) order
    tot += __shfl_sync(grp, v, __ffs(rest) - 1);
    cnt += __shfl_sync(grp, n, __ffs(rest) - 1);
  }
  if (lane == leader && key >= 0) {
    atomicAdd(sums + key, (double)tot);
    if (counts) atomicAdd(counts + key, cnt);
  }
}

// thread = 4 consecutive pixels (HW % 4 == 0): float4 streaming loads of z along the pixels, code rows through L1
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
vq_el_accum_vec_kernel(const float* __restrict__ z, const int32_t* __restrict__ labels, const float* __restrict__ E,
                       int D, int HW, int K, long long nquads, double* __restrict__ sums, int* __restrict__ counts,
                       int nrep, int table, int dsplit, int same_row) {
  // blockIdx.y = slice of the channels (finer CTAs: the grid is only ~2 waves of whole-pixel CTAs otherwise); the pixel
  // counts come from slice 0 alone
  sums += (size_t)(blockIdx.x % nrep) * table;
  counts += (size_t)(blockIdx.x % nrep) * table;
  const long long quad = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int dper = D / dsplit;                   // dsplit > 1 only when D % (4 * dsplit) == 0
  const int dbeg = blockIdx.y * dper, dend = dbeg + dper;
  int key[4] = {-1, -1, -1, -1};
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (quad < nquads) {
    const int qpi = HW >> 2;
    const long long b = quad / qpi;
    const int p = (int)(quad % qpi) << 2;
    const int4 l4 = *reinterpret_cast<const int4*>(labels + b * HW + p);
    const int lab[4] = {l4.x, l4.y, l4.z, l4.w};
    const float* er[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = lab[i] >= 1 && lab[i] <= K;
      key[i] = ok ? (int)(b * K + lab[i] - 1) : -1;
      er[i] = E + (size_t)(ok ? lab[i] - 1 : 0) * D;
    }
    const long long base = b * (long long)D * HW + p;
    int d = dbeg;
    const bool vec_rows = (D & 3) == 0 && (dper & 3) == 0 && ((uintptr_t)E & 15) == 0;
    if (vec_rows && same_row && lab[0] == lab[1] && lab[1] == lab[2] && lab[2] == lab[3]) {
      // the four pixels share a class (the usual case on a segmentation map): ONE code-row load per step instead of
      // four -- the scattered 16-byte row loads, not the z stream, are what fills the L1 wavefront budget here (a warp's
      // row load touches up to 16 lines, its z load 4).  Same operations per pixel in the same order: bit-identical.
      for (; d < dend; d += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(er[0] + d));
        const float ev[4] = {a.x, a.y, a.z, a.w};
        float4 zv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) zv[c] = __ldcs(reinterpret_cast<const float4*>(z + base + (long long)(d + c) * HW));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float d0 = zv[c].x - ev[c], d1 = zv[c].y - ev[c], d2 = zv[c].z - ev[c], d3 = zv[c].w - ev[c];
          acc[0] = __fmaf_rn(d0, d0, acc[0]);
          acc[1] = __fmaf_rn(d1, d1, acc[1]);
          acc[2] = __fmaf_rn(d2, d2, acc[2]);
          acc[3] = __fmaf_rn(d3, d3, acc[3]);
        }
      }
    } else if (vec_rows) {
      for (; d < dend; d += 4) {                    // four channels per step: float4 loads of the code rows too
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(er[0] + d));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(er[1] + d));
        const float4 a2 = __ldg(reinterpret_cast<const float4*>(er[2] + d));
        const float4 a3 = __ldg(reinterpret_cast<const float4*>(er[3] + d));
        const float ev[4][4] = {{a0.x, a1.x, a2.x, a3.x}, {a0.y, a1.y, a2.y, a3.y},
                                {a0.z, a1.z, a2.z, a3.z}, {a0.w, a1.w, a2.w, a3.w}};
        float4 zv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) zv[c] = __ldcs(reinterpret_cast<const float4*>(z + base + (long long)(d + c) * HW));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float d0 = zv[c].x - ev[c][0], d1 = zv[c].y - ev[c][1], d2 = zv[c].z - ev[c][2], d3 = zv[c].w - ev[c][3];
          acc[0] = __fmaf_rn(d0, d0, acc[0]);
          acc[1] = __fmaf_rn(d1, d1, acc[1]);
          acc[2] = __fmaf_rn(d2, d2, acc[2]);
          acc[3] = __fmaf_rn(d3, d3, acc[3]);
        }
      }
    }
    for (; d < dend; ++d) {
      const float4 zv = __ldcs(reinterpret_cast<const float4*>(z + base + (long long)d * HW));
      const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float df = zz[i] - __ldg(er[i] + d);
        acc[i] = __fmaf_rn(df, df, acc[i]);
      }
    }
  }
  // a thread's four pixels usually share a class: fold equal keys inside the thread first (a quarter of the atomics on
  // piecewise-constant maps), then across the warp
  int cnt[4] = {1, 1, 1, 1};
#pragma unroll
  for (int i = 1; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) {
      if (key[i] >= 0 && key[i] == key[j]) {
        acc[j] += acc[i];
        cnt[j] += cnt[i];
        key[i] = -1;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (__any_sync(0xffffffffu, key[i] >= 0))
      warp_group_add(key[i], acc[i], cnt[i], sums, blockIdx.y == 0 ? counts : nullptr);
  }
}

// generic path: thread = one pixel
__global__ void __launch_bounds__(256)
vq_el_accum_generic_kernel(const float* __restrict__ z, const int32_t* __restrict__ labels, const float* __restrict__ E,
                           int D, int HW, int K, long long N, double* __restrict__ sums, int* __restrict__ counts,
                           int nrep, int table) {
  sums += (size_t)(blockIdx.x % nrep) * table;
  counts += (size_t)(blockIdx.x % nrep) * table;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int key = -1;
  float acc = 0.f;
  if (n < N) {
    const long long b = n / HW;
    const int p = (int)(n % HW);
    const int lab = labels[n];
    if (lab >= 1 && lab <= K) {
      key = (int)(b * K + lab - 1);
      const float* er = E + (size_t)(lab - 1) * D;
      const float* zp = z + b * (long long)D * HW + p;
      for (int d = 0; d < D; ++d) {
        const float df = __ldg(zp + (long long)d * HW) - __ldg(er + d);
        acc = __fmaf_rn(df, df, acc);
      }
    }
  }
  if (__any_sync(0xffffffffu, key >= 0)) warp_group_add(key, acc, 1, sums, counts);
}

// one CTA: loss = mean over present (b, k) of sums / (count + eps); weights for the backward pass
__global__ void __launch_bounds__(1024)
vq_el_finish_kernel(const double* __restrict__ sums, const int* __restrict__ counts, int n, int nrep,
                    float* __restrict__ loss, float* __restrict__ weights) {
  __shared__ double red_v[32];
  __shared__ int red_c[32];
  __shared__ int s_present;
  double part = 0.0;
  int present = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int c = 0;
    double sm = 0.0;
    for (int r = 0; r < nrep; ++r) { c += counts[(size_t)r * n + i]; sm += sums[(size_t)r * n + i]; }   // fixed order
    if (c > 0) {
      // fp32 like the reference: cross_dist.sum(2) / (r_ids.sum(2) + epsilon)
      part += (double)((float)sm / ((float)c + EL_EPS));
      ++present;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    part += __shfl_xor_sync(0xffffffffu, part, o);
    present += __shfl_xor_sync(0xffffffffu, present, o);
  }
  if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = part; red_c[threadIdx.x >> 5] = present; }
  __syncthreads();
  if (threadIdx.x < 32) {
    part = threadIdx.x < (blockDim.x >> 5) ? red_v[threadIdx.x] : 0.0;
    present = threadIdx.x < (blockDim.x >> 5) ? red_c[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) {
      part += __shfl_xor_sync(0xffffffffu, part, o);
      present += __shfl_xor_sync(0xffffffffu, present, o);
    }
    if (threadIdx.x == 0) {
      s_present = present;
      *loss = (float)(part / (double)present);          // no class present anywhere: 0/0 = NaN, as torch's mean of nothing
    }
  }
  __syncthreads();
  const float inv_present = 1.0f / (float)s_present;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int c = 0;
    for (int r = 0; r < nrep; ++r) c += counts[(size_t)r * n + i];
    weights[i] = c > 0 ? inv_present / ((float)c + EL_EPS) : 0.f;
  }
}

// backward: g_z = g_loss * 2 * w[b, k] * (z - c_k) for labelled pixels, 0 elsewhere
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
vq_el_bwd_vec_kernel(const float* __restrict__ g_loss, const float* __restrict__ z, const int32_t* __restrict__ labels,
                     const float* __restrict__ E, const float* __restrict__ weights, float* __restrict__ g_z,
                     int D, int HW, int K, long long nquads, int dsplit, int same_row) {
  const long long quad = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (quad >= nquads) return;
  const int qpi = HW >> 2;
  const long long b = quad / qpi;
  const int p = (int)(quad % qpi) << 2;
  const float g2 = 2.0f * (*g_loss);
  const int4 l4 = *reinterpret_cast<const int4*>(labels + b * HW + p);
  const int lab[4] = {l4.x, l4.y, l4.z, l4.w};
  const float* er[4];
  float coef[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool ok = lab[i] >= 1 && lab[i] <= K;
    er[i] = E + (size_t)(ok ? lab[i] - 1 : 0) * D;
    coef[i] = ok ? g2 * __ldg(weights + b * K + lab[i] - 1) : 0.f;
  }
  const int dper = D / dsplit;
  const int dbeg = blockIdx.y * dper, dend = dbeg + dper;
  const long long base = b * (long long)D * HW + p;
  int d = dbeg;
  const bool vec_rows = (dper & 3) == 0 && ((uintptr_t)E & 15) == 0 && (D & 3) == 0;
  if (vec_rows && same_row && lab[0] == lab[1] && lab[1] == lab[2] && lab[2] == lab[3]) {
    for (; d < dend; d += 4) {                      // one class for the four pixels: one code-row load per step (see the accumulation)
      const float4 a = __ldg(reinterpret_cast<const float4*>(er[0] + d));
      const float ev[4] = {a.x, a.y, a.z, a.w};
      float4 zv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) zv[c] = __ldcs(reinterpret_cast<const float4*>(z + base + (long long)(d + c) * HW));
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 r;
        r.x = coef[0] * (zv[c].x - ev[c]);
        r.y = coef[1] * (zv[c].y - ev[c]);
        r.z = coef[2] * (zv[c].z - ev[c]);
        r.w = coef[3] * (zv[c].w - ev[c]);
        __stcs(reinterpret_cast<float4*>(g_z + base + (long long)(d + c) * HW), r);
      }
    }
  } else if (vec_rows) {
    for (; d < dend; d += 4) {                      // four channels per step: float4 loads of the code rows (as vq_bwd_vec)
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(er[0] + d));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(er[1] + d));
      const float4 a2 = __ldg(reinterpret_cast<const float4*>(er[2] + d));
      const float4 a3 = __ldg(reinterpret_cast<const float4*>(er[3] + d));
      const float ev[4][4] = {{a0.x, a1.x, a2.x, a3.x}, {a0.y, a1.y, a2.y, a3.y},
                              {a0.z, a1.z, a2.z, a3.z}, {a0.w, a1.w, a2.w, a3.w}};
      float4 zv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) zv[c] = __ldcs(reinterpret_cast<const float4*>(z + base + (long long)(d + c) * HW));
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 r;
        r.x = coef[0] * (zv[c].x - ev[c][0]);
        r.y = coef[1] * (zv[c].y - ev[c][1]);
        r.z = coef[2] * (zv[c].z - ev[c][2]);
        r.w = coef[3] * (zv[c].w - ev[c][3]);
        __stcs(reinterpret_cast<float4*>(g_z + base + (long long)(d + c) * HW), r);
      }
    }
  }
  for (; d < dend; ++d) {
    const float4 zv = __ldcs(reinterpret_cast<const float4*>(z + base + (long long)d * HW));
    float4 r;
    r.x = coef[0] * (zv.x - __ldg(er[0] + d));
    r.y = coef[1] * (zv.y - __ldg(er[1] + d));
    r.z = coef[2] * (zv.z - __ldg(er[2] + d));
    r.w = coef[3] * (zv.w - __ldg(er[3] + d));
    __stcs(reinterpret_cast<float4*>(g_z + base + (long long)d * HW), r);
  }
}

__global__ void __launch_bounds__(256)
vq_el_bwd_generic_kernel(const float* __restrict__ g_loss, const float* __restrict__ z, const int32_t* __restrict__ labels,
                         const float* __restrict__ E, const float* __restrict__ weights, float* __restrict__ g_z,
                         int D, int HW, int K, long long N) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long b = n / HW;
  const int p = (int)(n % HW);
  const int lab = labels[n];
  const bool ok = lab >= 1 && lab <= K;
  const float coef = ok ? 2.0f * (*g_loss) * weights[b * K + lab - 1] : 0.f;
  const float* er = E + (size_t)(ok ? lab - 1 : 0) * D;
  const long long base = b * (long long)D * HW + p;
  for (int d = 0; d < D; ++d) g_z[base + (long long)d * HW] = coef * (z[base + (long long)d * HW] - __ldg(er + d));
}

static int el_sm_count() { return device_sm_count(); }

#ifndef VQ_EL_SAMEROW_DEFAULT
#define VQ_EL_SAMEROW_DEFAULT 1
#endif
#ifndef VQ_EL_BWD_MINB_DEFAULT
#define VQ_EL_BWD_MINB_DEFAULT 4
#endif
#ifndef VQ_EL_ACC_MINB_DEFAULT
#define VQ_EL_ACC_MINB_DEFAULT 4
#endif
#ifndef VQ_EL_ACC_DSPLIT_DEFAULT
#define VQ_EL_ACC_DSPLIT_DEFAULT 1
#endif
#ifndef VQ_EL_BWD_DSPLIT_DEFAULT
#define VQ_EL_BWD_DSPLIT_DEFAULT 4
#endif

int launch_embed_loss_fwd(const float* z, const int32_t* labels, const float* embed, int B, int D, int H, int W, int K,
                          float* loss, float* weights, void* work, cudaStream_t s) {
  const int HW = H * W;
  const long long N = (long long)B * HW;
  const ElWork w = carve_el(work, B, K);
  VQ_CUDA_CHECK(cudaMemsetAsync(work, 0, w.bytes, s));
  if (N > 0) {
    const bool vec = (HW % 4 == 0) && ((((uintptr_t)z) | ((uintptr_t)labels)) & 15) == 0;
    if (vec) {
      const long long nquads = N / 4;
      int dsplit = 1;
      const int want = tuning_knob("VQ_EL_ACC_DSPLIT", VQ_EL_ACC_DSPLIT_DEFAULT);
      while (dsplit < want && D % (dsplit * 8) == 0) dsplit *= 2;
      dim3 grid((unsigned)((nquads + 255) / 256), (unsigned)dsplit);
      const int same_row = tuning_knob("VQ_EL_SAMEROW", VQ_EL_SAMEROW_DEFAULT);
      const int minb = tuning_knob("VQ_EL_ACC_MINB", VQ_EL_ACC_MINB_DEFAULT);      // register cap -> resident CTAs per SM
      if (minb >= 6)
        vq_el_accum_vec_kernel<6><<<grid, 256, 0, s>>>(z, labels, embed, D, HW, K, nquads, w.sums, w.counts, w.nrep, B * K, dsplit, same_row);
      else if (minb >= 5)
        vq_el_accum_vec_kernel<5><<<grid, 256, 0, s>>>(z, labels, embed, D, HW, K, nquads, w.sums, w.counts, w.nrep, B * K, dsplit, same_row);
      else if (minb >= 4)
        vq_el_accum_vec_kernel<4><<<grid, 256, 0, s>>>(z, labels, embed, D, HW, K, nquads, w.sums, w.counts, w.nrep, B * K, dsplit, same_row);
      else
        vq_el_accum_vec_kernel<0><<<grid, 256, 0, s>>>(z, labels, embed, D, HW, K, nquads, w.sums, w.counts, w.nrep, B * K, dsplit, same_row);
    } else {
      vq_el_accum_generic_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(z, labels, embed, D, HW, K, N, w.sums, w.counts, w.nrep, B * K);
    }
    count_launch();
    VQ_CUDA_CHECK(cudaGetLastError());
  }
  vq_el_finish_kernel<<<1, 1024, 0, s>>>(w.sums, w.counts, B * K, w.nrep, loss, weights);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

int launch_embed_loss_bwd(const float* g_loss, const float* z, const int32_t* labels, const float* embed,
                          const float* weights, float* g_z, int B, int D, int H, int W, int K, cudaStream_t s) {
  const int HW = H * W;
  const long long N = (long long)B * HW;
  if (N == 0) return VQ_OK;
  const bool vec = (HW % 4 == 0) && ((((uintptr_t)z) | ((uintptr_t)labels) | ((uintptr_t)g_z)) & 15) == 0;
  if (vec) {
    const long long nquads = N / 4;
    const long long bx = (nquads + 255) / 256;
    int dsplit = 1;
    while (bx * dsplit < 4LL * el_sm_count() && D % (dsplit * 2) == 0) dsplit *= 2;
    const int want = tuning_knob("VQ_EL_BWD_DSPLIT", VQ_EL_BWD_DSPLIT_DEFAULT);
    while (dsplit < want && D % (dsplit * 8) == 0 && D / (dsplit * 2) >= 16) dsplit *= 2;
    dim3 grid((unsigned)bx, (unsigned)dsplit);
    const int same_row = tuning_knob("VQ_EL_SAMEROW", VQ_EL_SAMEROW_DEFAULT);
    if (tuning_knob("VQ_EL_BWD_MINB", VQ_EL_BWD_MINB_DEFAULT) >= 4)
      vq_el_bwd_vec_kernel<4><<<grid, 256, 0, s>>>(g_loss, z, labels, embed, weights, g_z, D, HW, K, nquads, dsplit, same_row);
    else
      vq_el_bwd_vec_kernel<0><<<grid, 256, 0, s>>>(g_loss, z, labels, embed, weights, g_z, D, HW, K, nquads, dsplit, same_row);
  } else {
    vq_el_bwd_generic_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(g_loss, z, labels, embed, weights, g_z, D, HW, K, N);
  }
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

}  // namespace vqb200

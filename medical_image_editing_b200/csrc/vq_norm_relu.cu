// vq_norm_relu.cu -- the U-Net tail that produces the quantiser's input (sm_100a only).
//
// The last two layers in front of `VQModule` are `nn.InstanceNorm2d(C)` + `nn.ReLU(inplace=True)` (the end of
// `up_conv1_1.double_conv`, reference blocks.py:39-50, vqwnet.py:104-108).  Stock torch runs them as batch-norm statistics +
// normalise (x read twice, y written) and an in-place ReLU (y read and written): five N*D-sized passes in the forward and
// seven in the backward.  z itself must exist in HBM -- the networks return it as `embed` (vqwnet.py:108) and the stage-1
// trainers feed it to EmbeddingLoss -- so the fusion that is left is on the producer side: z is written ONCE, straight
// from the convolution output.
//
// Layout: NCHW fp32; a (b, c) plane is HW contiguous floats.  A plane is handled by a thread-block CLUSTER of S CTAs
// (S = 1, 2, 4, 8; one segment of the plane each): pass 1 reads the segment and reduces shifted sums, the partial sums
// are exchanged through distributed shared memory, pass 2 re-reads the segment -- which is still in L2: the launcher
// sizes segments and the persistent grid so that all segments in flight fit in the 126 MB L2 -- and writes the result.
// DRAM traffic is therefore the minimum: forward read x + write z, backward read x, g_z + write g_x.
#include <cooperative_groups.h>

#include "vq_common.cuh"

namespace cg = cooperative_groups;

namespace vqb200 {
namespace {

constexpr int NR_THREADS = 512;
constexpr int NR_WARPS = NR_THREADS / 32;
constexpr int NR_NSIZES = 5;                // cluster sizes 1, 2, 4, 8, 16 (8 is the portable limit; 16 needs the non-portable attribute)

struct NrCaps { int active[NR_NSIZES]; };   // co-resident clusters of size 1 << i on this device (0: size not available)

struct NrPlan {
  int S;            // cluster size = segments per plane
  long long seg;    // floats per segment (multiple of 4 when the vector path is used)
  int clusters;     // persistent clusters in the grid
};

// Segments of at most `seg_max` floats, so that (CTAs in flight) x (bytes a CTA streams per plane) stays below L2; among
// the cluster sizes that satisfy it, the one whose persistent schedule has the shortest tail (rounds x segment length).
// The persistent grid never holds more clusters than can be co-resident (a waiting cluster would be a second wave).
NrPlan nr_plan(long long planes, long long HW, bool vec, long long seg_max, const NrCaps& caps) {
  int i_min = 0;
  while (i_min + 1 < NR_NSIZES && caps.active[i_min + 1] > 0 && (HW + (1 << i_min) - 1) / (1 << i_min) > seg_max) ++i_min;
  int best = i_min;
  double best_cost = 1e300;
  for (int i = i_min; i < NR_NSIZES && caps.active[i] > 0; ++i) {
    const int S = 1 << i;
    const long long clusters = caps.active[i];
    const long long rounds = (planes + clusters - 1) / clusters;
    const double cost = (double)rounds / (double)S * (1.0 + 0.01 * S);      // mild preference for smaller clusters
    if (cost < best_cost) { best_cost = cost; best = i; }
  }
  NrPlan p;
  p.S = 1 << best;
  long long seg = (HW + p.S - 1) / p.S;
  if (vec) seg = (seg + 3) / 4 * 4;
  p.seg = seg;
  long long clusters = caps.active[best];
  if (clusters > planes) clusters = planes;
  if (clusters < 1) clusters = 1;
  p.clusters = (int)clusters;
  return p;
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ float nr_gx(float xv, float gv, float mean, float rstd, float m1, float m2) {
  const float xh = (xv - mean) * rstd;
  const float gy = xh > 0.f ? gv : 0.f;
  return rstd * ((gy - m1) - xh * m2);
}

// sum over the cluster of two doubles per CTA; every CTA gets the totals.  `slot` is this CTA's shared-memory pair.
__device__ __forceinline__ void cluster_sum2(double (&v)[2], double* slot, double* warp_part, int nranks) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    v[1] += __shfl_xor_sync(0xffffffffu, v[1], o);
  }
  if (lane == 0) { warp_part[2 * warp] = v[0]; warp_part[2 * warp + 1] = v[1]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < NR_WARPS; ++w) { a += warp_part[2 * w]; b += warp_part[2 * w + 1]; }
    slot[0] = a; slot[1] = b;
  }
  if (nranks > 1) {
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();                                              // partial sums of every CTA are visible cluster-wide
    double a = 0.0, b = 0.0;
    for (int r = 0; r < nranks; ++r) {                      // same order in every CTA: identical totals
      const double* rs = cl.map_shared_rank(slot, r);
      a += rs[0]; b += rs[1];
    }
    v[0] = a; v[1] = b;
    cl.sync();                                              // nobody overwrites its slot (next plane) while a peer still reads it
  } else {
    __syncthreads();
    v[0] = slot[0]; v[1] = slot[1];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward: z = relu((x - mean) * rstd), mean / biased variance over the plane, rstd = 1 / sqrt(var + eps)
// (torch.nn.functional.instance_norm without affine parameters or running statistics, then relu)
// ------------------------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(NR_THREADS, 2)
vq_norm_relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ z, float2* __restrict__ stats, long long planes,
                        long long HW, long long seg, int S, float eps) {
  __shared__ double slot[2];
  __shared__ double warp_part[2 * NR_WARPS];
  const int rank = S > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const long long cluster_id = blockIdx.x / S, nclusters = gridDim.x / S;
  const long long s0 = (long long)rank * seg;
  const long long s1 = s0 + seg < HW ? s0 + seg : HW;        // my segment [s0, s1) of every plane (may be empty)
  for (long long plane = cluster_id; plane < planes; plane += nclusters) {
    const float* xp = x + plane * HW;
    float* zp = z + plane * HW;
    const float shift = __ldg(xp);                           // any sample of the plane: keeps the sums well conditioned
    float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
    if (VEC) {
      const float4* x4 = reinterpret_cast<const float4*>(xp);
      const long long i0 = s0 >> 2, i1 = s1 >> 2;
      long long i = i0 + threadIdx.x;
      for (; i + 3 * NR_THREADS < i1; i += 4 * NR_THREADS) {   // four independent 16-byte loads per thread in flight
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(x4 + i + u * NR_THREADS);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float dx = v[u].x - shift, dy = v[u].y - shift, dz = v[u].z - shift, dw = v[u].w - shift;
          a0 += dx + dy; a1 += dz + dw;
          q0 = __fmaf_rn(dx, dx, q0); q0 = __fmaf_rn(dy, dy, q0);
          q1 = __fmaf_rn(dz, dz, q1); q1 = __fmaf_rn(dw, dw, q1);
        }
      }
      for (; i < i1; i += NR_THREADS) {
        const float4 v = __ldg(x4 + i);
        const float dx = v.x - shift, dy = v.y - shift, dz = v.z - shift, dw = v.w - shift;
        a0 += dx + dy; a1 += dz + dw;
        q0 = __fmaf_rn(dx, dx, q0); q0 = __fmaf_rn(dy, dy, q0);
        q1 = __fmaf_rn(dz, dz, q1); q1 = __fmaf_rn(dw, dw, q1);
      }
    } else {
      for (long long i = s0 + threadIdx.x; i < s1; i += NR_THREADS) {
        const float d = __ldg(xp + i) - shift;
        a0 += d;
        q0 = __fmaf_rn(d, d, q0);
      }
    }
    double v2[2] = {(double)a0 + (double)a1, (double)q0 + (double)q1};
    cluster_sum2(v2, slot, warp_part, S);
    const double n = (double)HW;
    const double md = v2[0] / n;                              // mean - shift
    double var = v2[1] / n - md * md;
    if (var < 0.0) var = 0.0;
    const float mean = (float)((double)shift + md);
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (rank == 0 && threadIdx.x == 0 && stats) stats[plane] = make_float2(mean, rstd);
    if (VEC) {
      const float4* x4 = reinterpret_cast<const float4*>(xp);
      float4* z4 = reinterpret_cast<float4*>(zp);
      const long long i0 = s0 >> 2, i1 = s1 >> 2;
      long long i = i0 + threadIdx.x;
      for (; i + 3 * NR_THREADS < i1; i += 4 * NR_THREADS) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg_stream(x4 + i + u * NR_THREADS);     // L2 hit; last use of the line
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float4 o;
          o.x = fmaxf((v[u].x - mean) * rstd, 0.f); o.y = fmaxf((v[u].y - mean) * rstd, 0.f);
          o.z = fmaxf((v[u].z - mean) * rstd, 0.f); o.w = fmaxf((v[u].w - mean) * rstd, 0.f);
          z4[i + u * NR_THREADS] = o;                         // plain store: the quantiser reads z next
        }
      }
      for (; i < i1; i += NR_THREADS) {
        const float4 v = ldg_stream(x4 + i);
        float4 o;
        o.x = fmaxf((v.x - mean) * rstd, 0.f); o.y = fmaxf((v.y - mean) * rstd, 0.f);
        o.z = fmaxf((v.z - mean) * rstd, 0.f); o.w = fmaxf((v.w - mean) * rstd, 0.f);
        z4[i] = o;
      }
    } else {
      for (long long i = s0 + threadIdx.x; i < s1; i += NR_THREADS) zp[i] = fmaxf((__ldg(xp + i) - mean) * rstd, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: g_y = g_z where xhat > 0 else 0 (relu), g_x = rstd * (g_y - mean(g_y) - xhat * mean(g_y * xhat))
// xhat is recomputed from x with the forward's own expression, so the relu mask is the forward's
// ------------------------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(NR_THREADS, 2)
vq_norm_relu_bwd_kernel(const float* __restrict__ g_z, const float* __restrict__ x, const float2* __restrict__ stats,
                        float* __restrict__ g_x, long long planes, long long HW, long long seg, int S) {
  __shared__ double slot[2];
  __shared__ double warp_part[2 * NR_WARPS];
  const int rank = S > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const long long cluster_id = blockIdx.x / S, nclusters = gridDim.x / S;
  const long long s0 = (long long)rank * seg;
  const long long s1 = s0 + seg < HW ? s0 + seg : HW;
  for (long long plane = cluster_id; plane < planes; plane += nclusters) {
    const float* xp = x + plane * HW;
    const float* gp = g_z + plane * HW;
    float* op = g_x + plane * HW;
    const float2 st = __ldg(stats + plane);
    const float mean = st.x, rstd = st.y;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#define NR_ACC(xv, gv, A, Bq)                                       \
    {                                                               \
      const float xh = ((xv) - mean) * rstd;                        \
      const float gy = xh > 0.f ? (gv) : 0.f;                       \
      A += gy;                                                      \
      Bq = __fmaf_rn(gy, xh, Bq);                                   \
    }
    if (VEC) {
      const float4* x4 = reinterpret_cast<const float4*>(xp);
      const float4* g4 = reinterpret_cast<const float4*>(gp);
      const long long i0 = s0 >> 2, i1 = s1 >> 2;
      long long i = i0 + threadIdx.x;
      for (; i + NR_THREADS < i1; i += 2 * NR_THREADS) {       // four independent 16-byte loads per thread in flight
        const float4 xa = __ldg(x4 + i), xb = __ldg(x4 + i + NR_THREADS);
        const float4 ga = __ldg(g4 + i), gb = __ldg(g4 + i + NR_THREADS);
        NR_ACC(xa.x, ga.x, a0, b0) NR_ACC(xa.y, ga.y, a1, b1) NR_ACC(xa.z, ga.z, a0, b0) NR_ACC(xa.w, ga.w, a1, b1)
        NR_ACC(xb.x, gb.x, a0, b0) NR_ACC(xb.y, gb.y, a1, b1) NR_ACC(xb.z, gb.z, a0, b0) NR_ACC(xb.w, gb.w, a1, b1)
      }
      for (; i < i1; i += NR_THREADS) {
        const float4 xa = __ldg(x4 + i), ga = __ldg(g4 + i);
        NR_ACC(xa.x, ga.x, a0, b0) NR_ACC(xa.y, ga.y, a1, b1) NR_ACC(xa.z, ga.z, a0, b0) NR_ACC(xa.w, ga.w, a1, b1)
      }
    } else {
      for (long long i = s0 + threadIdx.x; i < s1; i += NR_THREADS) NR_ACC(__ldg(xp + i), __ldg(gp + i), a0, b0)
    }
#undef NR_ACC
    double v2[2] = {(double)a0 + (double)a1, (double)b0 + (double)b1};
    cluster_sum2(v2, slot, warp_part, S);
    const float m1 = (float)(v2[0] / (double)HW), m2 = (float)(v2[1] / (double)HW);
#define NR_OUT(xv, gv) nr_gx((xv), (gv), mean, rstd, m1, m2)
    if (VEC) {
      const float4* x4 = reinterpret_cast<const float4*>(xp);
      const float4* g4 = reinterpret_cast<const float4*>(gp);
      float4* o4 = reinterpret_cast<float4*>(op);
      const long long i0 = s0 >> 2, i1 = s1 >> 2;
      long long i = i0 + threadIdx.x;
      for (; i + NR_THREADS < i1; i += 2 * NR_THREADS) {
        const float4 xa = ldg_stream(x4 + i), xb = ldg_stream(x4 + i + NR_THREADS);
        const float4 ga = ldg_stream(g4 + i), gb = ldg_stream(g4 + i + NR_THREADS);
        float4 oa, ob;
        oa.x = NR_OUT(xa.x, ga.x); oa.y = NR_OUT(xa.y, ga.y); oa.z = NR_OUT(xa.z, ga.z); oa.w = NR_OUT(xa.w, ga.w);
        ob.x = NR_OUT(xb.x, gb.x); ob.y = NR_OUT(xb.y, gb.y); ob.z = NR_OUT(xb.z, gb.z); ob.w = NR_OUT(xb.w, gb.w);
        __stcs(o4 + i, oa);                                    // evict-first: g_x must not push the lines pass 2 still needs out of L2
        __stcs(o4 + i + NR_THREADS, ob);
      }
      for (; i < i1; i += NR_THREADS) {
        const float4 xa = ldg_stream(x4 + i), ga = ldg_stream(g4 + i);
        float4 oa;
        oa.x = NR_OUT(xa.x, ga.x); oa.y = NR_OUT(xa.y, ga.y); oa.z = NR_OUT(xa.z, ga.z); oa.w = NR_OUT(xa.w, ga.w);
        __stcs(o4 + i, oa);
      }
    } else {
      for (long long i = s0 + threadIdx.x; i < s1; i += NR_THREADS) op[i] = NR_OUT(__ldg(xp + i), __ldg(gp + i));
    }
#undef NR_OUT
  }
}

// ------------------------------------------------------------------------------------------------------------------
// small planes (HW <= 4096 forward / 1024 backward, HW % 4 == 0: the deep levels of the U-Net): one WARP per plane, the plane
// stays in registers between the statistics and the output (single pass, no block-level synchronisation -- the
// CTA-per-plane kernels above spend two barriers and an idle half of the CTA on a 256-element plane)
// ------------------------------------------------------------------------------------------------------------------
constexpr int NRS_WARPS = 8;
// float4 per lane (template parameter NV): 2 -> planes of <= 256 floats, 8 -> <= 1024, 32 -> <= 4096 (forward only: the
// backward keeps two values per element)
constexpr long long NRS_MAX_HW_FWD = 32 * 32 * 4, NRS_MAX_HW_BWD = 32 * 8 * 4;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NRS_NV>
__global__ void __launch_bounds__(32 * NRS_WARPS)
vq_norm_relu_fwd_small_kernel(const float* __restrict__ x, float* __restrict__ z, float2* __restrict__ stats, long long planes,
                              int nq /* HW / 4 */, float eps) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (long long)blockIdx.x * NRS_WARPS + (threadIdx.x >> 5), nw = (long long)gridDim.x * NRS_WARPS;
  for (long long plane = w0; plane < planes; plane += nw) {
    const float4* x4 = reinterpret_cast<const float4*>(x) + plane * nq;
    float4* z4 = reinterpret_cast<float4*>(z) + plane * nq;
    float4 v[NRS_NV];
#pragma unroll
    for (int u = 0; u < NRS_NV; ++u) {
      const int i = lane + 32 * u;
      v[u] = i < nq ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float shift = __shfl_sync(0xffffffffu, v[0].x, 0);
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int u = 0; u < NRS_NV; ++u) {
      if (lane + 32 * u < nq) {
        const float dx = v[u].x - shift, dy = v[u].y - shift, dz = v[u].z - shift, dw = v[u].w - shift;
        a += (dx + dy) + (dz + dw);
        q = __fmaf_rn(dx, dx, q); q = __fmaf_rn(dy, dy, q); q = __fmaf_rn(dz, dz, q); q = __fmaf_rn(dw, dw, q);
      }
    }
    const double n = 4.0 * (double)nq;
    const double md = warp_sum_d((double)a) / n;
    double var = warp_sum_d((double)q) / n - md * md;
    if (var < 0.0) var = 0.0;
    const float mean = (float)((double)shift + md);
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (lane == 0 && stats) stats[plane] = make_float2(mean, rstd);
#pragma unroll
    for (int u = 0; u < NRS_NV; ++u) {
      const int i = lane + 32 * u;
      if (i < nq) {
        float4 o;
        o.x = fmaxf((v[u].x - mean) * rstd, 0.f); o.y = fmaxf((v[u].y - mean) * rstd, 0.f);
        o.z = fmaxf((v[u].z - mean) * rstd, 0.f); o.w = fmaxf((v[u].w - mean) * rstd, 0.f);
        z4[i] = o;
      }
    }
  }
}

template <int NRS_NV>
__global__ void __launch_bounds__(32 * NRS_WARPS)
vq_norm_relu_bwd_small_kernel(const float* __restrict__ g_z, const float* __restrict__ x, const float2* __restrict__ stats,
                              float* __restrict__ g_x, long long planes, int nq) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (long long)blockIdx.x * NRS_WARPS + (threadIdx.x >> 5), nw = (long long)gridDim.x * NRS_WARPS;
  for (long long plane = w0; plane < planes; plane += nw) {
    const float4* x4 = reinterpret_cast<const float4*>(x) + plane * nq;
    const float4* g4 = reinterpret_cast<const float4*>(g_z) + plane * nq;
    float4* o4 = reinterpret_cast<float4*>(g_x) + plane * nq;
    const float2 st = __ldg(stats + plane);
    const float mean = st.x, rstd = st.y;
    float4 xh[NRS_NV], gy[NRS_NV];              // normalised input and relu-masked gradient, kept for the second half
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int u = 0; u < NRS_NV; ++u) {
      const int i = lane + 32 * u;
      const bool in = i < nq;
      const float4 xv = in ? __ldg(x4 + i) : make_float4(mean, mean, mean, mean);
      const float4 gv = in ? __ldg(g4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      xh[u].x = (xv.x - mean) * rstd; xh[u].y = (xv.y - mean) * rstd; xh[u].z = (xv.z - mean) * rstd; xh[u].w = (xv.w - mean) * rstd;
      gy[u].x = xh[u].x > 0.f ? gv.x : 0.f; gy[u].y = xh[u].y > 0.f ? gv.y : 0.f;
      gy[u].z = xh[u].z > 0.f ? gv.z : 0.f; gy[u].w = xh[u].w > 0.f ? gv.w : 0.f;
      a += (gy[u].x + gy[u].y) + (gy[u].z + gy[u].w);
      b = __fmaf_rn(gy[u].x, xh[u].x, b); b = __fmaf_rn(gy[u].y, xh[u].y, b);
      b = __fmaf_rn(gy[u].z, xh[u].z, b); b = __fmaf_rn(gy[u].w, xh[u].w, b);
    }
    const double n = 4.0 * (double)nq;
    const float m1 = (float)(warp_sum_d((double)a) / n), m2 = (float)(warp_sum_d((double)b) / n);
#pragma unroll
    for (int u = 0; u < NRS_NV; ++u) {
      const int i = lane + 32 * u;
      if (i < nq) {
        float4 o;
        o.x = rstd * ((gy[u].x - m1) - xh[u].x * m2); o.y = rstd * ((gy[u].y - m1) - xh[u].y * m2);
        o.z = rstd * ((gy[u].z - m1) - xh[u].z * m2); o.w = rstd * ((gy[u].w - m1) - xh[u].w * m2);
        o4[i] = o;
      }
    }
  }
}

// how many clusters of each size can be co-resident (two CTAs of 512 threads per SM at best), per kernel and device
template <typename Kern>
const NrCaps& cluster_caps(Kern kern, int slot) {
  static NrCaps cached[kMaxDevices][4];
  static bool have[kMaxDevices][4] = {};
  const int dev = current_device();
  if (have[dev][slot]) return cached[dev][slot];
  NrCaps c{};
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  const bool big = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  for (int i = 0; i < NR_NSIZES; ++i) {
    const int S = 1 << i;
    if (S > 8 && !big) break;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * sms / S * S));
    cfg.blockDim = dim3(NR_THREADS);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
      if (S == 1) n = 2 * sms;                              // plain CTAs: the launch bounds guarantee two per SM
      else break;
    }
    if (n > 2 * sms / S) n = 2 * sms / S;
    c.active[i] = n;
  }
  (void)cudaGetLastError();
  cached[dev][slot] = c;
  have[dev][slot] = true;
  return cached[dev][slot];
}

template <typename Kern, typename... Args>
int launch_clustered(Kern kern, const NrPlan& p, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.clusters * p.S));
  cfg.blockDim = dim3(NR_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VQ_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, args...));
  count_launch();
  return VQ_OK;
}

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// L2 budget: with 2 x 148 CTAs in flight, a forward segment of 128 KB keeps 38 MB of x (re-read by pass 2) plus as much
// freshly written z in the 126 MB L2; a backward segment of 64 KB keeps 2 x 19 MB of x / g_z plus 19 MB of g_x.
// (256 KB / 128 KB segments measured 71 % of the HBM peak at 256^2 planes but 54-57 % at 512^2: the re-read missed.)
constexpr long long NR_SEG_FWD = 32768, NR_SEG_BWD = 16384;

int launch_norm_relu_fwd(const float* x, float* z, float* stats, long long planes, long long HW, float eps, cudaStream_t s) {
  const bool vec = (HW % 4 == 0) && al16(x) && al16(z);
  float2* st = reinterpret_cast<float2*>(stats);
  if (vec && HW <= NRS_MAX_HW_FWD) {
    const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
    long long blocks = (planes + NRS_WARPS - 1) / NRS_WARPS;
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    const int nq = (int)(HW / 4);
    if (nq <= 64) vq_norm_relu_fwd_small_kernel<2><<<(unsigned)blocks, 32 * NRS_WARPS, 0, s>>>(x, z, st, planes, nq, eps);
    else if (nq <= 256) vq_norm_relu_fwd_small_kernel<8><<<(unsigned)blocks, 32 * NRS_WARPS, 0, s>>>(x, z, st, planes, nq, eps);
    else vq_norm_relu_fwd_small_kernel<32><<<(unsigned)blocks, 32 * NRS_WARPS, 0, s>>>(x, z, st, planes, nq, eps);
    count_launch();
    VQ_CUDA_CHECK(cudaGetLastError());
    return VQ_OK;
  }
  if (vec) {
    const NrPlan p = nr_plan(planes, HW, vec, NR_SEG_FWD, cluster_caps(vq_norm_relu_fwd_kernel<true>, 0));
    return launch_clustered(vq_norm_relu_fwd_kernel<true>, p, s, x, z, st, planes, HW, p.seg, p.S, eps);
  }
  const NrPlan p = nr_plan(planes, HW, vec, NR_SEG_FWD, cluster_caps(vq_norm_relu_fwd_kernel<false>, 1));
  return launch_clustered(vq_norm_relu_fwd_kernel<false>, p, s, x, z, st, planes, HW, p.seg, p.S, eps);
}

int launch_norm_relu_bwd(const float* g_z, const float* x, const float* stats, float* g_x, long long planes, long long HW,
                         cudaStream_t s) {
  const bool vec = (HW % 4 == 0) && al16(x) && al16(g_z) && al16(g_x);
  const float2* st = reinterpret_cast<const float2*>(stats);
  if (vec && HW <= NRS_MAX_HW_BWD) {
    const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
    long long blocks = (planes + NRS_WARPS - 1) / NRS_WARPS;
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    const int nq = (int)(HW / 4);
    if (nq <= 64) vq_norm_relu_bwd_small_kernel<2><<<(unsigned)blocks, 32 * NRS_WARPS, 0, s>>>(g_z, x, st, g_x, planes, nq);
    else vq_norm_relu_bwd_small_kernel<8><<<(unsigned)blocks, 32 * NRS_WARPS, 0, s>>>(g_z, x, st, g_x, planes, nq);
    count_launch();
    VQ_CUDA_CHECK(cudaGetLastError());
    return VQ_OK;
  }
  if (vec) {
    const NrPlan p = nr_plan(planes, HW, vec, NR_SEG_BWD, cluster_caps(vq_norm_relu_bwd_kernel<true>, 2));
    return launch_clustered(vq_norm_relu_bwd_kernel<true>, p, s, g_z, x, st, g_x, planes, HW, p.seg, p.S);
  }
  const NrPlan p = nr_plan(planes, HW, vec, NR_SEG_BWD, cluster_caps(vq_norm_relu_bwd_kernel<false>, 3));
  return launch_clustered(vq_norm_relu_bwd_kernel<false>, p, s, g_z, x, st, g_x, planes, HW, p.seg, p.S);
}

}  // namespace vqb200

// vq_kernels.cu -- fp32 CUDA-core kernels of the VQ bottleneck (sm_100a):
//   prep        |e|^2, transposed codebook, snapshot, accumulator zeroing
//   assign_simt exact fp32 nearest-code search (+ gather, loss, EMA statistics) -- the first
//               correct path and the exact fallback of the tensor-core path
//   finish      loss finalisation + packing of the int32 histogram
//   ema         EMA / Laplace-smoothing / codebook refresh            (vq_module.py:194-199)
//   bwd         straight-through + commitment-loss backward          (grad_approximation.py:7-29)
//   lookup      codebook gather                                      (vq_module.py:203-206)
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "vq_common.cuh"

namespace vqb200 {

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = {0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// ---------------------------------------------------------------------------------------------
// measurement hooks
// ---------------------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t g_prof_open = nullptr;

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(); }
void profile_enable(bool on) { g_prof_on.store(on); }

static cudaEvent_t prof_get_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  return e;
}
bool profile_begin(cudaStream_t s) {
  if (!g_prof_on.load()) return false;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_events.size() >= 8192) return false;
  g_prof_open = prof_get_event();
  if (!g_prof_open) return false;
  cudaEventRecord(g_prof_open, s);
  return true;
}
void profile_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_open) return;
  cudaEvent_t e = prof_get_event();
  if (e) { cudaEventRecord(e, s); g_prof_events.emplace_back(g_prof_open, e); }
  else g_prof_pool.push_back(g_prof_open);
  g_prof_open = nullptr;
}
int profile_read(double* total_ms, int* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double tot = 0.0; int n = 0;
  for (auto& pr : g_prof_events) {
    if (cudaEventSynchronize(pr.second) == cudaSuccess) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { tot += ms; ++n; }
    }
    g_prof_pool.push_back(pr.first); g_prof_pool.push_back(pr.second);
  }
  g_prof_events.clear();
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return VQ_OK;
}

// =============================================================================================
// prep
// =============================================================================================
__global__ void vq_prep_kernel(const float* __restrict__ E, int K, int D, int Kpad,
                               float* __restrict__ e2, float* __restrict__ et, float* __restrict__ snap,
                               float* __restrict__ stats, size_t stats_n, float* __restrict__ sums_rep, size_t rep_n,
                               int* __restrict__ counts, double* __restrict__ loss_acc, int* __restrict__ misc,
                               uint32_t* __restrict__ tc_rmax, uint32_t* __restrict__ tc_meta) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  // |e_k|^2 : one thread per code, ascending d, fused multiply-add chain
  for (size_t k = tid; k < (size_t)Kpad; k += nth) {
    float acc = 0.f;
    if (k < (size_t)K) {
      const float* row = E + k * D;
      int d = 0;
      if ((D & 3) == 0 && (((uintptr_t)E) & 15) == 0) {      // four 16-byte loads in flight, then their fmas in ascending d
        const float4* r4 = reinterpret_cast<const float4*>(row);
        for (; d + 16 <= D; d += 16) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = __ldg(r4 + (d >> 2) + u);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc = __fmaf_rn(v[u].x, v[u].x, acc); acc = __fmaf_rn(v[u].y, v[u].y, acc);
            acc = __fmaf_rn(v[u].z, v[u].z, acc); acc = __fmaf_rn(v[u].w, v[u].w, acc);
          }
        }
      }
      for (; d < D; ++d) acc = __fmaf_rn(row[d], row[d], acc);
    } else {
      acc = INFINITY;  // padding codes can never win: score = 2*0 - inf
    }
    e2[k] = acc;
  }
  // transposed codebook (coalesced writes) and snapshot
  const size_t tot = (size_t)D * Kpad;
  for (size_t i = tid; i < tot; i += nth) {
    const size_t d = i / Kpad, k = i % Kpad;
    et[i] = (k < (size_t)K) ? E[k * D + d] : 0.f;
  }
  if (snap) {
    const size_t kd = (size_t)K * D;
    for (size_t i = tid; i < kd; i += nth) snap[i] = E[i];
  }
  if (stats) {
    for (size_t i = tid; i < stats_n; i += nth) stats[i] = 0.f;
    for (size_t i = tid; i < rep_n; i += nth) sums_rep[i] = 0.f;
  }
  for (size_t i = tid; i < (size_t)K; i += nth) counts[i] = 0;
  if (tid == 0) *loss_acc = 0.0;
  if (tid < 8) misc[tid] = 0;
  // tensor-core path: per-chunk largest norm (atomicMax target) and the smallest excluded norm (atomicMin target)
  for (size_t i = tid; i < (size_t)Kpad / 32; i += nth) tc_rmax[i] = 0u;
  if (tid < 16) tc_meta[tid] = 0x7F800000u;
}

// =============================================================================================
// exact fp32 search
// =============================================================================================
constexpr int FB_ROWS_MAX = 8192; // longer fallback lists go to the tiled search
constexpr int SIMT_TP = 128;   // pixels per tile
constexpr int SIMT_TK = 64;    // codes per pass
constexpr int SIMT_DC = 32;    // channel chunk

// One CTA (256 threads) = 128 pixels.  Thread (tx = tid&31, ty = tid>>5) owns pixels 4tx..4tx+3
// and, in each pass over 64 codes, codes 8ty..8ty+7: a 4x8 register tile of dot products,
// accumulated over ascending d with two fma chains per (pixel, code): even / odd channel quads (exact_dot, vq_common.cuh).
__global__ void __launch_bounds__(256)
vq_assign_simt_kernel(const float* __restrict__ z, const float* __restrict__ et, const float* __restrict__ e2,
                      const float* __restrict__ E, int B, int D, int H, int W, int K, int Kpad,
                      const int* __restrict__ fb_rows, const int* __restrict__ fb_count,
                      int64_t* __restrict__ ids, int32_t* __restrict__ ids_nat, float* __restrict__ q,
                      double* __restrict__ loss_acc, int* __restrict__ counts, float* __restrict__ sums, int ids_mode) {
  __shared__ __align__(16) float zs[SIMT_DC][SIMT_TP];
  __shared__ __align__(16) float es[SIMT_DC][SIMT_TK];
  __shared__ float e2s[SIMT_TK];
  __shared__ float red_s[8][SIMT_TP];
  __shared__ int red_i[8][SIMT_TP];
  __shared__ long long pix_off[SIMT_TP];   // offset of z[b,0,p]; -1 = masked

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int HW = H * W;
  const long long N = (long long)B * HW;

  long long ntiles;
  const bool list_mode = (fb_rows != nullptr);
  int nrows = 0;
  if (list_mode) {
    nrows = *fb_count;
    if (nrows <= FB_ROWS_MAX) return;               // short lists were handled by vq_fallback_rows_kernel
    ntiles = (nrows + SIMT_TP - 1) / SIMT_TP;
  } else {
    ntiles = (long long)B * ((HW + SIMT_TP - 1) / SIMT_TP);
  }
  const int tiles_per_img = (HW + SIMT_TP - 1) / SIMT_TP;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < SIMT_TP) {
      long long off = -1;
      if (list_mode) {
        const long long r = tile * SIMT_TP + tid;
        if (r < nrows) {
          const long long n = fb_rows[r];
          const long long b = n / HW, p = n % HW;
          off = b * (long long)D * HW + p;
        }
      } else {
        const long long b = tile / tiles_per_img;
        const long long p = (tile % tiles_per_img) * SIMT_TP + tid;
        if (p < HW) off = b * (long long)D * HW + p;
      }
      pix_off[tid] = off;
    }
    __syncthreads();

    float best[4];
    int bidx[4];
    float z2[4] = {0.f, 0.f, 0.f, 0.f}, z2B[4] = {0.f, 0.f, 0.f, 0.f};   // |z|^2 = A + B (even / odd channel quads)
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[i] = -INFINITY; bidx[i] = 0; }

    // loader mapping: pixel i = tid & 127, rows dd = (tid >> 7) + 2*j
    const int li = tid & (SIMT_TP - 1);
    const long long loff = pix_off[li];

    for (int cb = 0; cb < Kpad; cb += SIMT_TK) {
      // exact dot product = A + B: one fma chain over the even channel quads, one over the odd ones (vq_common.cuh)
      float acc[4][8], accB[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) { acc[i][c] = 0.f; accB[i][c] = 0.f; }

      for (int d0 = 0; d0 < D; d0 += SIMT_DC) {
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < SIMT_DC / 2; ++j) {
          const int dd = (tid >> 7) + 2 * j;
          const int d = d0 + dd;
          float v = 0.f;
          if (d < D && loff >= 0) v = __ldg(z + loff + (long long)d * HW);
          zs[dd][li] = v;
        }
#pragma unroll
        for (int j = 0; j < (SIMT_DC * SIMT_TK) / 256; ++j) {
          const int e = tid + j * 256;
          const int dd = e / SIMT_TK, kk = e % SIMT_TK;
          const int d = d0 + dd;
          es[dd][kk] = (d < D) ? __ldg(et + (size_t)d * Kpad + cb + kk) : 0.f;
        }
        if (d0 == 0 && tid < SIMT_TK) e2s[tid] = e2[cb + tid];
        __syncthreads();

        const int dlim = min(SIMT_DC, D - d0);
        if (cb == 0) {
          for (int dd = 0; dd < dlim; ++dd) {              // d0 is a multiple of 32: quad parity = (dd >> 2) & 1
            const float4 zv = *reinterpret_cast<const float4*>(&zs[dd][tx * 4]);
            if (((dd >> 2) & 1) == 0) {
              z2[0] = __fmaf_rn(zv.x, zv.x, z2[0]);
              z2[1] = __fmaf_rn(zv.y, zv.y, z2[1]);
              z2[2] = __fmaf_rn(zv.z, zv.z, z2[2]);
              z2[3] = __fmaf_rn(zv.w, zv.w, z2[3]);
            } else {
              z2B[0] = __fmaf_rn(zv.x, zv.x, z2B[0]);
              z2B[1] = __fmaf_rn(zv.y, zv.y, z2B[1]);
              z2B[2] = __fmaf_rn(zv.z, zv.z, z2B[2]);
              z2B[3] = __fmaf_rn(zv.w, zv.w, z2B[3]);
            }
          }
          if (d0 + SIMT_DC >= D) {
#pragma unroll
            for (int i = 0; i < 4; ++i) z2[i] = __fadd_rn(z2[i], z2B[i]);
          }
        }
        for (int d8 = 0; d8 < dlim; d8 += 8) {            // d0 is a multiple of 32: quad parity = (dd >> 2) & 1
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int dd = d8 + u;
            if (dd < dlim) {
              const float4 zv = *reinterpret_cast<const float4*>(&zs[dd][tx * 4]);
              const float4 ea = *reinterpret_cast<const float4*>(&es[dd][ty * 8]);
              const float4 eb = *reinterpret_cast<const float4*>(&es[dd][ty * 8 + 4]);
              const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
              const float ee[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  if (u < 4) acc[i][c] = __fmaf_rn(zz[i], ee[c], acc[i][c]);
                  else accB[i][c] = __fmaf_rn(zz[i], ee[c], accB[i][c]);
                }
            }
          }
        }
      }
      // scores for this block of codes, reference op order; strict '>' keeps the lowest index
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int k = cb + ty * 8 + c;
        const float ek = e2s[ty * 8 + c];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float s = ref_score(__fadd_rn(acc[i][c], accB[i][c]), ek, z2[i]);
          if (s > best[i]) { best[i] = s; bidx[i] = k; }
        }
      }
    }

    // reduce across the 8 code groups of each pixel
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      red_s[ty][tx * 4 + i] = best[i];
      red_i[ty][tx * 4 + i] = bidx[i];
    }
    __syncthreads();

    float lsum = 0.f;
    if (tid < SIMT_TP) {
      const long long off = pix_off[tid];
      if (off >= 0) {
        float bs = red_s[0][tid];
        int bi = red_i[0][tid];
#pragma unroll
        for (int g = 1; g < 8; ++g) {
          const float s = red_s[g][tid];
          const int i2 = red_i[g][tid];
          if (s > bs || (s == bs && i2 < bi)) { bs = s; bi = i2; }
        }
        const long long b = off / ((long long)D * HW);
        const long long p = off - b * (long long)D * HW;
        const int h = (int)(p / W), w = (int)(p % W);
        if (ids) store_id(ids + b * HW, p, h, w, H, bi, ids_mode);
        if (ids_nat) ids_nat[b * HW + p] = bi;
        if (counts) atomicAdd(&counts[bi], 1);
        const float* erow = E + (size_t)bi * D;
        float* srow = sums ? sums + (size_t)bi * D : nullptr;
        if ((D & 3) == 0) {
          for (int d = 0; d < D; d += 4) {
            const float4 e4 = __ldg(reinterpret_cast<const float4*>(erow + d));
            const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
            float zv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) zv[c] = __ldg(z + off + (long long)(d + c) * HW);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float df = zv[c] - ev[c];
              lsum = __fmaf_rn(df, df, lsum);
              if (q) q[off + (long long)(d + c) * HW] = ev[c];
            }
            if (srow) atomicAdd(reinterpret_cast<float4*>(srow + d), make_float4(zv[0], zv[1], zv[2], zv[3]));
          }
        } else {
          for (int d = 0; d < D; ++d) {
            const float ev = __ldg(erow + d);
            const float zv = __ldg(z + off + (long long)d * HW);
            const float df = zv - ev;
            lsum = __fmaf_rn(df, df, lsum);
            if (q) q[off + (long long)d * HW] = ev;
            if (srow) atomicAdd(srow + d, zv);
          }
        }
      }
    }
    if (tid < SIMT_TP) {   // warps 0..3, warp-uniform branch
      lsum = warp_sum(lsum);
      if ((tid & 31) == 0 && loss_acc) atomicAdd(loss_acc, (double)lsum);
    }
  }
  (void)N;
}

// =============================================================================================
// exhaustive exact search for the few rows the tensor-core kernel could not decide: one CTA per row, codes
// spread over the threads (coalesced reads of the transposed codebook), same fma chains as above.  Lists longer
// than FB_ROWS_MAX are left to the tiled search (vq_assign_simt_kernel in list mode), which is launched right after.
// =============================================================================================
constexpr int FB_THREADS = 256;
__global__ void __launch_bounds__(FB_THREADS)
vq_fallback_rows_kernel(const float* __restrict__ z, const float* __restrict__ et, const float* __restrict__ e2,
                        const float* __restrict__ E, int D, int H, int W, int K, int Kpad,
                        const int* __restrict__ fb_rows, const int* __restrict__ fb_count,
                        int64_t* __restrict__ ids, int32_t* __restrict__ ids_nat, float* __restrict__ q,
                        double* __restrict__ loss_acc, int* __restrict__ counts, float* __restrict__ sums, int ids_mode) {
  extern __shared__ float zr[];                       // [D]
  __shared__ float red_s[FB_THREADS / 32];
  __shared__ int red_i[FB_THREADS / 32];
  __shared__ int s_best;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int nrows = *fb_count;
  if (nrows > FB_ROWS_MAX) return;                    // the tiled list search handles long lists
  const int HW = H * W;
  float lsum = 0.f;
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const long long n = fb_rows[r];
    const long long b = n / HW, p = n - b * HW;
    const long long off = b * (long long)D * HW + p;
    __syncthreads();
    for (int d = tid; d < D; d += FB_THREADS) zr[d] = __ldg(z + off + (long long)d * HW);
    __syncthreads();
    float z2 = 0.f, z2B = 0.f;                       // |z|^2 = A + B (even / odd channel quads, vq_common.cuh)
    for (int d = 0; d < D; ++d) {
      if (((d >> 2) & 1) == 0) z2 = __fmaf_rn(zr[d], zr[d], z2);
      else z2B = __fmaf_rn(zr[d], zr[d], z2B);
    }
    z2 = __fadd_rn(z2, z2B);
    float best = -INFINITY;
    int bi = 0;
    for (int k0 = 0; k0 < Kpad; k0 += 2 * FB_THREADS) {     // two codes per thread in flight: k, k + 256
      const bool two = k0 + FB_THREADS < Kpad;
      float acc0 = 0.f, acc1 = 0.f, acc0B = 0.f, acc1B = 0.f;   // dot = A + B (even / odd channel quads, vq_common.cuh)
      const float* ep = et + k0 + tid;
      // Full groups of 16 channels: all 32 loads first (unguarded, so the compiler keeps them in flight together), then
      // the fmas in the library's chain order.  With the loads guarded one by one (the first version) every load was a
      // separate L2 round trip: 36 us per call for one or two rows at D = 64, 8 % of the training step.
      int d8 = 0;
      for (; d8 + 16 <= D; d8 += 16) {
        float ea[16], eb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float* row = ep + (size_t)(d8 + u) * Kpad;
          ea[u] = __ldg(row);
          eb[u] = two ? __ldg(row + FB_THREADS) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float zv = zr[d8 + u];
          if ((u & 7) < 4) {
            acc0 = __fmaf_rn(zv, ea[u], acc0);
            if (two) acc1 = __fmaf_rn(zv, eb[u], acc1);
          } else {
            acc0B = __fmaf_rn(zv, ea[u], acc0B);
            if (two) acc1B = __fmaf_rn(zv, eb[u], acc1B);
          }
        }
      }
      for (; d8 < D; d8 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int d = d8 + u;
          if (d < D) {
            const float zv = zr[d];
            const float* row = ep + (size_t)d * Kpad;
            if (u < 4) {
              acc0 = __fmaf_rn(zv, __ldg(row), acc0);
              if (two) acc1 = __fmaf_rn(zv, __ldg(row + FB_THREADS), acc1);
            } else {
              acc0B = __fmaf_rn(zv, __ldg(row), acc0B);
              if (two) acc1B = __fmaf_rn(zv, __ldg(row + FB_THREADS), acc1B);
            }
          }
        }
      }
      acc0 = __fadd_rn(acc0, acc0B);
      acc1 = __fadd_rn(acc1, acc1B);
      {                                                      // ascending k per thread, strict '>' keeps the lowest index
        const int k = k0 + tid;
        const float sc = ref_score(acc0, __ldg(e2 + k), z2);
        if (sc > best) { best = sc; bi = k; }
      }
      if (two) {
        const int k = k0 + tid + FB_THREADS;
        const float sc = ref_score(acc1, __ldg(e2 + k), z2);
        if (sc > best) { best = sc; bi = k; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { red_s[wib] = best; red_i[wib] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w2 = 1; w2 < FB_THREADS / 32; ++w2) {
        const float ob = red_s[w2];
        const int oi = red_i[w2];
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (!(best > -INFINITY)) bi = 0;                 // all-NaN row: the reference's topk returns index 0
      s_best = bi;
      const int h = (int)(p / W), w = (int)(p % W);
      if (ids) store_id(ids + b * HW, p, h, w, H, bi, ids_mode);
      if (ids_nat) ids_nat[n] = bi;
      if (counts) atomicAdd(&counts[bi], 1);
    }
    __syncthreads();
    bi = s_best;
    for (int d = tid; d < D; d += FB_THREADS) {
      const float ev = __ldg(E + (size_t)bi * D + d);
      const float df = zr[d] - ev;
      lsum = __fmaf_rn(df, df, lsum);
      if (q) q[off + (long long)d * HW] = ev;
      if (sums) atomicAdd(sums + (size_t)bi * D + d, zr[d]);
    }
  }
  lsum = warp_sum(lsum);
  if (lane == 0 && loss_acc && lsum != 0.f) atomicAdd(loss_acc, (double)lsum);
}

// =============================================================================================
// finish: loss = acc / (N*D); pack the int32 histogram as two exactly-representable floats
// =============================================================================================
__global__ void vq_finish_kernel(const double* __restrict__ loss_acc, float* __restrict__ loss, double inv_numel,
                                 const int* __restrict__ counts, float* __restrict__ stats, int K, int D,
                                 size_t sums_off, const float* __restrict__ sums_rep, int nrep) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0 && loss) *loss = (float)(*loss_acc * inv_numel);
  if (stats) {
    for (int k = tid; k < K; k += gridDim.x * blockDim.x) {
      const int c = counts[k];
      stats[k] = (float)(c >> 12);
      stats[K + k] = (float)(c & 4095);
    }
    if (nrep > 1) {                       // fold the replicas of the per-code sums (fixed order)
      const int kd = K * D;
      for (int i = tid; i < kd; i += gridDim.x * blockDim.x) {
        float acc = stats[sums_off + i];
        int r = 0;
        for (; r + 8 <= nrep - 1; r += 8) {                   // eight replica loads in flight, added in the fixed order
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldcs(sums_rep + (size_t)(r + u) * kd + i);
#pragma unroll
          for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; r < nrep - 1; ++r) acc += sums_rep[(size_t)r * kd + i];
        stats[sums_off + i] = acc;
      }
    }
  }
}

// =============================================================================================
// EMA update (vq_module.py:132-136, 194-199)
// =============================================================================================
__global__ void __launch_bounds__(1024)
vq_ema_cs_kernel(float* __restrict__ cluster_size, const float* __restrict__ stats, int K, float momentum,
                 float alpha, float count_scale, float* __restrict__ scratch) {
  __shared__ double red[32];
  double part = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float cnt = __fmaf_rn(stats[k], 4096.0f, stats[K + k]);
    cnt = cnt * count_scale;
    const float cs = __fmaf_rn(alpha, cnt, __fmul_rn(cluster_size[k], momentum));   // base.mul_(m).add_(u, alpha=1-m)
    cluster_size[k] = cs;
    part += (double)cs;
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) scratch[0] = (float)v;   // n = cluster_size.sum()
  }
}

__global__ void __launch_bounds__(256)
vq_ema_embed_kernel(const float* __restrict__ cluster_size, float* __restrict__ embed_avg, long long sd, long long sk,
                    float* __restrict__ embed, const float* __restrict__ sums, int K, int D, float momentum,
                    float alpha, float sum_scale, float eps, float k_eps, const float* __restrict__ scratch) {
  // 32x32 tile transpose: embed_avg addressed [d*sd + k*sk] (coalesced when sk == 1); sums / embed are [K][D]
  __shared__ float t_in[32][33];
  __shared__ float t_out[32][33];
  const int k0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;   // 32 x 8
  const float n = scratch[0];
  for (int r = ly; r < 32; r += 8) {            // read sums[k0+r][d0+lx]
    const int k = k0 + r, d = d0 + lx;
    t_in[r][lx] = (k < K && d < D) ? sums[(size_t)k * D + d] * sum_scale : 0.f;
  }
  __syncthreads();
  for (int r = ly; r < 32; r += 8) {            // d = d0 + r, k = k0 + lx
    const int d = d0 + r, k = k0 + lx;
    if (k < K && d < D) {
      const size_t ia = (size_t)d * sd + (size_t)k * sk;
      const float avg = __fmaf_rn(alpha, t_in[lx][r], __fmul_rn(embed_avg[ia], momentum));
      embed_avg[ia] = avg;
      const float cs = __fdiv_rn(__fmul_rn(n, __fadd_rn(cluster_size[k], eps)), __fadd_rn(n, k_eps));
      t_out[lx][r] = __fdiv_rn(avg, cs);
    }
  }
  __syncthreads();
  for (int r = ly; r < 32; r += 8) {            // write embed[k0+r][d0+lx]
    const int k = k0 + r, d = d0 + lx;
    if (k < K && d < D) embed[(size_t)k * D + d] = t_out[r][lx];
  }
}

// embed_avg stored code-major (strides (1, D): what `embed.T.clone()` produces): pure element-wise
__global__ void __launch_bounds__(256)
vq_ema_embed_kd_kernel(const float* __restrict__ cluster_size, float* __restrict__ embed_avg, float* __restrict__ embed,
                       const float* __restrict__ sums, int K, int D, float momentum, float alpha, float sum_scale,
                       float eps, float k_eps, const float* __restrict__ scratch) {
  const float n = scratch[0];
  const size_t tot = (size_t)K * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / D);
    const float avg = __fmaf_rn(alpha, sums[i] * sum_scale, __fmul_rn(embed_avg[i], momentum));
    embed_avg[i] = avg;
    const float cs = __fdiv_rn(__fmul_rn(n, __fadd_rn(cluster_size[k], eps)), __fadd_rn(n, k_eps));
    embed[i] = __fdiv_rn(avg, cs);
  }
}

// =============================================================================================
// backward:  g_z = g_q + g_loss * 2 (z - E[id]) / numel
// =============================================================================================
// vector path: thread = 4 consecutive pixels x 4 channels per step
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
vq_bwd_vec_kernel(const float* __restrict__ g_q, const float* __restrict__ g_loss, const float* __restrict__ z,
                  const int32_t* __restrict__ ids_nat, const float* __restrict__ E, float* __restrict__ g_z,
                  int D, int HW, long long nquads, int dsplit, float two_over_numel) {
  const long long quad = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (quad >= nquads) return;
  const int qpi = HW >> 2;
  const long long b = quad / qpi;
  const int p = (int)(quad % qpi) << 2;
  const float coef = g_loss ? (*g_loss) * two_over_numel : 0.f;
  const int4 id4 = *reinterpret_cast<const int4*>(ids_nat + b * HW + p);
  const float* e0 = E + (size_t)id4.x * D;
  const float* e1 = E + (size_t)id4.y * D;
  const float* e2 = E + (size_t)id4.z * D;
  const float* e3 = E + (size_t)id4.w * D;
  const int dper = D / dsplit;                 // multiple of 4 by construction
  const int dbeg = blockIdx.y * dper, dend = dbeg + dper;
  const long long base = b * (long long)D * HW + p;
  for (int d = dbeg; d < dend; d += 4) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(e0 + d));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(e1 + d));
    const float4 a2 = __ldg(reinterpret_cast<const float4*>(e2 + d));
    const float4 a3 = __ldg(reinterpret_cast<const float4*>(e3 + d));
    const float ev[4][4] = {{a0.x, a1.x, a2.x, a3.x}, {a0.y, a1.y, a2.y, a3.y},
                            {a0.z, a1.z, a2.z, a3.z}, {a0.w, a1.w, a2.w, a3.w}};
    float4 zv[4], gv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long o = base + (long long)(d + c) * HW;
      zv[c] = __ldcs(reinterpret_cast<const float4*>(z + o));
      gv[c] = g_q ? __ldcs(reinterpret_cast<const float4*>(g_q + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 r;
      r.x = __fmaf_rn(coef, zv[c].x - ev[c][0], gv[c].x);
      r.y = __fmaf_rn(coef, zv[c].y - ev[c][1], gv[c].y);
      r.z = __fmaf_rn(coef, zv[c].z - ev[c][2], gv[c].z);
      r.w = __fmaf_rn(coef, zv[c].w - ev[c][3], gv[c].w);
      __stcs(reinterpret_cast<float4*>(g_z + base + (long long)(d + c) * HW), r);
    }
  }
}

// generic path: thread = one pixel, all channels
__global__ void __launch_bounds__(256)
vq_bwd_generic_kernel(const float* __restrict__ g_q, const float* __restrict__ g_loss, const float* __restrict__ z,
                      const int32_t* __restrict__ ids_nat, const float* __restrict__ E, float* __restrict__ g_z,
                      int D, int HW, long long N, float two_over_numel) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long b = n / HW;
  const int p = (int)(n % HW);
  const float coef = g_loss ? (*g_loss) * two_over_numel : 0.f;
  const float* e = E + (size_t)ids_nat[n] * D;
  const long long base = b * (long long)D * HW + p;
  for (int d = 0; d < D; ++d) {
    const long long o = base + (long long)d * HW;
    const float g = g_q ? g_q[o] : 0.f;
    g_z[o] = __fmaf_rn(coef, z[o] - __ldg(e + d), g);
  }
}

// =============================================================================================
// lookup
// =============================================================================================
__global__ void __launch_bounds__(256)
vq_lookup_rows_kernel(const int64_t* __restrict__ ids, long long n, const float* __restrict__ E, int K, int D,
                      float* __restrict__ out, int* __restrict__ status) {
  // thread = (row, 4-channel group) when D%4==0, else (row, channel)
  const int vec = ((D & 3) == 0) ? 4 : 1;
  const int gpr = D / vec;
  const long long total = n * gpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / gpr;
    const int g = (int)(i % gpr);
    const long long id = ids[r];
    const bool ok = (id >= 0 && id < K);
    if (!ok && status) *status = 1;
    if (vec == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v = __ldg(reinterpret_cast<const float4*>(E + (size_t)id * D) + g);
      __stcs(reinterpret_cast<float4*>(out + r * D) + g, v);
    } else {
      out[r * D + g] = ok ? __ldg(E + (size_t)id * D + g) : 0.f;
    }
  }
}

// ids [B,A,C] -> out [B,D,C,A]: out[b,d,c,a] = E[ids[b,a,c]][d].  Tile: 32 a x 32 c per CTA so that
// both the id reads (c fastest) and the output writes (a fastest) are coalesced.
template <bool VEC>
__global__ void __launch_bounds__(256)
vq_lookup_nchw_t_kernel(const int64_t* __restrict__ ids, const float* __restrict__ E, int K, int D,
                        float* __restrict__ out, int B, int A, int C, int* __restrict__ status) {
  __shared__ int sid[32][33];
  const int a0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int r = ly; r < 32; r += 8) {     // a = a0 + r, c = c0 + lx
    const int a = a0 + r, c = c0 + lx;
    int v = -1;
    if (a < A && c < C) {
      const long long id = ids[((long long)b * A + a) * C + c];
      if (id >= 0 && id < K) v = (int)id;
      else if (status) *status = 1;
    }
    sid[r][lx] = v;
  }
  __syncthreads();
  if (VEC) {
    // A % 4 == 0, out 16-byte aligned: thread = four consecutive a of one c, one streaming float4 store per channel
    const int a4 = (threadIdx.x & 7) * 4, cc = threadIdx.x >> 3;
    const int a = a0 + a4, c = c0 + cc;
    if (a < A && c < C) {                 // A % 4 == 0: the quad is inside or outside as a whole
      const int i0 = sid[a4][cc], i1 = sid[a4 + 1][cc], i2 = sid[a4 + 2][cc], i3 = sid[a4 + 3][cc];
      const float* e0 = E + (size_t)(i0 < 0 ? 0 : i0) * D;
      const float* e1 = E + (size_t)(i1 < 0 ? 0 : i1) * D;
      const float* e2 = E + (size_t)(i2 < 0 ? 0 : i2) * D;
      const float* e3 = E + (size_t)(i3 < 0 ? 0 : i3) * D;
      float* o = out + (((long long)b * D) * C + c) * A + a;
      const long long dstride = (long long)C * A;
      int d = 0;
      if ((D & 3) == 0 && ((((uintptr_t)E) & 15) == 0)) {
        // four channels at a time: one 16-byte load per code row (the codebook sits in L1), a 4 x 4 transpose in
        // registers, four 16-byte streaming stores -- 2.5 x fewer load/store instructions than a scalar gather
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (; d + 4 <= D; d += 4) {
          const float4 r0 = i0 < 0 ? zero : __ldg(reinterpret_cast<const float4*>(e0 + d));
          const float4 r1 = i1 < 0 ? zero : __ldg(reinterpret_cast<const float4*>(e1 + d));
          const float4 r2 = i2 < 0 ? zero : __ldg(reinterpret_cast<const float4*>(e2 + d));
          const float4 r3 = i3 < 0 ? zero : __ldg(reinterpret_cast<const float4*>(e3 + d));
          __stcs(reinterpret_cast<float4*>(o + (d + 0) * dstride), make_float4(r0.x, r1.x, r2.x, r3.x));
          __stcs(reinterpret_cast<float4*>(o + (d + 1) * dstride), make_float4(r0.y, r1.y, r2.y, r3.y));
          __stcs(reinterpret_cast<float4*>(o + (d + 2) * dstride), make_float4(r0.z, r1.z, r2.z, r3.z));
          __stcs(reinterpret_cast<float4*>(o + (d + 3) * dstride), make_float4(r0.w, r1.w, r2.w, r3.w));
        }
      }
#pragma unroll 4
      for (; d < D; ++d) {
        float4 v;
        v.x = i0 < 0 ? 0.f : __ldg(e0 + d);
        v.y = i1 < 0 ? 0.f : __ldg(e1 + d);
        v.z = i2 < 0 ? 0.f : __ldg(e2 + d);
        v.w = i3 < 0 ? 0.f : __ldg(e3 + d);
        __stcs(reinterpret_cast<float4*>(o + d * dstride), v);
      }
    }
  } else {
  // thread: a = a0 + lx (fastest in output), c = c0 + r
  for (int r = ly; r < 32; r += 8) {
    const int a = a0 + lx, c = c0 + r;
    if (a < A && c < C) {
      const int id = sid[lx][r];
      const float* e = E + (size_t)(id < 0 ? 0 : id) * D;
      float* o = out + (((long long)b * D) * C + c) * A + a;
      const long long dstride = (long long)C * A;
      for (int d = 0; d < D; ++d) o[d * dstride] = (id < 0) ? 0.f : __ldg(e + d);
    }
  }
  }
}

// =============================================================================================
// launchers
// =============================================================================================
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int device_sm_count() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

static int sm_count() { return device_sm_count(); }

int tuning_knob(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

int launch_prep(const FwdArgs& a, bool tc_path, cudaStream_t s) {
  const int Kpad = pad_codes(a.K);
  const size_t work = (size_t)a.D * Kpad;
  int blocks = (int)((work + 255) / 256);
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  if (blocks < 1) blocks = 1;
  const size_t stats_n = a.stats ? vq_stats_floats(a.K, a.D) : 0;
  const size_t rep_n = (a.stats && tc_path) ? (size_t)(tc_sums_replicas(a.K, a.D) - 1) * a.K * a.D : 0;
  if (rep_n > 0 && blocks < 2 * sm_count()) blocks = 2 * sm_count();
  vq_prep_kernel<<<blocks, 256, 0, s>>>(a.embed, a.K, a.D, Kpad, a.ws.e2, a.ws.et, a.snapshot, a.stats, stats_n,
                                        a.ws.sums_rep, rep_n, a.ws.counts, a.ws.loss_acc, a.ws.misc,
                                        reinterpret_cast<uint32_t*>(a.ws.tc_ctab), reinterpret_cast<uint32_t*>(a.ws.tc_meta));
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

int launch_assign_simt(const FwdArgs& a, bool fallback_list_mode, cudaStream_t s) {
  const int HW = a.H * a.W;
  const int Kpad = pad_codes(a.K);
  long long ntiles = (long long)a.B * ((HW + SIMT_TP - 1) / SIMT_TP);
  int blocks;
  if (fallback_list_mode) {
    blocks = 2 * sm_count();          // grid-stride over a device-side row count
  } else {
    blocks = (int)(ntiles < (long long)8 * sm_count() ? ntiles : (long long)8 * sm_count());
  }
  if (blocks < 1) blocks = 1;
  float* sums = a.stats ? a.stats + stats_sums_offset(a.K) : nullptr;
  const bool prof = !fallback_list_mode && profile_begin(s);
  vq_assign_simt_kernel<<<blocks, 256, 0, s>>>(a.z, a.ws.et, a.ws.e2, a.embed, a.B, a.D, a.H, a.W, a.K, Kpad,
                                               fallback_list_mode ? a.ws.fb_rows : nullptr, a.ws.misc,
                                               a.ids, a.ids_nat, a.q, a.ws.loss_acc,
                                               a.stats ? a.ws.counts : nullptr, sums, ids_mode_of(a.flags));
  if (prof) profile_end(s);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

int launch_fallback_rows(const FwdArgs& a, cudaStream_t s) {
  const int Kpad = pad_codes(a.K);
  float* sums = a.stats ? a.stats + stats_sums_offset(a.K) : nullptr;
  const size_t smem = (size_t)a.D * sizeof(float);
  // short lists: one CTA per row (latency-bound, finishes in a few microseconds) ...
  vq_fallback_rows_kernel<<<4 * sm_count(), FB_THREADS, smem, s>>>(
      a.z, a.ws.et, a.ws.e2, a.embed, a.D, a.H, a.W, a.K, Kpad, a.ws.fb_rows, a.ws.misc, a.ids, a.ids_nat, a.q,
      a.ws.loss_acc, a.stats ? a.ws.counts : nullptr, sums, ids_mode_of(a.flags));
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  // ... long lists (degenerate codebooks): the tiled fp32 search over the listed rows; exits at once otherwise
  vq_assign_simt_kernel<<<2 * sm_count(), 256, 0, s>>>(a.z, a.ws.et, a.ws.e2, a.embed, a.B, a.D, a.H, a.W, a.K, Kpad,
                                                      a.ws.fb_rows, a.ws.misc, a.ids, a.ids_nat, a.q, a.ws.loss_acc,
                                                      a.stats ? a.ws.counts : nullptr, sums, ids_mode_of(a.flags));
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

int launch_finish(const FwdArgs& a, bool tc_path, cudaStream_t s) {
  const double numel = (double)a.B * a.D * a.H * a.W;
  const int nrep = (a.stats && tc_path) ? tc_sums_replicas(a.K, a.D) : 1;
  int blocks = 1;
  if (a.stats) {
    const long long work = nrep > 1 ? (long long)a.K * a.D : (long long)a.K;
    blocks = (int)((work + 255) / 256);
    if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  }
  vq_finish_kernel<<<blocks, 256, 0, s>>>(a.ws.loss_acc, a.loss, 1.0 / numel, a.ws.counts, a.stats, a.K, a.D,
                                          stats_sums_offset(a.K), a.ws.sums_rep, nrep);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

int launch_ema(float* cluster_size, float* embed_avg, long long avg_sd, long long avg_sk, float* embed,
               const float* stats, int K, int D, double momentum_d, double eps_d, float count_scale, float sum_scale,
               float* scratch, cudaStream_t s) {
  // the reference forms these in Python doubles, then they are cast to fp32 when they meet the tensors
  const float momentum = (float)momentum_d;
  const float alpha = (float)(1.0 - momentum_d);              // add_(update, alpha=1 - momentum)
  const float eps = (float)eps_d;
  const float k_eps = (float)((double)K * eps_d);             // self.dict_size * self.eps
  vq_ema_cs_kernel<<<1, 1024, 0, s>>>(cluster_size, stats, K, momentum, alpha, count_scale, scratch);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  if (avg_sd == 1 && avg_sk == D) {
    const size_t tot = (size_t)K * D;
    int blocks = (int)((tot + 255) / 256);
    if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    vq_ema_embed_kd_kernel<<<blocks, 256, 0, s>>>(cluster_size, embed_avg, embed, stats + stats_sums_offset(K), K, D,
                                                  momentum, alpha, sum_scale, eps, k_eps, scratch);
  } else {
    dim3 grid((K + 31) / 32, (D + 31) / 32);
    vq_ema_embed_kernel<<<grid, 256, 0, s>>>(cluster_size, embed_avg, avg_sd, avg_sk, embed, stats + stats_sums_offset(K),
                                             K, D, momentum, alpha, sum_scale, eps, k_eps, scratch);
  }
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

#ifndef VQ_BWD_DSPLIT_DEFAULT
#define VQ_BWD_DSPLIT_DEFAULT 4
#endif
#ifndef VQ_BWD_MINB_DEFAULT
#define VQ_BWD_MINB_DEFAULT 4
#endif
#ifndef VQ_LOOKUP_MINB_DEFAULT
#define VQ_LOOKUP_MINB_DEFAULT 4
#endif

int launch_bwd(const float* g_q, const float* g_loss, const float* z, const int32_t* ids_nat, const float* snap,
               float* g_z, int B, int D, int H, int W, int K, cudaStream_t s) {
  (void)K;
  const int HW = H * W;
  const long long N = (long long)B * HW;
  const float two_over_numel = (float)(2.0 / ((double)N * D));
  auto al16 = [](const void* p) { return (((uintptr_t)p) & 15) == 0; };
  const bool vec = (HW % 4 == 0) && (D % 4 == 0) && al16(z) && al16(g_z) && (!g_q || al16(g_q)) && al16(ids_nat) &&
                   al16(snap);
  if (vec) {
    const long long nquads = N / 4;
    const long long bx = (nquads + 255) / 256;
    // split the channel loop across blockIdx.y until the grid has a few waves
    int dsplit = 1;
    while (bx * dsplit < 4LL * sm_count() && (D / (dsplit * 2)) % 4 == 0 && D / (dsplit * 2) >= 4) dsplit *= 2;
    // finer CTAs (>= 16 channels each): at config 2 the grid is only 2.3 waves of whole-pixel CTAs; measured 0.147 -> 0.129 ms
    // (0.85 -> 0.96 of the HBM peak) with four channel slices and the 64-register build (tools/knob_ab.py, profiles/r02_knob_ab.jsonl)
    const int want = tuning_knob("VQ_BWD_DSPLIT", VQ_BWD_DSPLIT_DEFAULT);
    while (dsplit < want && (D / (dsplit * 2)) % 4 == 0 && D / (dsplit * 2) >= 16) dsplit *= 2;
    dim3 grid((unsigned)bx, (unsigned)dsplit);
    if (tuning_knob("VQ_BWD_MINB", VQ_BWD_MINB_DEFAULT) >= 4)
      vq_bwd_vec_kernel<4><<<grid, 256, 0, s>>>(g_q, g_loss, z, ids_nat, snap, g_z, D, HW, nquads, dsplit, two_over_numel);
    else
      vq_bwd_vec_kernel<0><<<grid, 256, 0, s>>>(g_q, g_loss, z, ids_nat, snap, g_z, D, HW, nquads, dsplit, two_over_numel);
  } else {
    const long long bx = (N + 255) / 256;
    vq_bwd_generic_kernel<<<(unsigned)bx, 256, 0, s>>>(g_q, g_loss, z, ids_nat, snap, g_z, D, HW, N, two_over_numel);
  }
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

// Same gather, 64 a x 16 c tiles (A % 4 == 0, out 16-byte aligned): every (channel, c) row of the tile is 256 contiguous
// bytes of the output (the 32 x 32 tile wrote 128-byte pieces 2 KB apart: poor DRAM page locality for a pure write
// stream), the ids are still read in full 128-byte rows.  thread = four consecutive a of one c.
template <int TA, int TC, int MINB>
__global__ void __launch_bounds__(256, MINB)
vq_lookup_nchw_tw_kernel(const int64_t* __restrict__ ids, const float* __restrict__ E, int K, int D,
                          float* __restrict__ out, int B, int A, int C, int* __restrict__ status) {
  static_assert(TA * TC == 1024 && TA % 4 == 0, "256 threads x 4 pixels");
  __shared__ int sid[TA][TC + 1];
  const int a0 = blockIdx.x * TA, c0 = blockIdx.y * TC, b = blockIdx.z;
  {
    const int lc = threadIdx.x % TC, la = threadIdx.x / TC;           // TC consecutive c (one row of int64) x 256 / TC a per pass
#pragma unroll
    for (int r = 0; r < TA; r += 256 / TC) {
      const int a = a0 + r + la, c = c0 + lc;
      int v = -1;
      if (a < A && c < C) {
        const long long id = ids[((long long)b * A + a) * C + c];
        if (id >= 0 && id < K) v = (int)id;
        else if (status) *status = 1;
      }
      sid[r + la][lc] = v;
    }
  }
  __syncthreads();
  const int a4 = (threadIdx.x % (TA / 4)) * 4, cc = threadIdx.x / (TA / 4);
  const int a = a0 + a4, c = c0 + cc;
  if (a >= A || c >= C) return;                          // A % 4 == 0: the quad is inside or outside as a whole
  const int i0 = sid[a4][cc], i1 = sid[a4 + 1][cc], i2 = sid[a4 + 2][cc], i3 = sid[a4 + 3][cc];
  const float* e0 = E + (size_t)(i0 < 0 ? 0 : i0) * D;
  const float* e1 = E + (size_t)(i1 < 0 ? 0 : i1) * D;
  const float* e2 = E + (size_t)(i2 < 0 ? 0 : i2) * D;
  const float* e3 = E + (size_t)(i3 < 0 ? 0 : i3) * D;
  const bool v0 = i0 >= 0, v1 = i1 >= 0, v2 = i2 >= 0, v3 = i3 >= 0;       // ids outside [0, K): zeros (and `status`)
  float* o = out + (((long long)b * D) * C + c) * A + a;
  const long long dstride = (long long)C * A;
  int d = 0;
  if ((D & 3) == 0 && ((((uintptr_t)E) & 15) == 0)) {
#pragma unroll 4
    for (; d + 4 <= D; d += 4) {
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(e0 + d));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(e1 + d));
      const float4 r2 = __ldg(reinterpret_cast<const float4*>(e2 + d));
      const float4 r3 = __ldg(reinterpret_cast<const float4*>(e3 + d));
      __stcs(reinterpret_cast<float4*>(o + (d + 0) * dstride), make_float4(v0 ? r0.x : 0.f, v1 ? r1.x : 0.f, v2 ? r2.x : 0.f, v3 ? r3.x : 0.f));
      __stcs(reinterpret_cast<float4*>(o + (d + 1) * dstride), make_float4(v0 ? r0.y : 0.f, v1 ? r1.y : 0.f, v2 ? r2.y : 0.f, v3 ? r3.y : 0.f));
      __stcs(reinterpret_cast<float4*>(o + (d + 2) * dstride), make_float4(v0 ? r0.z : 0.f, v1 ? r1.z : 0.f, v2 ? r2.z : 0.f, v3 ? r3.z : 0.f));
      __stcs(reinterpret_cast<float4*>(o + (d + 3) * dstride), make_float4(v0 ? r0.w : 0.f, v1 ? r1.w : 0.f, v2 ? r2.w : 0.f, v3 ? r3.w : 0.f));
    }
  }
  for (; d < D; ++d)
    __stcs(reinterpret_cast<float4*>(o + d * dstride),
           make_float4(v0 ? __ldg(e0 + d) : 0.f, v1 ? __ldg(e1 + d) : 0.f, v2 ? __ldg(e2 + d) : 0.f, v3 ? __ldg(e3 + d) : 0.f));
}

int launch_lookup(const int64_t* ids, int64_t n, const float* embed, int K, int D, float* out, int layout, int B,
                  int A, int C, int* status, cudaStream_t s) {
  if (n == 0) return VQ_OK;
  if (layout == VQ_LAYOUT_ROWS) {
    const bool al = ((((uintptr_t)embed) | ((uintptr_t)out)) & 15) == 0;
    if ((D & 3) == 0 && !al) {
      set_error("vq_lookup: embed/out must be 16-byte aligned when D %% 4 == 0");
      return VQ_ERR_INVALID_ARG;
    }
    const long long total = (long long)n * (((D & 3) == 0) ? D / 4 : D);
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    vq_lookup_rows_kernel<<<(unsigned)blocks, 256, 0, s>>>(ids, n, embed, K, D, out, status);
  } else {
    dim3 grid((A + 31) / 32, (C + 31) / 32, B);
    static int t64 = -1;
    if (t64 < 0) { const char* e = getenv("VQ_LOOKUP_T64"); t64 = e ? atoi(e) : 1; }
    if ((A & 3) == 0 && (((uintptr_t)out) & 15) == 0 && t64 == 1) {
      dim3 g64((A + 63) / 64, (C + 15) / 16, B);
      const int minb = tuning_knob("VQ_LOOKUP_MINB", VQ_LOOKUP_MINB_DEFAULT);   // register cap -> resident CTAs per SM
      if (minb >= 5) vq_lookup_nchw_tw_kernel<64, 16, 5><<<g64, 256, 0, s>>>(ids, embed, K, D, out, B, A, C, status);
      else if (minb >= 4) vq_lookup_nchw_tw_kernel<64, 16, 4><<<g64, 256, 0, s>>>(ids, embed, K, D, out, B, A, C, status);
      else vq_lookup_nchw_tw_kernel<64, 16, 0><<<g64, 256, 0, s>>>(ids, embed, K, D, out, B, A, C, status);
    } else if ((A & 3) == 0 && (((uintptr_t)out) & 15) == 0 && t64 == 2) {
      dim3 g64((A + 127) / 128, (C + 7) / 8, B);
      vq_lookup_nchw_tw_kernel<128, 8, 0><<<g64, 256, 0, s>>>(ids, embed, K, D, out, B, A, C, status);
    } else if ((A & 3) == 0 && (((uintptr_t)out) & 15) == 0 && t64 == 3) {
      dim3 g64((A + 255) / 256, (C + 3) / 4, B);
      vq_lookup_nchw_tw_kernel<256, 4, 0><<<g64, 256, 0, s>>>(ids, embed, K, D, out, B, A, C, status);
    } else if ((A & 3) == 0 && (((uintptr_t)out) & 15) == 0)
      vq_lookup_nchw_t_kernel<true><<<grid, 256, 0, s>>>(ids, embed, K, D, out, B, A, C, status);
    else
      vq_lookup_nchw_t_kernel<false><<<grid, 256, 0, s>>>(ids, embed, K, D, out, B, A, C, status);
  }
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}


// =============================================================================================
// one-hot of an integer label map, channel-major: labels [B, HW] -> out [B, C, HW] fp32 (functions/onehot.py:5-20 builds it
// with eye(C).index_select + permute + contiguous + float: three passes over B*HW*C elements).  Pure write stream:
// thread = four consecutive pixels, one 16-byte streaming store per class; labels outside [0, C) give an all-zero column.
// =============================================================================================
template <typename LabelT>
__global__ void __launch_bounds__(256)
vq_onehot_kernel(const LabelT* __restrict__ labels, long long HW, int C, float* __restrict__ out, int cblock) {
  const long long quads = (HW + 3) >> 2;
  const long long qd = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qd >= quads) return;
  const long long b = blockIdx.z, p = qd << 2;
  const int c0 = blockIdx.y * cblock, c1 = min(C, c0 + cblock);
  long long l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) l[i] = (p + i < HW) ? (long long)labels[b * HW + p + i] : -1;
  float* o = out + (b * C + c0) * HW + p;
  if (p + 3 < HW && ((HW & 3) == 0) && (((uintptr_t)out & 15) == 0)) {
    for (int c = c0; c < c1; ++c, o += HW)
      __stcs(reinterpret_cast<float4*>(o), make_float4(l[0] == c ? 1.f : 0.f, l[1] == c ? 1.f : 0.f, l[2] == c ? 1.f : 0.f,
                                                         l[3] == c ? 1.f : 0.f));
  } else {
    for (int c = c0; c < c1; ++c, o += HW)
      for (int i = 0; i < 4; ++i)
        if (p + i < HW) o[i] = l[i] == c ? 1.f : 0.f;
  }
}

int launch_onehot(const void* labels, int label_bytes, int64_t B, int64_t HW, int C, float* out, cudaStream_t s) {
  if (B == 0 || HW == 0) return VQ_OK;
  const long long quads = (HW + 3) / 4;
  const int cblock = C < 32 ? C : 32;                      // classes per CTA: keeps the label quad in registers for 32 stores
  dim3 grid((unsigned)((quads + 255) / 256), (unsigned)((C + cblock - 1) / cblock), (unsigned)B);
  VQ_REQUIRE(grid.y <= 65535 && B <= 65535, VQ_ERR_UNSUPPORTED, "vq_onehot: too many classes / images");
  if (label_bytes == 8) vq_onehot_kernel<int64_t><<<grid, 256, 0, s>>>((const int64_t*)labels, HW, C, out, cblock);
  else vq_onehot_kernel<int32_t><<<grid, 256, 0, s>>>((const int32_t*)labels, HW, C, out, cblock);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

}  // namespace vqb200

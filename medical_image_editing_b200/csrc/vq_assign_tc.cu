// vq_assign_tc.cu -- tcgen05 / TMEM / TMA nearest-code search with exact fp32 re-rank (sm_100a).
//
// Two kernels, one persistent CTA per SM (640 threads, 227 KB shared memory, all 512 TMEM columns), warp-specialised:
//
// vq_assign_tc_kernel  (codebook resident in shared memory: K*D*4 <= ~140 KB)
//   warp 0       TMA producer: the norm-sorted codebook once, then one 128-pixel z tile per stage (2 stages), loaded
//                straight from NCHW (pixel-contiguous => MN-major A operand, no flatten copy)
//   warp 1       MMA issuer: tcgen05.mma kind::tf32, M=128 pixels x N<=256 codes x K=8 per instruction, fp32 accumulators
//                in TMEM (2 stages x 256 columns); one extra K-step multiplies a block of ones with a 3-way tf32 split of
//                the augmentation -|e|^2 / (2 (1 + 2^-10)), so the accumulator is the (centred) score
//   warps 2,3    |z|^2 of the tile's pixels (bound of the tf32 error, last term of the exact score)
//   warps 4..11  scan, warp = (TMEM lane quadrant, column group), thread = pixel: tcgen05.ld 32 columns at a time,
//                running max (FMNMX3) and a sign-bit candidate mask against (max - bound); publishes
//                {bounds, <= 2 candidate chunks + masks} per pixel to shared memory
//   warps 12..19 output, lane = (pixel, channel-quad parity): z of the tile into registers (the stage goes back to the
//                producer at once), merge of the column groups, exact fp32 re-rank of pixels with several candidates by
//                the pixel's own two lanes, then ids, q, (z-q)^2, histogram and EMA sums
// vq_assign_tcs_kernel (codebook streamed: D >= 128 at K = 512, K = 4096)
//   one ring of (z chunk + codebook slice) stages, sixteen epilogue warps in two teams that each scan, merge, re-rank
//   and write every other tile; z and the exact code rows of the output phase come from L2.
//
// Exactness: kind::tf32 TRUNCATES the fp32 operands (tools/trunc_check.py), so every product shrinks by a factor in
//   (1 - 2^-9, 1]; with the factor (1 + 2^-10) folded into the augmentation an approximate score is off by at most
//   delta = 2^-10 |z| |e| (+ accumulation slop).  Every code whose approximate score is within 2 delta of the approximate
//   maximum is re-scored in exact fp32 (the reference's op order), so the winner is the fp32 winner.  Rows with more
//   candidates than the kernel keeps, non-finite rows, or rows where a norm-outlier ("exploded") code could still win are
//   appended to a list that a CUDA-core kernel then searches exhaustively.  Nothing is probabilistic.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "vq_common.cuh"

namespace vqb200 {

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
constexpr int TC_TILE = 128;        // pixels per tile (UMMA M)
constexpr int TC_MAXBN = 256;       // codes per accumulator stage (UMMA N)
constexpr int TC_DCH = 32;          // channels per shared-memory chunk (128-byte swizzle rows)
#ifndef VQ_TC_NCG
#define VQ_TC_NCG 2
#endif
#ifndef VQ_TC_OPX_SHIFT
#define VQ_TC_OPX_SHIFT 4
#endif
constexpr int TC_NCG = VQ_TC_NCG;   // column groups: scan warps per TMEM lane quadrant
constexpr int TC_SCAN_WARPS = 4 * TC_NCG;
constexpr int TC_OPX_SHIFT = VQ_TC_OPX_SHIFT;
constexpr int TC_OPX = 1 << TC_OPX_SHIFT;        // pixels per output warp
constexpr int TC_OUT_WARPS = 128 / TC_OPX;       // TC_OPX pixels of the tile each
constexpr int TC_OCS = 32 / TC_OPX;              // lanes per pixel = channel splits (4): lane = px + TC_OPX * hf
static_assert((1 << TC_OPX_SHIFT) == TC_OPX && TC_OPX * TC_OCS == 32 && TC_OCS == 2, "output-warp lane mapping: lane = (pixel, quad parity)");
constexpr int TC_AUX_WARPS = 4;     // TMA producer, MMA issuer, two |z|^2 workers
constexpr int TC_THREADS = 32 * (TC_AUX_WARPS + TC_SCAN_WARPS + TC_OUT_WARPS);
constexpr int TC_MAXCAND = 8;       // candidates re-scored exactly per pixel (more: exhaustive fallback)
constexpr int TC_SMEM_LIMIT = 227 * 1024;
constexpr int TC_SORT_MAX = 4096;     // codes (16-bit sorted positions, 7-bit chunk indices in the epilogue)
constexpr int TC_MAX_REP = 32;        // replicas of the per-code sums (spreads the L2 reduction traffic)

struct TcGeom {
  int BN, nb, nD, nst;
  size_t off_emain, off_eaug, off_aaug, off_z, off_pub, off_wl, off_zn, off_hist, off_perm, off_ctab, off_bar, total;
  bool ok;
};

static TcGeom tc_geometry(int D, int K) {
  TcGeom g{};
  g.ok = false;
  g.BN = 32;                                   // power of two (the epilogue shifts by it), <= 256
  while (g.BN < K && g.BN < TC_MAXBN) g.BN <<= 1;
  g.nb = (K + g.BN - 1) / g.BN;
  g.nD = (D + TC_DCH - 1) / TC_DCH;
  const size_t ktot = (size_t)g.nb * g.BN;
  const size_t emain = ktot * g.nD * 128;
  const size_t eaug = ktot * 32;
  const size_t zstage = (size_t)g.nD * TC_TILE * 128;
  size_t off = 0;
  g.off_emain = off; off += emain;
  g.off_eaug = off;  off += align_up(eaug, 1024);
  g.off_aaug = off;  off += 4096;
  g.off_z = off;
  const size_t sz_pub = (size_t)2 * TC_NCG * TC_TILE * 16;     // [tile parity][column group][pixel] x 16 B
  const size_t sz_wl = 16, sz_zn = 2 * TC_TILE * 4;
  const size_t sz_hist = align_up((size_t)K * 4, 16), sz_perm = align_up(ktot * 2, 16);
  const size_t sz_ctab = align_up((size_t)g.nb * (g.BN / 32) * 8, 16);
  const size_t tail = sz_pub + sz_wl + sz_zn + sz_hist + sz_perm + sz_ctab + 256;
  for (int nst = 2; nst >= 1; --nst) {
    if (off + nst * zstage + tail + 1024 <= (size_t)TC_SMEM_LIMIT) { g.nst = nst; g.ok = true; break; }
  }
  if (!g.ok) return g;
  off += g.nst * zstage;
  g.off_pub = off;  off += sz_pub;
  g.off_wl = off;   off += sz_wl;
  g.off_zn = off;   off += sz_zn;
  g.off_hist = off; off += sz_hist;
  g.off_perm = off; off += sz_perm;
  g.off_ctab = off; off += sz_ctab;
  g.off_bar = off;  off += 256;
  g.total = off + 1024;   // slack for manual 1024-byte alignment of the dynamic smem base
  return g;
}

int tc_sums_replicas(int K, int D) {
  // enough replicas that concurrent CTAs rarely reduce into the same L2 line, capped at 4 MB in total (the replicas
  // are zeroed by vq_prep_kernel and folded by vq_finish_kernel on every call)
  size_t r = ((size_t)4 << 20) / ((size_t)K * D * 4);
  if (r > (size_t)TC_MAX_REP) r = TC_MAX_REP;
  if (r < 1) r = 1;
  return (int)r;
}

static bool tcs_supported(int D, int K);

bool tc_path_supported(int B, int D, int H, int W, int K) {
  const long long HW = (long long)H * W;
  if (B <= 0 || HW <= 0) return false;
  if (HW % TC_TILE != 0) return false;          // tiles never straddle images; TMA strides need HW % 4 == 0
  if (D % 4 != 0 || D < 4) return false;        // 16-byte rows for TMA / float4 gathers
  if (K < 1) return false;
  const TcGeom g = tc_geometry(D, K);
  if (g.ok && g.nb * g.BN <= TC_SORT_MAX) return true;      // codebook resident in shared memory
  return tcs_supported(D, K);                                // codebook streamed through shared memory
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint expires),
  // so a waiting warp does not burn issue slots of its SM sub-partition
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
}
// polling wait with nanosleep back-off: used by the single-thread producer / MMA roles so that their spinning
// does not steal issue slots from the epilogue warps sharing the same SM sub-partition
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    asm volatile("nanosleep.u32 %0;" ::"r"(ns));
  }
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_v2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)(layout & 7) << 61;      // 0 = no swizzle, 1 = 128 B swizzle / 32 B atom, 2 = 128 B swizzle
  return d;
}
// instruction descriptor: tf32 x tf32 -> f32, A MN-major (pixels contiguous), B K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                  // c_format  = F32
  d |= 2u << 7;                  // a_format  = TF32
  d |= 2u << 10;                 // b_format  = TF32
  d |= 1u << 15;                 // a_major   = MN
  d |= 0u << 16;                 // b_major   = K
  d |= (uint32_t)(n >> 3) << 17; // N
  d |= (uint32_t)(128 >> 4) << 24;  // M
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- q leaves through the TMA: a [16 pixel x 32 channel] box staged in shared memory per warp
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster share one M = 256 MMA; loads of both CTAs signal the
// leader's (cluster rank 0) barrier, tcgen05.commit arrives on the same barrier slot of both CTAs
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_idx() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t leader_addr(uint32_t local_addr) {      // same offset in the shared memory of cluster rank 0
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local_addr));
  return r;
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t local_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(local_bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep_cluster(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    asm volatile("nanosleep.u32 %0;" ::"r"(ns));
  }
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_pair(int n) {           // as make_idesc, M = 256 over the two CTAs
  return (make_idesc(n) & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {       // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// prep 2 (single CTA): sort the codebook by norm, per-chunk error-bound tables, augmentation image
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Codes are laid out in ascending-norm order so that the 32 codes of one scan chunk have similar norms: the
// approximate-score error of code k is bounded by (c1 |z| |e_k| + c2 |e_k| (|z| + |e_k|)) / 2, and the scan uses
// one bound per chunk (largest norm in the chunk).  Codes whose norm exceeds 64x the lower edge of the median
// exponent bin ("big": the exploded dead codes of EMA training) are excluded from the
// approximate search (augmentation = -1e30) and handled by a rigorous per-row test in the main kernel.
//   rmax[c]  = largest live norm in sorted chunk c (uint32 bits of a non-negative float; zeroed by vq_prep_kernel);
//              the main kernel turns it into bound_c(|z|) = |z| * A_c + B_c  (accumulator units)
//   meta[0] = r_cap, meta[1] = smallest norm of an excluded code as uint32 bits (0x7F800000 = none; set by prep)
//
// Rank by counting: every thread owns one sorted-order key (norm bits, original index) and counts the smaller keys
// in shared memory -- O(K^2 / threads) compares, no sorting network, any number of CTAs.
constexpr int TC_PREP2_THREADS = 256;
__global__ void __launch_bounds__(TC_PREP2_THREADS)
vq_tc_prep2_kernel(const float* __restrict__ E, const float* __restrict__ e2, int K, int D, int BN, int nb,
                   float* __restrict__ es, float* __restrict__ eaug_img, int* __restrict__ perm,
                   uint32_t* __restrict__ rmax, uint32_t* __restrict__ meta) {
  extern __shared__ __align__(16) uint32_t keys[];          // [K] norm bits (NaN / negative -> +inf)
  __shared__ int hist[256];
  __shared__ float s_rcap;
  const int tid = threadIdx.x;
  const int ktot = nb * BN;
  hist[tid] = 0;                              // TC_PREP2_THREADS == 256
  __syncthreads();
  for (int k = tid; k < K; k += blockDim.x) {
    const float r = sqrtf(e2[k]);
    uint32_t rb = __float_as_uint(r);
    if (!(r >= 0.f)) rb = 0x7F800000u;
    keys[k] = rb;
    atomicAdd(&hist[(rb >> 23) & 0xFF], 1);
  }
  __syncthreads();
  if (tid < 32) {                             // median exponent bin: warp-parallel prefix over the 256 bins
    int cum = 0, emed = 255;
    bool found = false;
    for (int base = 0; base < 256 && !found; base += 32) {
      int v = hist[base + tid];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (tid >= o) v += t;
      }
      const uint32_t hit = __ballot_sync(0xffffffffu, 2 * (cum + v) >= K);
      if (hit) { emed = base + __ffs(hit) - 1; found = true; }
      cum += __shfl_sync(0xffffffffu, v, 31);
    }
    int ecap = emed + 6;                      // 64 x the lower edge of the median exponent bin
    if (ecap > 254) ecap = 254;
    if (tid == 0) s_rcap = __uint_as_float((uint32_t)ecap << 23);
  }
  __syncthreads();
  const float rcap = s_rcap;
  const float slop = 1.f + (float)D * 1.2e-7f + 1e-5f;          // |e| was computed in fp32 from a rounded |e|^2
  const int i = blockIdx.x * blockDim.x + tid;                 // code (i < K) or padding slot (K <= i < ktot)
  if (i == 0) meta[0] = __float_as_uint(rcap);
  if (i < ktot) {
    int pos = i;                              // padding keeps its slot: K .. ktot-1
    float a0 = -1e30f, a1 = 0.f, a2 = 0.f;    // padding / excluded codes never win
    const bool real = i < K;
    if (real) {
      const uint32_t mine = keys[i];
      int rank = 0;
      int j = 0;
      const uint4* k4 = reinterpret_cast<const uint4*>(keys);
      for (; j + 4 <= K; j += 4) {            // broadcast reads, four keys per 16-byte load
        const uint4 kk = k4[j >> 2];
        rank += (kk.x < mine || (kk.x == mine && j < i)) ? 1 : 0;
        rank += (kk.y < mine || (kk.y == mine && j + 1 < i)) ? 1 : 0;
        rank += (kk.z < mine || (kk.z == mine && j + 2 < i)) ? 1 : 0;
        rank += (kk.w < mine || (kk.w == mine && j + 3 < i)) ? 1 : 0;
      }
      for (; j < K; ++j) {
        const uint32_t kj = keys[j];
        rank += (kj < mine || (kj == mine && j < i)) ? 1 : 0;
      }
      pos = rank;
      const float r = __uint_as_float(mine);
      if (r <= rcap) {
        // tf32 operands are TRUNCATED (tools/trunc_check.py), so every product shrinks by a factor in (1 - 2^-9, 1]:
        // (1 + 2^-10) * dot~ is within 2^-10 sum|z_d e_d| of the exact dot -- half of the uncentred bound.  Comparing
        // (1 + 2^-10) dot~ - |e|^2/2 is comparing dot~ - |e|^2 / (2 (1 + 2^-10)): fold the factor into the augmentation.
        const float x = -0.5f * e2[i] * (1.0f / (1.0f + 0.0009765625f));
        a0 = tf32_trunc(x);
        const float r1 = x - a0;
        a1 = tf32_trunc(r1);
        a2 = tf32_trunc(r1 - a1);
        atomicMax(&rmax[pos >> 5], __float_as_uint(r * slop));
      } else {
        atomicMin(&meta[1], __float_as_uint(r / slop));
      }
    }
    perm[pos] = real ? i : 0;
    const int blk = pos / BN, rr = pos % BN, grp = rr >> 3, row = rr & 7;
    float* base = eaug_img + (size_t)blk * BN * 8 + grp * 64 + row * 4;
    base[0] = a0; base[1] = a1; base[2] = a2; base[3] = real ? e2[i] : 0.f;   // [3]: exact |e|^2 (times the zero row of the ones block)
    base[32] = 0.f; base[33] = 0.f; base[34] = 0.f; base[35] = 0.f;
    // sorted copy of the codebook row (TMA source)
    const int dq = D >> 2;
    float4* dst = reinterpret_cast<float4*>(es + (size_t)pos * D);
    const float4* src = reinterpret_cast<const float4*>(E + (size_t)(real ? i : 0) * D);
    int j = 0;
    for (; j + 8 <= dq; j += 8) {                 // eight 16-byte loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = real ? __ldg(src + j + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[j + u] = v[u];
    }
    for (; j < dq; ++j) dst[j] = real ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
struct TcParams {
  const float* z; const float* E; const float* e2; const float* eaug_img; const uint32_t* meta;
  const int* perm; const uint32_t* rmax;
  int B, D, H, W, HW, K;
  int BN, nb, nD, nst;
  int bn_shift;           // BN == 1 << bn_shift
  int w_shift;            // W == 1 << w_shift, or -1
  int tiles_per_img; int ntiles;
  uint32_t off_emain, off_eaug, off_aaug, off_z, off_pub, off_wl, off_zn, off_hist, off_perm, off_ctab, off_bar;
  int64_t* ids; int32_t* ids_nat; float* q; double* loss_acc; int* counts;
  float* sums;            // replica 0 of the per-code sums (inside the packed statistics buffer), or null
  float* sums_rep;        // replicas 1..nrep-1 (workspace), [nrep-1][K*D]
  int nrep;
  int* fb_count; int* fb_rows;
  int ids_mode;    // layout / base of the int64 code map (store_id)
  float* dbg;      // optional [N][nb*BN] dump of the raw accumulators
};

// byte offset of z(p, d) inside a z stage: chunk (d/32) -> group (p/32) -> row (d%32) of 128 B -> 32-byte atom
// ((p%32)/8) XOR (row%4).  tf32 MN-major operands only exist in the "128-byte swizzle, 32-byte atom" layout
// (UMMA layout type 1 / CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); the plain 128-byte swizzle silently yields zeros.
__device__ __forceinline__ uint32_t zs_off(int p, int d) {
  const int row = d & 31;
  return (uint32_t)((d >> 5) * (TC_TILE * 128) + (p >> 5) * 4096 + row * 128 +
                    ((((p & 31) >> 2) ^ ((row & 3) << 1)) << 4) + ((p & 3) << 2));
}

// 1.0f if x < 0 (any negative normal number), else 0.0f -- one saturating multiply on the fp32 pipe.  .ftz flushes a
// denormal x to zero, so the result is exactly 0 or 1 (a difference that small is far inside the bound's slack).
__device__ __forceinline__ float sign01(float x) {
  float r;
  asm("mul.ftz.sat.f32 %0, %1, 0fFE800000;" : "=f"(r) : "f"(x));      // x * -2^126, clamped to [0, 1]
  return r;
}
__device__ __forceinline__ uint32_t f32_orderable(float x) {   // monotone map float -> uint32
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// round a float UP to a value whose low 16 bits are zero (conservative 16-bit copy of an upper bound)
__device__ __forceinline__ uint32_t f32_up16(float x) {
  const uint32_t u = __float_as_uint(x);
  if (u & 0x80000000u) return u & 0xFFFF0000u;              // negative: shrinking the magnitude moves up
  return (u + 0xFFFFu) & 0xFFFF0000u;                        // positive: bump the magnitude (inf stays inf)
}

// Optional role timing (build with -DVQ_TC_TIMING; read with vq_debug_tc_timing): clock64 deltas summed over tiles by
// lane 0 of every scan / output warp: [cta][warp 0..15][slot], slots: scan warps 0 wait |z|^2, 1 wait tmem, 2 scan work,
// 3 wait pub slot, 4 publish ; output warps 0 wait z/|z|^2, 1 wait scan results, 2 merge, 3 pair list + re-rank,
// 4 outputs ; slot 6 tiles, slot 7 pairs
#ifdef VQ_TC_TIMING
__device__ long long g_tc_timing[148 * 16 * 8];
#define TC_TIMING_DECL long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tlast = clock64();
#define TC_TICK(slot)                                                   \
  do {                                                                  \
    const long long _now = clock64();                                   \
    tacc[slot] += _now - tlast;                                         \
    tlast = _now;                                                       \
  } while (0)
#define TC_TIMING_STORE(widx, ntiles_)                                                                   \
  do {                                                                                                  \
    if (lane == 0 && blockIdx.x < 148) {                                                                \
      tacc[6] = (ntiles_);                                                                              \
      for (int _i = 0; _i < 8; ++_i) g_tc_timing[((size_t)blockIdx.x * 16 + (widx)) * 8 + _i] = tacc[_i]; \
    }                                                                                                   \
  } while (0)
#else
#define TC_TIMING_DECL
#define TC_TICK(slot) do { } while (0)
#define TC_TIMING_STORE(widx, ntiles_) do { } while (0)
#endif

// DT: compile-time emb_dim (0 = run-time P.D); the specialisations fully unroll the per-channel loops
template <bool DBG, bool STATS, int DT>
__global__ void __launch_bounds__(TC_THREADS, 1)
vq_assign_tc_kernel(const __grid_constant__ CUtensorMap zmap, const __grid_constant__ CUtensorMap emap, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint64_t* bars = (uint64_t*)(smem + P.off_bar);
  const uint32_t bar0 = sbase + P.off_bar;
  // barrier slots: 0 e_full | 1,2 z_full | 3,4 z_empty | 5,6 tmem_full | 7,8 tmem_empty | 9,10 zn_full |
  //                11,12 pub_full | 13,14 pub_empty ; slot 15: tmem base
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + 15);
  int* hist = (int*)(smem + P.off_hist);
  uint16_t* perm_s = (uint16_t*)(smem + P.off_perm);

  const int Dc = DT ? DT : P.D;                              // emb_dim
  const int nD = DT ? (DT + TC_DCH - 1) / TC_DCH : P.nD;     // 32-channel chunks (zero-padded by TMA)
  const uint32_t zstage_bytes = (uint32_t)nD * TC_TILE * 128;
  const int ktot = P.nb * P.BN;

  if (threadIdx.x == 32) {
    mbar_init(BAR(0), 1);
    mbar_init(BAR(1), 1); mbar_init(BAR(2), 1);
    mbar_init(BAR(3), TC_OUT_WARPS + 3); mbar_init(BAR(4), TC_OUT_WARPS + 3);   // output warps, |z|^2 warps, MMA commit
    mbar_init(BAR(5), 1); mbar_init(BAR(6), 1);
    mbar_init(BAR(7), TC_SCAN_WARPS); mbar_init(BAR(8), TC_SCAN_WARPS);
    mbar_init(BAR(9), 2); mbar_init(BAR(10), 2);
    mbar_init(BAR(11), TC_SCAN_WARPS); mbar_init(BAR(12), TC_SCAN_WARPS);
    mbar_init(BAR(13), TC_OUT_WARPS); mbar_init(BAR(14), TC_OUT_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= TC_AUX_WARPS) {
    // ones block of the augmentation K-step: 4 groups x 8 rows x 128 B; rows 0..2 = 1, rows 3..7 = 0
    const int t = threadIdx.x - 32 * TC_AUX_WARPS;
    constexpr int NT = 32 * (TC_SCAN_WARPS + TC_OUT_WARPS);
    float4* a = (float4*)(smem + P.off_aaug);
    for (int i = t; i < 256; i += NT) {                   // 256 float4 = 4 KB
      const int row = (i >> 3) & 7;
      const float v = row < 3 ? 1.f : 0.f;
      a[i] = make_float4(v, v, v, v);
    }
    for (int k = t; k < P.K; k += NT) hist[k] = 0;
    for (int k = t; k < ktot; k += NT) perm_s[k] = (uint16_t)P.perm[k];
    {   // per-chunk error-bound coefficients from the chunk's largest norm (already includes the rounding slop)
      const float c1 = 0.001953125f * 1.03f;                    // 2^-9 (score units): truncated tf32 operands, centred (prep2)
      const float c2 = (float)(Dc + 16) * 4.76837158e-7f;       // (D+16) 2^-21: fp32 accumulation in the tensor core
      for (int c = t; c < P.nb * (P.BN >> 5); c += NT) {
        const float rm = __uint_as_float(P.rmax[c]);
        ((float*)(smem + P.off_ctab))[2 * c] = 0.5f * (c1 + c2) * rm;                 // A_c
        ((float*)(smem + P.off_ctab))[2 * c + 1] = 0.5f * c2 * rm * rm + 1e-30f;      // B_c
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool two_stages = P.nst == 2;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      const uint32_t ebytes = (uint32_t)P.nb * nD * P.BN * 128 + (uint32_t)P.nb * P.BN * 32;
      mbar_expect_tx(BAR(0), ebytes);
      for (int blk = 0; blk < P.nb; ++blk)
        for (int c = 0; c < nD; ++c)
          tma_load_2d(sbase + P.off_emain + (uint32_t)(blk * nD + c) * P.BN * 128, &emap, BAR(0), c * TC_DCH, blk * P.BN);
      bulk_load_1d(sbase + P.off_eaug, P.eaug_img, (uint32_t)P.nb * P.BN * 32, BAR(0));
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int s = two_stages ? (it & 1) : 0, ph = two_stages ? ((it >> 1) & 1) : (it & 1);
        mbar_wait_sleep(BAR(3 + s), ph ^ 1, 128);
        mbar_expect_tx(BAR(1 + s), zstage_bytes);
        const int b = tile / P.tiles_per_img, pt = tile % P.tiles_per_img;
        for (int c = 0; c < nD; ++c)
          for (int grp = 0; grp < 4; ++grp)    // one 32-pixel x 32-channel box per swizzle atom column
            tma_load_3d(sbase + P.off_z + s * zstage_bytes + c * (TC_TILE * 128) + grp * 4096, &zmap, BAR(1 + s),
                        pt * TC_TILE + grp * 32, c * TC_DCH, b);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(P.BN);
      mbar_wait(BAR(0), 0);
      int g = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = two_stages ? (it & 1) : 0, ph = two_stages ? ((it >> 1) & 1) : (it & 1);
        mbar_wait_sleep(BAR(1 + s), ph, 32);
        tc_fence_after();
        const uint32_t zaddr = sbase + P.off_z + s * zstage_bytes;
        for (int blk = 0; blk < P.nb; ++blk, ++g) {
          const int a = g & 1, aph = (g >> 1) & 1;
          mbar_wait_sleep(BAR(7 + a), aph ^ 1, 32);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)a * TC_MAXBN;
          uint32_t acc = 0;
          for (int c = 0; c < nD; ++c) {
            const int ksteps = min(4, (Dc - c * TC_DCH + 7) >> 3);
            const uint32_t eaddr = sbase + P.off_emain + (uint32_t)(blk * nD + c) * P.BN * 128;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t ad = make_desc(zaddr + c * (TC_TILE * 128) + ks * 1024, 4096, 512, 1);
              const uint64_t bd = make_desc(eaddr + ks * 32, 16, 1024, 2);
              umma_tf32(d_tmem, ad, bd, idesc, acc);
              acc = 1;
            }
          }
          {   // augmentation K-step: ones x (-|e|^2/2 split in three tf32 pieces)
            const uint64_t ad = make_desc(sbase + P.off_aaug, 1024, 512, 1);
            const uint64_t bd = make_desc(sbase + P.off_eaug + (uint32_t)blk * P.BN * 32, 128, 256, 0);
            umma_tf32(d_tmem, ad, bd, idesc, acc);
          }
          umma_commit(BAR(5 + a));
        }
        umma_commit(BAR(3 + s));               // the tensor core is done reading this z stage
      }
    }
  } else if (warp < TC_AUX_WARPS) {
    // ===================================== |z|^2 workers ====================================
    // two warps, two ADJACENT pixels per lane (one 8-byte shared-memory load, one packed fma per channel); per
    // pixel this is the same ascending-d fma chain as in the CUDA-core kernels
    const int pA = (warp - 2) * 64 + 2 * lane;            // pixels pA, pA + 1 (same 16-byte atom)
    const uint32_t zrow0 = sbase + P.off_z + (uint32_t)(pA >> 5) * 4096 + ((pA & 3) << 2);
    const uint32_t zn_s = sbase + P.off_zn;
    uint32_t zx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zx[i] = (uint32_t)i * 128 + (uint32_t)((((pA & 31) >> 2) ^ (i << 1)) << 4);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = two_stages ? (it & 1) : 0, ph = two_stages ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(BAR(1 + s), ph);
      uint32_t zc = zrow0 + s * zstage_bytes;
      float2 zz = make_float2(0.f, 0.f), zzB = make_float2(0.f, 0.f);    // |z|^2 = A + B (even / odd channel quads)
#pragma unroll
      for (int c = 0; c < nD; ++c) {                      // channels beyond D are zero-filled by TMA: no guards
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 v = lds_v2(zc + jj * 512 + zx[i]);
            if ((jj & 1) == 0) zz = __ffma2_rn(v, v, zz);
            else zzB = __ffma2_rn(v, v, zzB);
          }
        }
        zc += 16384;
      }
      zz = __fadd2_rn(zz, zzB);
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(zn_s + (uint32_t)(s * TC_TILE + pA) * 4), "f"(zz.x), "f"(zz.y) : "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(BAR(9 + s)); mbar_arrive(BAR(3 + s)); }
    }
  } else if (warp < TC_AUX_WARPS + TC_SCAN_WARPS) {
    // ===================================== scan warps =======================================
    // warp = (TMEM lane quadrant, column group): thread = (pixel, 1/TC_NCG of the 32-code chunks).  Running max and
    // sign-bit candidate masks over the accumulators; the result (bounds + <= 2 candidate chunks) is published in
    // shared memory for the output warps.  No exact arithmetic, no synchronisation with sibling warps.
    const int quad = warp & 3, cg = (warp - TC_AUX_WARPS) >> 2;
    const int p = quad * 32 + lane;                       // pixel within the tile == TMEM lane
    const uint32_t ctab_s = sbase + P.off_ctab;
    const uint32_t zn_s = sbase + P.off_zn + (uint32_t)p * 4;
    const uint32_t pub_s = sbase + P.off_pub + (uint32_t)(cg * TC_TILE + p) * 16;
    const int nchunks = P.BN >> 5;
    int g = 0;
    int tb = (int)blockIdx.x / P.tiles_per_img, tpt = (int)blockIdx.x % P.tiles_per_img;   // only for the debug dump
    TC_TIMING_DECL
    for (int it = 0; it < my_tiles; ++it) {
      const int s = two_stages ? (it & 1) : 0, ph = two_stages ? ((it >> 1) & 1) : (it & 1);
      TC_TICK(4);
      mbar_wait(BAR(9 + s), ph);                          // |z|^2 of this tile
      TC_TICK(0);
      float z2;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(z2) : "r"(zn_s + (uint32_t)s * (TC_TILE * 4)));
      const float zn = sqrtf(z2) * 1.00001f;
      // Running state (accumulator units, a_k = z.e_k - |e_k|^2/2):
      //   L     lower bound on the best exact a_k among the columns this thread has seen = max_c (chunkmax_c - delta_c)
      //   Urec  upper bound on the exact a_k of every recorded candidate
      // A column of chunk c is a candidate iff approx + delta_c >= L, i.e. approx >= L - delta_c.
      float L = -INFINITY, Urec = -INFINITY;
      int cnt = 0;
      uint32_t rcA = 0, rcB = 0, rm0 = 0, rm1 = 0;        // records: global chunk index, candidate mask
      for (int blk = 0; blk < P.nb; ++blk, ++g) {
        const int a = g & 1, aph = (g >> 1) & 1;
        TC_TICK(2);
        mbar_wait(BAR(5 + a), aph);
        TC_TICK(1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)a * TC_MAXBN;
        for (int c = cg; c < nchunks; c += TC_NCG) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          const int gc = blk * nchunks + c;               // global chunk index (sorted codes gc*32 .. gc*32+31)
          float cA, cB;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cA), "=f"(cB) : "r"(ctab_s + (uint32_t)gc * 8));
          const float delta = __fmaf_rn(zn, cA, cB);
          tmem_ld_wait();
          if (DBG) {
            float* o = P.dbg + ((size_t)tb * P.HW + tpt * TC_TILE + p) * ktot + gc * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = v[j];
          }
          float m4[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {                   // four independent max chains (8 columns each)
            float m = fmaxf(fmaxf(v[8 * h], v[8 * h + 1]), v[8 * h + 2]);
            m = fmaxf(fmaxf(m, v[8 * h + 3]), v[8 * h + 4]);
            m = fmaxf(fmaxf(m, v[8 * h + 5]), v[8 * h + 6]);
            m4[h] = fmaxf(m, v[8 * h + 7]);
          }
          const float cm = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          L = fmaxf(L, cm - delta);
          if (Urec < L) { cnt = 0; Urec = -INFINITY; }    // nothing recorded so far can still win
          const float T = L - delta;
          const float2 nT2 = make_float2(-T, -T);
          // "below threshold" bits of the 32 columns.  The integer pipe (funnel shifts) and the fp32 pipe (saturating
          // multiply -> exact 0/1, then acc = 2*acc + bit) each build half of them, so the two pipes work in parallel.
          uint32_t n4[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (h & 1) {                                  // columns 8h..8h+7 on the fp32 pipe
              float facc = 0.f;
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const float2 d2 = __fadd2_rn(make_float2(v[8 * h + j], v[8 * h + j + 1]), nT2);
                facc = __fmaf_rn(facc, 2.f, sign01(d2.x));
                facc = __fmaf_rn(facc, 2.f, sign01(d2.y));
              }
              n4[h] = (uint32_t)facc;                     // exact: 8 bits
            } else {                                      // columns 8h..8h+7 on the integer pipe
              uint32_t nm = 0;
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const float2 d2 = __fadd2_rn(make_float2(v[8 * h + j], v[8 * h + j + 1]), nT2);
                nm = __funnelshift_l(__float_as_uint(d2.x), nm, 1);
                nm = __funnelshift_l(__float_as_uint(d2.y), nm, 1);
              }
              n4[h] = nm;
            }
          }
          const uint32_t nmall = (n4[0] << 24) | (n4[1] << 16) | (n4[2] << 8) | n4[3];
          const uint32_t cand = ~nmall;                   // bit (31-j) set <=> column j is within the bound
          // branch-free record update
          const bool has = cand != 0u;
          const bool s0 = has && cnt == 0, s1 = has && cnt == 1;
          rcA = s0 ? (uint32_t)gc : rcA; rm0 = s0 ? cand : rm0;
          rcB = s1 ? (uint32_t)gc : rcB; rm1 = s1 ? cand : rm1;
          cnt += has ? 1 : 0;
          Urec = has ? fmaxf(Urec, cm + delta) : Urec;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(7 + a));
      }
      if (DBG && cg == 0) {   // second debug area (after the accumulators): what the warps see in shared memory
        float* o2 = P.dbg + (size_t)P.B * P.HW * ktot + ((size_t)tb * P.HW + tpt * TC_TILE + p) * 8;
        const uint8_t* zs = smem + P.off_z + s * zstage_bytes;
        const uint8_t* eb0 = smem + P.off_emain + (size_t)p * 128;                       // code p of block 0, chunk 0
        o2[0] = z2;
        o2[1] = *(const float*)(zs + zs_off(p, 0));
        o2[2] = *(const float*)(zs + zs_off(p, 1));
        o2[3] = *(const float*)(eb0 + ((0 ^ (p & 7)) << 4));                             // E[p][0]
        o2[4] = *(const float*)(eb0 + ((1 ^ (p & 7)) << 4) + 4);                         // E[p][5]
        o2[5] = *(const float*)(smem + P.off_eaug + (p >> 3) * 256 + (p & 7) * 16);      // aug[p][0]
        o2[6] = *(const float*)(smem + P.off_aaug + p * 4);
        o2[7] = __uint_as_float(tmem_base);
      }
      if (DBG) { tpt += (int)gridDim.x; while (tpt >= P.tiles_per_img) { tpt -= P.tiles_per_img; ++tb; } }
      // ---- publish: {L, U16 | chunkA<<8 | chunkB<<1 | overflow, maskA, maskB} -------------------------------
      if (cnt < 2) rm1 = 0;
      if (cnt < 1) rm0 = 0;
      const uint32_t w1 = f32_up16(Urec) | (rcA << 8) | (rcB << 1) | (cnt > 2 ? 1u : 0u);   // chunk indices < 128
      const int par = it & 1, pph = (it >> 1) & 1;
      TC_TICK(2);
      mbar_wait(BAR(13 + par), pph ^ 1);                  // the output warps are done with this slot (tile it-2)
      TC_TICK(3);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(pub_s + (uint32_t)par * (TC_NCG * TC_TILE * 16)),
                   "r"(__float_as_uint(L)), "r"(w1), "r"(rm0), "r"(rm1) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(11 + par));
    }
    TC_TIMING_STORE(warp - TC_AUX_WARPS, my_tiles);
  } else {
    // ===================================== output warps =====================================
    // warp ow owns TC_OPX pixels of the tile; lane = (pixel px, channel split hf): all decisions for a
    // pixel are taken inside one warp.  Per tile: merge the scan warps' bounds -> single candidate or a list of
    // (pixel, code) pairs -> exact fp32 re-rank of the pairs (one lane per pair, ascending-d fma chain) -> ids, q,
    // (z-q)^2, EMA statistics for the channel quads j == hf (mod TC_OCS).
    // ZREG (emb_dim known at compile time, <= 64): the thread keeps its 4*NZQ z values in registers, so the z stage
    // goes back to the TMA producer before the scan results even arrive; the re-rank reads z through shuffles.
    constexpr bool ZREG = DT != 0 && DT % (4 * TC_OCS) == 0 && DT / (4 * TC_OCS) <= 8;
    constexpr int NZQ = ZREG ? DT / (4 * TC_OCS) : 1;     // channel quads per thread
    const int ow = warp - TC_AUX_WARPS - TC_SCAN_WARPS;
    const int px = lane & (TC_OPX - 1), hf = lane >> TC_OPX_SHIFT;
    const int p = ow * TC_OPX + px;                           // pixel within the tile
    const float rminbig = __uint_as_float(P.meta[1]);   // +inf when no code is excluded
    const int bnsh = P.bn_shift;                          // BN == 1 << bnsh
    const uint32_t bn128 = (uint32_t)P.BN * 128;
    const int nq = Dc >> 2;
    // shared-memory address of z(p, d) = zrow + (d>>5)*16384 + (d&31)*128 + zx[d&3]   (see zs_off)
    const uint32_t zrow0 = sbase + P.off_z + (uint32_t)(p >> 5) * 4096 + ((p & 3) << 2);
    uint32_t zx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zx[i] = (uint32_t)i * 128 + (uint32_t)((((p & 31) >> 2) ^ (i << 1)) << 4);
    const uint32_t emain = sbase + P.off_emain;
    const uint32_t pub_s = sbase + P.off_pub + (uint32_t)p * 16;
    const uint32_t zn_s = sbase + P.off_zn;
    const uint32_t perm_a = sbase + P.off_perm;
    float* sums_mine = nullptr;
    if (STATS) {
      const int rep = (int)(blockIdx.x % (unsigned)P.nrep);
      sums_mine = rep == 0 ? P.sums : P.sums_rep + (size_t)(rep - 1) * P.K * Dc;
    }
    const size_t hw = (size_t)P.HW;
    const size_t img_stride = (size_t)Dc * hw;
    float2 ls2 = make_float2(0.f, 0.f);                   // two partial sums of (z-q)^2
    int tb = (int)blockIdx.x / P.tiles_per_img, tpt = (int)blockIdx.x % P.tiles_per_img;   // tile -> (image, tile in image)
    mbar_wait(BAR(0), 0);                                 // codebook resident (read below with plain loads)
    TC_TIMING_DECL
    for (int it = 0; it < my_tiles; ++it) {
      const int s = two_stages ? (it & 1) : 0, ph = two_stages ? ((it >> 1) & 1) : (it & 1);
      const int par = it & 1, pph = (it >> 1) & 1;
      const int b = tb, p0 = tpt * TC_TILE;
      tpt += (int)gridDim.x;
      while (tpt >= P.tiles_per_img) { tpt -= P.tiles_per_img; ++tb; }
      const uint32_t zst = s * zstage_bytes;
      const uint32_t zrow = zrow0 + zst;
      TC_TICK(4);
      mbar_wait(BAR(1 + s), ph);                          // z tile (TMA writes) visible to this thread
      float zq[NZQ][4];                                   // ZREG: z of channels 4*(TC_OCS*t+hf)+i
      if (ZREG) {
#pragma unroll
        for (int t = 0; t < NZQ; ++t) {
          // quad j = TC_OCS * t + hf (hf < TC_OCS divides 8): chunk and quad-in-chunk of TC_OCS * t, plus hf
          const uint32_t zj = zrow + (uint32_t)((TC_OCS * t) >> 3) * 16384 + (uint32_t)((TC_OCS * t) & 7) * 512 + (uint32_t)hf * 512;
#pragma unroll
          for (int i = 0; i < 4; ++i) zq[t][i] = lds_f32(zj + zx[i]);
        }
      }
      mbar_wait(BAR(9 + s), ph);                          // |z|^2
      float z2;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(z2) : "r"(zn_s + (uint32_t)(s * TC_TILE + p) * 4));
      // ZREG: the |z|^2 of the pixels this warp may re-rank (lane l keeps pixel l & 15), then the stage is free
      if (ZREG) {
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(3 + s));
      }
      TC_TICK(0);
      mbar_wait(BAR(11 + par), pph);                      // scan results of this tile
      TC_TICK(1);
      const float zn = sqrtf(z2) * 1.00001f;
      const bool bad = !(z2 <= 3.0e38f);

      // ---- merge the column groups of the pixel ---------------------------------------------------
      uint32_t pw[TC_NCG][4];
      float Lg = -INFINITY;
#pragma unroll
      for (int i = 0; i < TC_NCG; ++i) {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pw[i][0]), "=r"(pw[i][1]), "=r"(pw[i][2]), "=r"(pw[i][3])
                     : "r"(pub_s + (uint32_t)(par * TC_NCG + i) * (TC_TILE * 16)));
        Lg = fmaxf(Lg, __uint_as_float(pw[i][0]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(13 + par));          // the slot may be overwritten (tile it+2)
      int total = 0;
      bool ovf = false;
#pragma unroll
      for (int i = 0; i < TC_NCG; ++i) {
        const bool al = __uint_as_float(pw[i][1] & 0xFFFF0000u) >= Lg;
        if (!al) { pw[i][2] = 0; pw[i][3] = 0; }
        else ovf |= (pw[i][1] & 1u) != 0;
        total += __popc(pw[i][2]) + __popc(pw[i][3]);
      }
      const float lbest = 2.f * (Lg < 0.f ? Lg * 1.0009765625f : Lg);   // lower bound on the best exact 2 a_k (accumulators are a_k / (1 + 2^-10))
      // excluded ("big") codes: s_k <= r_k (2|z| - r_k), decreasing in r_k for r_k >= |z|
      bool big_safe = true;
      if (rminbig < 3.0e38f) {
        const float bigub = rminbig * (2.f * zn - rminbig);
        big_safe = (rminbig >= zn) && (bigub + 1e-5f * (fabsf(bigub) + fabsf(lbest)) < lbest);
      }
      bool fb = bad || ovf || total == 0 || !big_safe;
      int w = 0;                                          // winner: position in the norm-sorted codebook
      if (total == 1) {
#pragma unroll
        for (int i = 0; i < TC_NCG; ++i) {
          if (pw[i][2]) w = (int)(((pw[i][1] >> 8) & 0x7Fu) * 32u) + __clz(pw[i][2]);
          if (pw[i][3]) w = (int)(((pw[i][1] >> 1) & 0x7Fu) * 32u) + __clz(pw[i][3]);
        }
      }
      TC_TICK(2);
      // ---- pixels with several candidates: exact re-rank by the pixel's own two lanes -----------------------
      // Exact dot product (all kernels of this library): dot = A + B, A / B = ascending-d fma chains over the even /
      // odd channel quads.  Lane hf of the pixel owns the quads of parity hf, so each lane runs one chain over its own
      // z values (registers when ZREG) and one shuffle joins the halves; both lanes then take identical decisions.
      // The warp iterates until its busiest pixel is done (usually two candidates).
      int rem = (!fb && total > 1) ? total : 0;
      if (rem > TC_MAXCAND) { fb = true; rem = 0; }      // too many ties: exhaustive search for this pixel
#ifdef VQ_ABL_NORERANK
      rem = 0;
#endif
#ifdef VQ_TC_TIMING
      tacc[7] += __reduce_add_sync(0xffffffffu, (hf == 0) ? rem : 0);
#endif
      if (__any_sync(0xffffffffu, rem > 0)) {
        static_assert(TC_NCG <= 2, "candidate cascade below handles two column groups");
        uint32_t m0 = pw[0][2], m1 = pw[0][3], m2 = TC_NCG > 1 ? pw[TC_NCG - 1][2] : 0u, m3 = TC_NCG > 1 ? pw[TC_NCG - 1][3] : 0u;
        const uint32_t c0 = ((pw[0][1] >> 8) & 0x7Fu) * 32u, c1 = ((pw[0][1] >> 1) & 0x7Fu) * 32u;
        const uint32_t c2 = ((pw[TC_NCG - 1][1] >> 8) & 0x7Fu) * 32u, c3 = ((pw[TC_NCG - 1][1] >> 1) & 0x7Fu) * 32u;
        unsigned long long key = 0ull;                    // best (score, -original index, position) so far
        do {
          // next candidate of my pixel (idle pixels score code 0 and drop the result)
          const uint32_t mm = m0 ? m0 : m1 ? m1 : m2 ? m2 : m3;
          const uint32_t cbase = m0 ? c0 : m1 ? c1 : m2 ? c2 : c3;
          const bool act = rem > 0;
          const int jb = act ? __clz(mm) : 0;
          const uint32_t clr = act ? ~(0x80000000u >> jb) : 0xFFFFFFFFu;
          if (m0) m0 &= clr; else if (m1) m1 &= clr; else if (m2) m2 &= clr; else m3 &= clr;
          const int k = act ? (int)(cbase + (uint32_t)jb) : 0;
          const int kb = k >> bnsh, row = k & (P.BN - 1);
          const uint32_t eb = emain + (uint32_t)(kb * nD) * bn128 + (uint32_t)row * 128;
          const uint32_t r7 = (uint32_t)(row & 7) << 4;
          float dot = 0.f;
          if (ZREG) {
#pragma unroll
            for (int t = 0; t < NZQ; ++t) {
              const uint32_t j = (uint32_t)(2 * t) + (uint32_t)hf;
              const float4 e4 = lds_v4(eb + (uint32_t)((2 * t) >> 3) * bn128 + (((j & 7) << 4) ^ r7));
              dot = __fmaf_rn(zq[t][0], e4.x, dot);
              dot = __fmaf_rn(zq[t][1], e4.y, dot);
              dot = __fmaf_rn(zq[t][2], e4.z, dot);
              dot = __fmaf_rn(zq[t][3], e4.w, dot);
            }
          } else {
            for (int j = hf; j < nq; j += 2) {
              const float4 e4 = lds_v4(eb + (uint32_t)(j >> 3) * bn128 + ((((uint32_t)j & 7) << 4) ^ r7));
              const uint32_t zj = zrow + (uint32_t)(j >> 3) * 16384 + (uint32_t)(j & 7) * 512;
              dot = __fmaf_rn(lds_f32(zj + zx[0]), e4.x, dot);
              dot = __fmaf_rn(lds_f32(zj + zx[1]), e4.y, dot);
              dot = __fmaf_rn(lds_f32(zj + zx[2]), e4.z, dot);
              dot = __fmaf_rn(lds_f32(zj + zx[3]), e4.w, dot);
            }
          }
          dot = __fadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, 16));     // A + B (commutative: same bits in both lanes)
          uint32_t korig;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(korig) : "r"(perm_a + (uint32_t)k * 2));
          // exact |e|^2: fourth float of the code's augmentation entry (it meets the zero row of the ones block in the MMA)
          const float4 au = lds_v4(sbase + P.off_eaug + (uint32_t)(((kb << bnsh) << 5) + (row >> 3) * 256 + (row & 7) * 16));
          const float e2k = au.w;
          const float sc = ref_score(dot, e2k, z2);
          // ties go to the lowest ORIGINAL index
          const unsigned long long kcur =
              ((unsigned long long)f32_orderable(sc) << 32) | ((unsigned long long)(0xFFFFu - korig) << 16) | (unsigned long long)k;
          if (act && kcur > key) key = kcur;
          rem -= act ? 1 : 0;
        } while (__any_sync(0xffffffffu, rem > 0));
        if (!fb && total > 1) w = (int)(key & 0xFFFFull);
      }

      TC_TICK(3);
      // ---- outputs: ids, q, (z-q)^2, EMA statistics ----------------------------------------------------
      const int pp = p0 + p;
      if (fb) {
        if (hf == 0) {
          const int slot = atomicAdd(P.fb_count, 1);
          P.fb_rows[slot] = b * P.HW + pp;
        }
      } else {
        uint32_t worig;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(worig) : "r"(perm_a + (uint32_t)w * 2));
        if (hf == 0) {
          int h, wc;
          if (P.w_shift >= 0) { h = pp >> P.w_shift; wc = pp & (P.W - 1); }
          else { h = pp / P.W; wc = pp - h * P.W; }
          const size_t nb_ = (size_t)b * hw;
          if (P.ids) store_id(P.ids + nb_, pp, h, wc, P.H, (int)worig, P.ids_mode);
          if (P.ids_nat) P.ids_nat[nb_ + pp] = (int)worig;
          if (STATS) atomicAdd(&hist[worig], 1);
        }
        const int kb = w >> bnsh, row = w & (P.BN - 1);
        const uint32_t r7 = (uint32_t)(row & 7);
        // channel quads j = TC_OCS*t + hf: chunk j >> 3, quad-in-chunk j & 7
        const uint32_t ea = emain + (uint32_t)(kb * nD) * bn128 + (uint32_t)row * 128;
        float* qo = P.q + (size_t)b * img_stride + pp;
        float* so = STATS ? sums_mine + (size_t)worig * Dc : nullptr;
        auto quad_out = [&](int j, float z0, float z1, float z2v, float z3) {
          const float4 e4 = lds_v4(ea + (uint32_t)(j >> 3) * bn128 + ((((uint32_t)j & 7) ^ r7) << 4));
          const float2 m1 = make_float2(-1.f, -1.f);      // z - e as one packed fma (exact: e * -1 + z)
          const float2 d01 = __ffma2_rn(make_float2(e4.x, e4.y), m1, make_float2(z0, z1));
          const float2 d23 = __ffma2_rn(make_float2(e4.z, e4.w), m1, make_float2(z2v, z3));
          ls2 = __ffma2_rn(d01, d01, ls2);
          ls2 = __ffma2_rn(d23, d23, ls2);
#ifndef VQ_ABL_NOQ
          if (!DBG || P.q)
#else
          if (false)
#endif
          {
            float* qj = qo + (size_t)(4 * j) * hw;
            __stcs(qj, e4.x);
            __stcs(qj + hw, e4.y);
            __stcs(qj + 2 * hw, e4.z);
            __stcs(qj + 3 * hw, e4.w);
          }
          if (STATS) atomicAdd(reinterpret_cast<float4*>(so + 4 * j), make_float4(z0, z1, z2v, z3));
        };
#ifndef VQ_ABL_NOOUT
        if (ZREG) {
#pragma unroll
          for (int t = 0; t < NZQ; ++t) quad_out(TC_OCS * t + hf, zq[t][0], zq[t][1], zq[t][2], zq[t][3]);
        } else {
#pragma unroll 1
          for (int j = hf; j < nq; j += TC_OCS) {
            const uint32_t zj = zrow + (uint32_t)(j >> 3) * 16384 + (uint32_t)(j & 7) * 512;
            quad_out(j, lds_f32(zj + zx[0]), lds_f32(zj + zx[1]), lds_f32(zj + zx[2]), lds_f32(zj + zx[3]));
          }
        }
#endif
      }
      if (!ZREG) {
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(3 + s));           // z stage free
      }
    }
    TC_TICK(4);
    TC_TIMING_STORE(8 + ow, my_tiles);
    float lsum = ls2.x + ls2.y;
    // ---- per-CTA reductions ------------------------------------------------------------------
    lsum = warp_sum(lsum);
    if (lane == 0 && P.loss_acc && lsum != 0.f) atomicAdd(P.loss_acc, (double)lsum);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_OUT_WARPS) : "memory");        // all output warps
    if (STATS) {
      for (int k = threadIdx.x - 32 * (TC_AUX_WARPS + TC_SCAN_WARPS); k < P.K; k += 32 * TC_OUT_WARPS) {
        const int c = hist[k];
        if (c) atomicAdd(&P.counts[k], c);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// streaming variant: codebooks that do not fit in shared memory (K*D*4 > ~140 KB, e.g. K=512 x D=256, K=4096 x D=64)
// ---------------------------------------------------------------------------------------------
// Nothing is resident: one ring of stages, each stage = one 32-channel chunk of the z tile (128 pixels x 128 B) plus the
// matching [BN codes x 32 channels] slice of the norm-sorted codebook (classic K-loop pipelining, one tcgen05.commit per
// stage).  The z tile is streamed once per code block (it comes out of L2 after the first pass), so all of the shared
// memory holds bytes in flight -- with a resident z tile (128 KB at D = 256) only two codebook stages fitted and the
// kernel was bound by TMA latency.  Scan and merge are the same as in the resident kernel; the output warps read z and
// the exact fp32 code rows from global memory (both L2 hits: the tile and the codebook were just streamed).
struct TcsGeom {
  int BN, nb, nD, nst;
  size_t stage_bytes, off_stage, off_aug, off_aaug, off_pub, off_zn, off_ctab, off_qst, off_bar, total;
  bool ok;
};
constexpr int TCS_QST_BYTES = 16 * TC_DCH * 4;      // q staging box of one epilogue warp: 16 pixels x 32 channels
constexpr int TCS_MAX_ND = 8;       // D <= 256
constexpr int TCS_MAX_ST = 8;       // ring stages
constexpr int TCS_MAXCAND = 16;     // candidates re-scored exactly per pixel (large codebooks tie more often)
// barrier slots of the streaming kernel
constexpr int TCS_B_FULL = 0, TCS_B_EMPTY = 8, TCS_B_AFULL = 16, TCS_B_AEMPTY = 18, TCS_B_TFULL = 20, TCS_B_TEMPTY = 22,
              TCS_B_ZN = 24, TCS_B_TMEM = 30, TCS_B_CONS = 32;

// pair = CTA-pair mode (cta_group::2, BN = 256 only): each CTA of a 2-CTA cluster keeps its own 128-pixel z chunk but
// only HALF of every codebook slice (the tensor cores of both SMs read both halves), so the codebook crosses
// L2 -> SM once per 256 pixels instead of once per 128
static TcsGeom tcs_geometry(int D, int K, bool pair = false) {
  TcsGeom g{};
  g.ok = false;
  g.BN = 32;
  while (g.BN < K && g.BN < TC_MAXBN) g.BN <<= 1;
  g.nb = (K + g.BN - 1) / g.BN;
  g.nD = (D + TC_DCH - 1) / TC_DCH;
  if (g.nD > TCS_MAX_ND || g.nb * g.BN > TC_SORT_MAX) return g;
  if (pair && g.BN != TC_MAXBN) return g;
  const int bmine = pair ? g.BN / 2 : g.BN;                             // codes of a slice held by this CTA
  g.stage_bytes = (size_t)TC_TILE * 128 + (size_t)bmine * 128;         // z chunk + codebook slice (both 1024-aligned)
  size_t off = 0;
  g.off_aug = off;  off += align_up((size_t)2 * bmine * 32, 1024);
  g.off_aaug = off; off += 4096;
  g.off_stage = off;
  const size_t sz_pub = (size_t)4 * TC_NCG * TC_TILE * 16, sz_zn = 2 * TC_TILE * 4;   // pub: [team][tile parity][column group][pixel]
  const size_t sz_ctab = align_up((size_t)g.nb * (g.BN / 32) * 8, 16);
  const size_t sz_qst = (size_t)(TC_SCAN_WARPS + TC_OUT_WARPS) * TCS_QST_BYTES;
  const size_t tail = sz_pub + sz_zn + align_up(sz_ctab, 128) + sz_qst + 512;
  long long room = (long long)TC_SMEM_LIMIT - 1024 - (long long)off - (long long)tail;
  int nst = (int)(room / (long long)g.stage_bytes);
  if (nst > TCS_MAX_ST) nst = TCS_MAX_ST;
  // The |z|^2 slot (and barrier) of tile it is reused for tile it+2, so |z|^2(it+2) must not complete before the team of
  // tile it has read |z|^2(it).  The team reads it before it scans block (it, 0); the MMA of block it*nb + 2 starts only
  // after that scan (TMEM has two stages); |z|^2(it+2) needs the last chunk of block (it+2, 0) loaded, and the producer
  // is at most nst stages ahead of the tensor core: nst <= ((it+2) nb nD + nD - 1) - (it nb + 2) nD = (2 nb - 1) nD - 1.
  if (nst > (2 * g.nb - 1) * g.nD - 1) nst = (2 * g.nb - 1) * g.nD - 1;
  if (nst < 2) return g;
  g.nst = nst;
  off += (size_t)nst * g.stage_bytes;
  g.off_pub = off;  off += sz_pub;
  g.off_zn = off;   off += sz_zn;
  g.off_ctab = off; off += align_up(sz_ctab, 128);
  g.off_qst = off;  off += sz_qst;
  g.off_bar = off;  off += 512;
  g.total = off + 1024;
  g.ok = true;
  return g;
}

struct TcsParams {
  const float* z; const float* E; const float* e2; const float* eaug_img; const uint32_t* meta;
  const int* perm; const uint32_t* rmax;
  int B, D, H, W, HW, K;
  int BN, nb, nD, nst;
  int bn_shift, w_shift;
  int tiles_per_img; int ntiles;
  uint32_t stage_bytes, off_stage, off_aug, off_aaug, off_pub, off_zn, off_ctab, off_qst, off_bar;
  int64_t* ids; int32_t* ids_nat; float* q; double* loss_acc; int* counts;
  float* sums; float* sums_rep; int nrep;
  int* fb_count; int* fb_rows;
  int ids_mode;
  float* dbg;
};

template <bool DBG, bool STATS, bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
vq_assign_tcs_kernel(const __grid_constant__ CUtensorMap zmap, const __grid_constant__ CUtensorMap emap,
                     const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap qmap, const TcsParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = (uint64_t*)(smem + P.off_bar);
  const uint32_t bar0 = sbase + P.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + TCS_B_TMEM);
  const int nD = P.nD, nst = P.nst;
  const int ktot = P.nb * P.BN;
  const uint32_t stage_bytes = P.stage_bytes;                 // z chunk (16 KB) followed by the codebook slice
  // CTA pair: rank 0 (the leader) issues the M = 256 MMAs for both CTAs; work unit = a pair of adjacent tiles
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int unit0 = PAIR ? (int)cluster_idx() : (int)blockIdx.x;
  const int ustride = PAIR ? (int)cluster_count() : (int)gridDim.x;
  const int nunits = PAIR ? (P.ntiles + 1) >> 1 : P.ntiles;
  const int bmine = PAIR ? P.BN >> 1 : P.BN;                   // codes of every slice held in this CTA's shared memory

  if (threadIdx.x == 32) {
    // single CTA: a stage is released by the MMA commit + the two |z|^2 warps.  Pair: the commit arrives on CONS (both
    // CTAs), the |z|^2 warps wait for it (the peer CTA never sees the leader's FULL barrier) and release the stage.
    for (int i = 0; i < TCS_MAX_ST; ++i) {
      mbar_init(BAR(TCS_B_FULL + i), 1); mbar_init(BAR(TCS_B_EMPTY + i), PAIR ? 2 : 3); mbar_init(BAR(TCS_B_CONS + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(TCS_B_AFULL + i), 1); mbar_init(BAR(TCS_B_AEMPTY + i), 1);
      mbar_init(BAR(TCS_B_TFULL + i), 1); mbar_init(BAR(TCS_B_TEMPTY + i), PAIR ? 2 * TC_SCAN_WARPS : TC_SCAN_WARPS);
      mbar_init(BAR(TCS_B_ZN + i), 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (warp >= TC_AUX_WARPS) {
    const int t = threadIdx.x - 32 * TC_AUX_WARPS;
    constexpr int NT = 32 * (TC_SCAN_WARPS + TC_OUT_WARPS);
    float4* a = (float4*)(smem + P.off_aaug);
    for (int i = t; i < 256; i += NT) {                   // ones block of the augmentation K-step (see the resident kernel)
      const int row = (i >> 3) & 7;
      const float v = row < 3 ? 1.f : 0.f;
      a[i] = make_float4(v, v, v, v);
    }
    const float c1 = 0.001953125f * 1.03f;                      // see the resident kernel
    const float c2 = (float)(P.D + 16) * 4.76837158e-7f;
    for (int c = t; c < P.nb * (P.BN >> 5); c += NT) {
      const float rm = __uint_as_float(P.rmax[c]);
      ((float*)(smem + P.off_ctab))[2 * c] = 0.5f * (c1 + c2) * rm;
      ((float*)(smem + P.off_ctab))[2 * c + 1] = 0.5f * c2 * rm * rm + 1e-30f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();               // the peer's barriers must exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (nunits - unit0 + ustride - 1) / ustride;        // work units (tiles, or tile pairs) of this CTA
  // tile of my it-th unit; in pair mode the odd CTA of the last pair may get a phantom tile (>= ntiles): it is loaded
  // (out-of-range boxes are zero-filled) and scanned like any other, and produces no output
  auto tile_of = [&](int it) { return PAIR ? 2 * (unit0 + it * ustride) + (int)crank : unit0 + it * ustride; };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int scount = 0, acount = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = tile_of(it);
        const int b = tile / P.tiles_per_img, pt = tile % P.tiles_per_img;
        for (int blk = 0; blk < P.nb; ++blk) {
          {   // this block's -|e|^2/2 image
            const int as = acount & 1;
            mbar_wait_sleep(BAR(TCS_B_AEMPTY + as), ((acount >> 1) & 1) ^ 1, 64);
            if (PAIR) {      // both halves are counted on the leader's barrier
              if (crank == 0) mbar_expect_tx(BAR(TCS_B_AFULL + as), (uint32_t)P.BN * 32);
              tma_load_2d_pair(sbase + P.off_aug + (uint32_t)as * bmine * 32, &amap, leader_addr(BAR(TCS_B_AFULL + as)), 0,
                               (blk * P.BN + (int)crank * bmine) >> 3);
            } else {
              mbar_expect_tx(BAR(TCS_B_AFULL + as), (uint32_t)P.BN * 32);
              bulk_load_1d(sbase + P.off_aug + (uint32_t)as * P.BN * 32, P.eaug_img + (size_t)blk * P.BN * 8, (uint32_t)P.BN * 32,
                           BAR(TCS_B_AFULL + as));
            }
            ++acount;
          }
          for (int c = 0; c < nD; ++c, ++scount) {
            const int st = scount % nst;
            mbar_wait_sleep(BAR(TCS_B_EMPTY + st), ((scount / nst) & 1) ^ 1, 32);
            const uint32_t dst = sbase + P.off_stage + (uint32_t)st * stage_bytes;
            if (PAIR) {
              if (crank == 0) mbar_expect_tx(BAR(TCS_B_FULL + st), 2u * stage_bytes);     // my stage and the peer's
              const uint32_t fb = leader_addr(BAR(TCS_B_FULL + st));
              for (int grp = 0; grp < 4; ++grp)
                tma_load_3d_pair(dst + grp * 4096, &zmap, fb, pt * TC_TILE + grp * 32, c * TC_DCH, b);
              tma_load_2d_pair(dst + TC_TILE * 128, &emap, fb, c * TC_DCH, blk * P.BN + (int)crank * bmine);
            } else {
              mbar_expect_tx(BAR(TCS_B_FULL + st), stage_bytes);
              for (int grp = 0; grp < 4; ++grp)             // z chunk: four 32-pixel x 32-channel boxes
                tma_load_3d(dst + grp * 4096, &zmap, BAR(TCS_B_FULL + st), pt * TC_TILE + grp * 32, c * TC_DCH, b);
              tma_load_2d(dst + TC_TILE * 128, &emap, BAR(TCS_B_FULL + st), c * TC_DCH, blk * P.BN);
            }
          }
        }
      }
      if (PAIR) {     // the leader's last commits arrive on this CTA's AEMPTY barriers: let them land before the CTA exits
        for (int k = acount > 2 ? acount - 2 : 0; k < acount; ++k) mbar_wait_sleep(BAR(TCS_B_AEMPTY + (k & 1)), (k >> 1) & 1, 64);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0 && crank == 0) {
      const uint32_t idesc = PAIR ? make_idesc_pair(P.BN) : make_idesc(P.BN);
      int g = 0, scount = 0, acount = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int blk = 0; blk < P.nb; ++blk, ++g) {
          const int a = g & 1, aph = (g >> 1) & 1;
          if (PAIR) mbar_wait_sleep_cluster(BAR(TCS_B_TEMPTY + a), aph ^ 1, 32);       // scan warps of both CTAs
          else mbar_wait_sleep(BAR(TCS_B_TEMPTY + a), aph ^ 1, 32);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)a * TC_MAXBN;
          uint32_t acc = 0;
          for (int c = 0; c < nD; ++c, ++scount) {
            const int st = scount % nst;
            mbar_wait_sleep(BAR(TCS_B_FULL + st), (scount / nst) & 1, 32);
            tc_fence_after();
            const int ksteps = min(4, (P.D - c * TC_DCH + 7) >> 3);
            const uint32_t zaddr = sbase + P.off_stage + (uint32_t)st * stage_bytes;
            const uint32_t eaddr = zaddr + TC_TILE * 128;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t ad = make_desc(zaddr + ks * 1024, 4096, 512, 1);
              const uint64_t bd = make_desc(eaddr + ks * 32, 16, 1024, 2);
              if (PAIR) umma_tf32_pair(d_tmem, ad, bd, idesc, acc);
              else umma_tf32(d_tmem, ad, bd, idesc, acc);
              acc = 1;
            }
            if (PAIR) umma_commit_pair(BAR(TCS_B_CONS + st));
            else umma_commit(BAR(TCS_B_EMPTY + st));       // stage consumed by the tensor core
          }
          {
            const int as = acount & 1;
            mbar_wait_sleep(BAR(TCS_B_AFULL + as), (acount >> 1) & 1, 32);
            tc_fence_after();
            const uint64_t ad = make_desc(sbase + P.off_aaug, 1024, 512, 1);
            const uint64_t bd = make_desc(sbase + P.off_aug + (uint32_t)as * bmine * 32, 128, 256, 0);
            if (PAIR) { umma_tf32_pair(d_tmem, ad, bd, idesc, acc); umma_commit_pair(BAR(TCS_B_AEMPTY + as)); }
            else { umma_tf32(d_tmem, ad, bd, idesc, acc); umma_commit(BAR(TCS_B_AEMPTY + as)); }
            ++acount;
          }
          if (PAIR) umma_commit_pair(BAR(TCS_B_TFULL + a));
          else umma_commit(BAR(TCS_B_TFULL + a));
        }
      }
    }
  } else if (warp < TC_AUX_WARPS) {
    // ===================================== |z|^2 workers ====================================
    // read the z chunks of the first code block's pass; the later passes of the tile only hand the stage back
    const int pA = (warp - 2) * 64 + 2 * lane;
    const uint32_t zrow0 = sbase + P.off_stage + (uint32_t)(pA >> 5) * 4096 + ((pA & 3) << 2);
    const uint32_t zn_s = sbase + P.off_zn;
    uint32_t zx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zx[i] = (uint32_t)i * 128 + (uint32_t)((((pA & 31) >> 2) ^ (i << 1)) << 4);
    int scount = 0;
    for (int it = 0; it < my_tiles; ++it) {
      float2 zz = make_float2(0.f, 0.f), zzB = make_float2(0.f, 0.f);    // |z|^2 = A + B (even / odd channel quads)
      for (int blk = 0; blk < P.nb; ++blk) {
        for (int c = 0; c < nD; ++c, ++scount) {
          const int st = scount % nst;
          // always wait for the fill: an arrival may only count for the phase it belongs to (the producer refills a
          // stage after the previous phase of its EMPTY barrier completed)
          mbar_wait(BAR((PAIR ? TCS_B_CONS : TCS_B_FULL) + st), (scount / nst) & 1);
          if (blk == 0) {
            const uint32_t zc = zrow0 + (uint32_t)st * stage_bytes;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 v = lds_v2(zc + jj * 512 + zx[i]);
                if ((jj & 1) == 0) zz = __ffma2_rn(v, v, zz);
                else zzB = __ffma2_rn(v, v, zzB);
              }
            }
            __syncwarp();
          }
          if (lane == 0) mbar_arrive(BAR(TCS_B_EMPTY + st));
        }
        if (blk == 0) {
          const int sl = it & 1;
          zz = __fadd2_rn(zz, zzB);
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(zn_s + (uint32_t)(sl * TC_TILE + pA) * 4), "f"(zz.x), "f"(zz.y) : "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(TCS_B_ZN + sl));
        }
      }
    }
  } else {
    // ===================================== epilogue warps: two teams of eight ================================
    // Team t takes the tiles it = t, t+2, ... of this CTA and does everything for them: scan (warp = TMEM lane quadrant
    // x column group, as in the resident kernel), merge, exact re-rank, outputs.  The scan of a tile is short next to
    // its output phase here (z and the code rows come from L2, D/8 dependent round trips per lane), so while one team
    // is writing a tile out the other one already scans and writes the next: twice the loads in flight.
    const int ew = warp - TC_AUX_WARPS, team = ew >> 3, tw = ew & 7;
    const int quad = warp & 3, cg = tw >> 2;              // warp % 4 == TMEM lane quadrant this warp may read
    const int p = quad * 32 + lane;                       // scan: pixel within the tile == TMEM lane
    const uint32_t ctab_s = sbase + P.off_ctab;
    const uint32_t zn_s = sbase + P.off_zn + (uint32_t)(team * TC_TILE + p) * 4;
    const uint32_t qst_s = sbase + P.off_qst + (uint32_t)ew * TCS_QST_BYTES;
    const int nchunks = P.BN >> 5;
    // output phase: lane = (pixel px, quad parity hf) of the 16 pixels quad*32 + cg*16 .. +15
    const int px = lane & (TC_OPX - 1), hf = lane >> TC_OPX_SHIFT;
    const int po = quad * 32 + cg * TC_OPX + px;
    const float rminbig = __uint_as_float(P.meta[1]);
    const int D = P.D, nq = D >> 2;
    float* sums_mine = nullptr;
    if (STATS) {
      const int rep = (int)((blockIdx.x * 2 + team) % (unsigned)P.nrep);
      sums_mine = rep == 0 ? P.sums : P.sums_rep + (size_t)(rep - 1) * P.K * D;
    }
    const size_t hw = (size_t)P.HW;
    const size_t img_stride = (size_t)D * hw;
    float2 ls2 = make_float2(0.f, 0.f);
    TC_TIMING_DECL
    for (int it = team; it < my_tiles; it += 2) {
      const int tile = tile_of(it);
      const bool phantom = PAIR && tile >= P.ntiles;
      const int b = tile / P.tiles_per_img, p0 = (tile % P.tiles_per_img) * TC_TILE;
      const int n2 = it >> 1;                             // this team's tile counter
      TC_TICK(5);
      mbar_wait(BAR(TCS_B_ZN + team), n2 & 1);
      TC_TICK(0);
      float z2s;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(z2s) : "r"(zn_s));
      const float zn_scan = sqrtf(z2s) * 1.00001f;
      float L = -INFINITY, Urec = -INFINITY;
      int cnt = 0;
      uint32_t rcA = 0, rcB = 0, rm0 = 0, rm1 = 0;
      for (int blk = 0; blk < P.nb; ++blk) {
        const int g = it * P.nb + blk;
        const int a = g & 1, aph = (g >> 1) & 1;
        TC_TICK(2);
        mbar_wait(BAR(TCS_B_TFULL + a), aph);
        TC_TICK(1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)a * TC_MAXBN;
        for (int c = cg; c < nchunks; c += TC_NCG) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          const int gc = blk * nchunks + c;
          float cA, cB;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cA), "=f"(cB) : "r"(ctab_s + (uint32_t)gc * 8));
          const float delta = __fmaf_rn(zn_scan, cA, cB);
          tmem_ld_wait();
          if (DBG && !phantom) {
            float* o = P.dbg + ((size_t)b * P.HW + p0 + p) * ktot + gc * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = v[j];
          }
          float m4[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float m = fmaxf(fmaxf(v[8 * h], v[8 * h + 1]), v[8 * h + 2]);
            m = fmaxf(fmaxf(m, v[8 * h + 3]), v[8 * h + 4]);
            m = fmaxf(fmaxf(m, v[8 * h + 5]), v[8 * h + 6]);
            m4[h] = fmaxf(m, v[8 * h + 7]);
          }
          const float cm = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          L = fmaxf(L, cm - delta);
          if (Urec < L) { cnt = 0; Urec = -INFINITY; }
          const float T = L - delta;
          const float2 nT2 = make_float2(-T, -T);
          uint32_t n4[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            uint32_t nm = 0;
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const float2 d2 = __fadd2_rn(make_float2(v[8 * h + j], v[8 * h + j + 1]), nT2);
              nm = __funnelshift_l(__float_as_uint(d2.x), nm, 1);
              nm = __funnelshift_l(__float_as_uint(d2.y), nm, 1);
            }
            n4[h] = nm;
          }
          const uint32_t cand = ~((n4[0] << 24) | (n4[1] << 16) | (n4[2] << 8) | n4[3]);
          const bool has = cand != 0u;
          const bool s0 = has && cnt == 0, s1 = has && cnt == 1;
          rcA = s0 ? (uint32_t)gc : rcA; rm0 = s0 ? cand : rm0;
          rcB = s1 ? (uint32_t)gc : rcB; rm1 = s1 ? cand : rm1;
          cnt += has ? 1 : 0;
          Urec = has ? fmaxf(Urec, cm + delta) : Urec;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(BAR(TCS_B_TEMPTY + a));        // the leader's MMA warp waits for both CTAs' scans
          else mbar_arrive(BAR(TCS_B_TEMPTY + a));
        }
      }
      // ---- publish to the team: {L, U16 | chunkA<<8 | chunkB<<1 | overflow, maskA, maskB}, double-buffered by tile parity
      if (cnt < 2) rm1 = 0;
      if (cnt < 1) rm0 = 0;
      const uint32_t w1 = f32_up16(Urec) | (rcA << 8) | (rcB << 1) | (cnt > 2 ? 1u : 0u);
      const uint32_t pub_t = sbase + P.off_pub + (uint32_t)((team * 2 + (n2 & 1)) * TC_NCG) * (TC_TILE * 16);
      TC_TICK(2);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(pub_t + (uint32_t)(cg * TC_TILE + p) * 16),
                   "r"(__float_as_uint(L)), "r"(w1), "r"(rm0), "r"(rm1) : "memory");
      // team barrier (hardware named barrier: a waiting warp issues nothing).  A warp cannot be two tiles ahead of a
      // team mate (it needs the mate's arrival at the next barrier), so two publication buffers per team are enough.
      if (team == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 256;" ::: "memory");
      if (phantom) continue;                               // (warp-uniform) nothing to write for a tile past the end

      // ---- merge the column groups of my output pixel ---------------------------------------------------------
      const float z2 = __shfl_sync(0xffffffffu, z2s, cg * TC_OPX + px);     // scan lane (pixel po) of this very warp
      const float zn = sqrtf(z2) * 1.00001f;
      const bool bad = !(z2 <= 3.0e38f);
      const int pp = p0 + po;
      const float* zp = P.z + (size_t)b * img_stride + pp;           // z(pixel, channel d) = zp[d * hw]
      uint32_t pw[TC_NCG][4];
      float Lg = -INFINITY;
#pragma unroll
      for (int i = 0; i < TC_NCG; ++i) {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pw[i][0]), "=r"(pw[i][1]), "=r"(pw[i][2]), "=r"(pw[i][3])
                     : "r"(pub_t + (uint32_t)(i * TC_TILE + po) * 16));
        Lg = fmaxf(Lg, __uint_as_float(pw[i][0]));
      }
      int total = 0;
      bool ovf = false;
#pragma unroll
      for (int i = 0; i < TC_NCG; ++i) {
        const bool al = __uint_as_float(pw[i][1] & 0xFFFF0000u) >= Lg;
        if (!al) { pw[i][2] = 0; pw[i][3] = 0; }
        else ovf |= (pw[i][1] & 1u) != 0;
        total += __popc(pw[i][2]) + __popc(pw[i][3]);
      }
      const float lbest = 2.f * (Lg < 0.f ? Lg * 1.0009765625f : Lg);
      bool big_safe = true;
      if (rminbig < 3.0e38f) {
        const float bigub = rminbig * (2.f * zn - rminbig);
        big_safe = (rminbig >= zn) && (bigub + 1e-5f * (fabsf(bigub) + fabsf(lbest)) < lbest);
      }
      bool fb = bad || ovf || total == 0 || !big_safe;
      int w = 0;
      if (total == 1) {
#pragma unroll
        for (int i = 0; i < TC_NCG; ++i) {
          if (pw[i][2]) w = (int)(((pw[i][1] >> 8) & 0x7Fu) * 32u) + __clz(pw[i][2]);
          if (pw[i][3]) w = (int)(((pw[i][1] >> 1) & 0x7Fu) * 32u) + __clz(pw[i][3]);
        }
      }
      int worig = __ldg(P.perm + w);
      TC_TICK(3);
      // ---- pixels with several candidates: exact fp32 re-rank by the pixel's own two lanes (see the resident kernel) ----
      // (a variant with the whole warp on one (pixel, code) pair and the fma chains travelling from lane to lane was
      //  slower: at D = 256 a warp has 3-4 such pixels per tile and they are better served in parallel)
      int rem = (!fb && total > 1) ? total : 0;
      if (rem > TCS_MAXCAND) { fb = true; rem = 0; }
      if (__any_sync(0xffffffffu, rem > 0)) {
        uint32_t m0 = pw[0][2], m1 = pw[0][3], m2 = TC_NCG > 1 ? pw[TC_NCG - 1][2] : 0u, m3 = TC_NCG > 1 ? pw[TC_NCG - 1][3] : 0u;
        const uint32_t c0 = ((pw[0][1] >> 8) & 0x7Fu) * 32u, c1 = ((pw[0][1] >> 1) & 0x7Fu) * 32u;
        const uint32_t c2 = ((pw[TC_NCG - 1][1] >> 8) & 0x7Fu) * 32u, c3 = ((pw[TC_NCG - 1][1] >> 1) & 0x7Fu) * 32u;
        unsigned long long key = 0ull;
        do {
          const uint32_t mm = m0 ? m0 : m1 ? m1 : m2 ? m2 : m3;
          const uint32_t cbase = m0 ? c0 : m1 ? c1 : m2 ? c2 : c3;
          const bool act = rem > 0;
          const int jb = act ? __clz(mm) : 0;
          const uint32_t clr = act ? ~(0x80000000u >> jb) : 0xFFFFFFFFu;
          if (m0) m0 &= clr; else if (m1) m1 &= clr; else if (m2) m2 &= clr; else m3 &= clr;
          const int k = act ? (int)(cbase + (uint32_t)jb) : 0;
          const int korig = __ldg(P.perm + k);
          float dot = 0.f;
          if (act) {                                      // divergent on purpose: idle pixels issue no loads
            const float4* er = reinterpret_cast<const float4*>(P.E + (size_t)korig * D);
            for (int j0 = hf; j0 < nq; j0 += 8) {         // four of my quads' loads, then their sixteen chained fmas
              float4 e4[4];
              float zv[4][4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int j = j0 + 2 * t;
                const bool in = j < nq;
                e4[t] = in ? __ldg(er + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float* zj = zp + (size_t)(4 * j) * hw;
                zv[t][0] = in ? __ldg(zj) : 0.f;          zv[t][1] = in ? __ldg(zj + hw) : 0.f;
                zv[t][2] = in ? __ldg(zj + 2 * hw) : 0.f; zv[t][3] = in ? __ldg(zj + 3 * hw) : 0.f;
              }
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                if (j0 + 2 * t < nq) {
                  dot = __fmaf_rn(zv[t][0], e4[t].x, dot);
                  dot = __fmaf_rn(zv[t][1], e4[t].y, dot);
                  dot = __fmaf_rn(zv[t][2], e4[t].z, dot);
                  dot = __fmaf_rn(zv[t][3], e4[t].w, dot);
                }
              }
            }
          }
          dot = __fadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, 16));     // A + B
          const float sc = ref_score(dot, __ldg(P.e2 + korig), z2);
          const unsigned long long kcur = ((unsigned long long)f32_orderable(sc) << 32) |
                                          ((unsigned long long)(0xFFFFu - (uint32_t)korig) << 16) | (unsigned long long)k;
          if (act && kcur > key) key = kcur;
          rem -= act ? 1 : 0;
        } while (__any_sync(0xffffffffu, rem > 0));
        if (!fb && total > 1) worig = 0xFFFF - (int)((key >> 16) & 0xFFFFull);
      }

      TC_TICK(4);
      // ---- outputs ------------------------------------------------------------------------------------
      // q does not leave through the load/store unit: with scalar q stores the LSU data pipe was 91 % busy (ncu,
      // profiles/r01e_*), two fifths of it those stores (a global store costs the pipe two passes per 32-byte sector).
      // Each warp stages the [16 pixel x 32 channel] box of one loop iteration in shared memory (one st.shared per
      // value) and one lane hands it to the TMA (cp.async.bulk.tensor store).  Pixels left to the exhaustive fallback
      // leave stale bytes in the box; the fallback kernel runs afterwards and rewrites them.
      // Ablation hooks (tools/build_variant.sh abl -DVQ_ABL_TCS_NOOUT | _NOZ | _NOE): the phase without its loop / its z
      // loads / its code-row gathers, to time what the rest of the kernel costs.
      if (hf == 0) {
        if (fb) {
          const int slot = atomicAdd(P.fb_count, 1);
          P.fb_rows[slot] = b * P.HW + pp;
        } else {
          int h, wc;
          if (P.w_shift >= 0) { h = pp >> P.w_shift; wc = pp & (P.W - 1); }
          else { h = pp / P.W; wc = pp - h * P.W; }
          const size_t nb_ = (size_t)b * hw;
          if (P.ids) store_id(P.ids + nb_, pp, h, wc, P.H, (int)worig, P.ids_mode);
          if (P.ids_nat) P.ids_nat[nb_ + pp] = worig;
          if (STATS) atomicAdd(&P.counts[worig], 1);
        }
      }
      {
        const bool live = !fb;
        const bool qout = !DBG || P.q;
        const float4* er = reinterpret_cast<const float4*>(P.E + (size_t)(live ? worig : 0) * D);
        float* so = STATS ? sums_mine + (size_t)(live ? worig : 0) * D : nullptr;
        const uint32_t qs_lane = qst_s + (uint32_t)hf * (4 * 64) + (uint32_t)px * 4;      // + t * 512 + i * 64
#ifdef VQ_ABL_TCS_NOOUT
        for (int jc = 0; jc < 0; jc += 8) {
#else
        for (int jc = 0; jc < nq; jc += 8) {              // one 32-channel chunk per iteration (warp-uniform trip count:
#endif
          const int j0 = jc + hf;                         // the loop synchronises the warp); my quads j = j0 + 2t
          float4 e4[4];
          float zv[4][4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int j = j0 + 2 * t;
            const bool in = live && j < nq;
#ifdef VQ_ABL_TCS_NOE
            e4[t] = make_float4(1.f, 2.f, 3.f, (float)j);
#else
            e4[t] = in ? __ldg(er + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
            const float* zj = zp + (size_t)(4 * j) * hw;
#ifdef VQ_ABL_TCS_NOZ
            zv[t][0] = zv[t][1] = zv[t][2] = zv[t][3] = (float)j;
#else
            zv[t][0] = in ? __ldg(zj) : 0.f;          zv[t][1] = in ? __ldg(zj + hw) : 0.f;
            zv[t][2] = in ? __ldg(zj + 2 * hw) : 0.f; zv[t][3] = in ? __ldg(zj + 3 * hw) : 0.f;
#endif
          }
          if (qout) {                                     // the TMA has read the previous box out of the staging buffer
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int j = j0 + 2 * t;
            if (live && j < nq) {
              const float2 m1 = make_float2(-1.f, -1.f);
              const float2 d01 = __ffma2_rn(make_float2(e4[t].x, e4[t].y), m1, make_float2(zv[t][0], zv[t][1]));
              const float2 d23 = __ffma2_rn(make_float2(e4[t].z, e4[t].w), m1, make_float2(zv[t][2], zv[t][3]));
              ls2 = __ffma2_rn(d01, d01, ls2);
              ls2 = __ffma2_rn(d23, d23, ls2);
              if (qout) {                                 // box row = channel within the chunk: 4 (2t + hf) + i, 64 bytes per row
                const uint32_t qa = qs_lane + (uint32_t)t * 512;
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(qa), "f"(e4[t].x) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(qa + 64), "f"(e4[t].y) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(qa + 128), "f"(e4[t].z) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(qa + 192), "f"(e4[t].w) : "memory");
              }
              if (STATS) atomicAdd(reinterpret_cast<float4*>(so + 4 * j), make_float4(zv[t][0], zv[t][1], zv[t][2], zv[t][3]));
            }
          }
          if (qout) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&qmap, qst_s, p0 + quad * 32 + cg * TC_OPX, (j0 >> 3) * TC_DCH, b);
              bulk_commit();
            }
          }
        }
      }
    }
    if (lane == 0) bulk_wait0();                          // every q box has landed before the CTA retires
    TC_TICK(5);
    TC_TIMING_STORE(ew, my_tiles);
    float lsum = ls2.x + ls2.y;
    lsum = warp_sum(lsum);
    if (lane == 0 && P.loss_acc && lsum != 0.f) atomicAdd(P.loss_acc, (double)lsum);
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();               // neither CTA may leave (or free TMEM) while the pair's MMAs can touch it
  else __syncthreads();
  if (warp == 0) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// cuTensorMapEncodeTiled costs a microsecond or two of host time and a forward builds up to four maps: the maps of
// the last calls are kept (per host thread), keyed on everything that goes into them -- a training loop that reuses its
// buffers (or replays the caching allocator's blocks) encodes nothing after the first steps.
static CUresult encode_cached(EncodeTiledFn enc, CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* ptr,
                              const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                              const cuuint32_t* es, CUtensorMapInterleave il, CUtensorMapSwizzle sw,
                              CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
  struct Key {
    void* ptr; int dev; uint32_t rank, sw, l2;
    uint64_t dims[3], strides[2];
    uint32_t box[3];
  };
  struct Entry { Key k; CUtensorMap m; bool used; };
  constexpr int N = 16;
  static thread_local Entry cache[N];
  static thread_local int next = 0;
  Key k;
  memset(&k, 0, sizeof(k));
  k.ptr = ptr; k.dev = current_device(); k.rank = rank; k.sw = (uint32_t)sw; k.l2 = (uint32_t)l2;
  for (cuuint32_t i = 0; i < rank && i < 3; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
  for (cuuint32_t i = 0; i + 1 < rank && i < 2; ++i) k.strides[i] = strides[i];
  for (int i = 0; i < N; ++i)
    if (cache[i].used && memcmp(&cache[i].k, &k, sizeof(k)) == 0) { *out = cache[i].m; return CUDA_SUCCESS; }
  const CUresult r = enc(out, dt, rank, ptr, dims, strides, box, es, il, sw, l2, oob);
  if (r == CUDA_SUCCESS) {
    cache[next].k = k; cache[next].m = *out; cache[next].used = true;
    next = (next + 1) % N;
  }
  return r;
}

static int sm_count_tc() { return device_sm_count(); }

static bool tcs_supported(int D, int K) { return tcs_geometry(D, K).ok; }

static int launch_assign_tcs_impl(const FwdArgs& a, float* dbg, cudaStream_t s);

int launch_assign_tc_impl(const FwdArgs& a, float* dbg, cudaStream_t s) {
  const int HW = a.H * a.W;
  const TcGeom g = tc_geometry(a.D, a.K);
  // resident codebook only when two z stages fit beside it (one stage = no prefetch: D = 256 at K = 64 ran 1.5x slower
  // than the streamed kernel)
  if (!(g.ok && g.nst == 2 && g.nb * g.BN <= TC_SORT_MAX) && tcs_supported(a.D, a.K)) return launch_assign_tcs_impl(a, dbg, s);
  VQ_REQUIRE(g.ok && g.nb * g.BN <= TC_SORT_MAX && HW % TC_TILE == 0 && a.D % 4 == 0, VQ_ERR_UNSUPPORTED, "tensor-core path: unsupported shape");
  VQ_REQUIRE(a.q != nullptr || dbg != nullptr, VQ_ERR_INVALID_ARG, "tensor-core path: q must not be null");
  EncodeTiledFn enc = get_encode_fn();
  VQ_REQUIRE(enc != nullptr, VQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  VQ_REQUIRE((((uintptr_t)a.z) & 15) == 0 && (((uintptr_t)a.embed) & 15) == 0, VQ_ERR_INVALID_ARG,
             "tensor-core path: z / embed must be 16-byte aligned");

  CUtensorMap zmap, emap;
  {   // z [B][D][HW]: dim0 = pixel (contiguous), dim1 = channel, dim2 = image; box = 32 pixels x 32 channels
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)a.D, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)a.D * HW * 4};
    cuuint32_t box[3] = {32, TC_DCH, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_cached(enc, &zmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a.z, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQ_REQUIRE(r == CUDA_SUCCESS, VQ_ERR_CUDA, "cuTensorMapEncodeTiled(z) failed: %d", (int)r);
  }
  {   // norm-sorted codebook copy [nb*BN][D] row-major: dim0 = channel, dim1 = sorted code position
    cuuint64_t dims[2] = {(cuuint64_t)a.D, (cuuint64_t)(g.nb * g.BN)};
    cuuint64_t strides[1] = {(cuuint64_t)a.D * 4};
    cuuint32_t box[2] = {TC_DCH, (cuuint32_t)g.BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_cached(enc, &emap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.ws.tc_es, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQ_REQUIRE(r == CUDA_SUCCESS, VQ_ERR_CUDA, "cuTensorMapEncodeTiled(embed) failed: %d", (int)r);
  }

  float* eaug_img = a.ws.tc_aug;
  uint32_t* meta = reinterpret_cast<uint32_t*>(a.ws.tc_meta);
  uint32_t* rmax = reinterpret_cast<uint32_t*>(a.ws.tc_ctab);
  const int ktot = g.nb * g.BN;
  vq_tc_prep2_kernel<<<(ktot + TC_PREP2_THREADS - 1) / TC_PREP2_THREADS, TC_PREP2_THREADS, (size_t)a.K * 4, s>>>(
      a.embed, a.ws.e2, a.K, a.D, g.BN, g.nb, a.ws.tc_es, eaug_img, a.ws.tc_perm, rmax, meta);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());

  TcParams P{};
  P.z = a.z; P.E = a.embed; P.e2 = a.ws.e2; P.eaug_img = eaug_img; P.meta = meta;
  P.perm = a.ws.tc_perm; P.rmax = rmax;
  P.B = a.B; P.D = a.D; P.H = a.H; P.W = a.W; P.HW = HW; P.K = a.K;
  P.BN = g.BN; P.nb = g.nb; P.nD = g.nD; P.nst = g.nst;
  P.bn_shift = 0;
  while ((1 << P.bn_shift) < g.BN) ++P.bn_shift;
  P.w_shift = -1;
  if ((a.W & (a.W - 1)) == 0) { P.w_shift = 0; while ((1 << P.w_shift) < a.W) ++P.w_shift; }
  P.tiles_per_img = HW / TC_TILE;
  P.ntiles = a.B * P.tiles_per_img;
  P.off_emain = (uint32_t)g.off_emain; P.off_eaug = (uint32_t)g.off_eaug; P.off_aaug = (uint32_t)g.off_aaug;
  P.off_z = (uint32_t)g.off_z; P.off_pub = (uint32_t)g.off_pub;
  P.off_wl = (uint32_t)g.off_wl; P.off_zn = (uint32_t)g.off_zn;
  P.off_hist = (uint32_t)g.off_hist; P.off_perm = (uint32_t)g.off_perm; P.off_ctab = (uint32_t)g.off_ctab;
  P.off_bar = (uint32_t)g.off_bar;
  P.ids = a.ids; P.ids_nat = a.ids_nat; P.q = a.q; P.loss_acc = a.ws.loss_acc;
  P.counts = a.stats ? a.ws.counts : nullptr;
  P.sums = a.stats ? a.stats + stats_sums_offset(a.K) : nullptr;
  P.sums_rep = a.ws.sums_rep;
  P.nrep = tc_sums_replicas(a.K, a.D);
  P.fb_count = a.ws.misc; P.fb_rows = a.ws.fb_rows;
  P.ids_mode = ids_mode_of(a.flags);
  P.dbg = dbg;

  int grid = sm_count_tc();
  if (grid > P.ntiles) grid = P.ntiles;
  const bool stats = a.stats != nullptr;
  typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const TcParams);
  KernFn kern;
  int ki;
  if (dbg) { kern = stats ? vq_assign_tc_kernel<true, true, 0> : vq_assign_tc_kernel<true, false, 0>; ki = stats ? 1 : 0; }
  else if (a.D == 64) { kern = stats ? vq_assign_tc_kernel<false, true, 64> : vq_assign_tc_kernel<false, false, 64>; ki = stats ? 3 : 2; }
  else { kern = stats ? vq_assign_tc_kernel<false, true, 0> : vq_assign_tc_kernel<false, false, 0>; ki = stats ? 5 : 4; }
  static bool attr_set[kMaxDevices][6] = {};               // the attribute is per device
  const int dev = current_device();
  if (!attr_set[dev][ki]) {
    VQ_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    attr_set[dev][ki] = true;
  }
  const bool prof = profile_begin(s);
  kern<<<grid, TC_THREADS, g.total, s>>>(zmap, emap, P);
  if (prof) profile_end(s);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

// CTA-pair mode of the streamed kernel (256-wide code blocks, at least two tiles).  Opt-in: VQ_FLAG_PAIR, or VQ_TCS_PAIR=1
// in the environment for A/B timing -- measured on B200 it is bit-identical to and no faster than one CTA per tile
// (DESIGN.md 4.2), because the streamed kernel is not bound by the codebook's L2 -> SM traffic.
static bool tcs_use_pair(const FwdArgs& a) {
  static int env_on = -1;
  if (env_on < 0) { const char* e = getenv("VQ_TCS_PAIR"); env_on = (e && e[0] == '1') ? 1 : 0; }
  if (!env_on && !(a.flags & VQ_FLAG_PAIR)) return false;
  const long long ntiles = (long long)a.B * (a.H * a.W / TC_TILE);
  return ntiles >= 2 && sm_count_tc() >= 2 && tcs_geometry(a.D, a.K, true).ok;
}

static int launch_assign_tcs_impl(const FwdArgs& a, float* dbg, cudaStream_t s) {
  const int HW = a.H * a.W;
  const bool pair = tcs_use_pair(a);
  const TcsGeom g = tcs_geometry(a.D, a.K, pair);
  VQ_REQUIRE(g.ok && HW % TC_TILE == 0 && a.D % 4 == 0, VQ_ERR_UNSUPPORTED, "tensor-core path (streamed codebook): unsupported shape");
  VQ_REQUIRE(a.q != nullptr || dbg != nullptr, VQ_ERR_INVALID_ARG, "tensor-core path: q must not be null");
  EncodeTiledFn enc = get_encode_fn();
  VQ_REQUIRE(enc != nullptr, VQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  VQ_REQUIRE((((uintptr_t)a.z) & 15) == 0 && (((uintptr_t)a.embed) & 15) == 0, VQ_ERR_INVALID_ARG,
             "tensor-core path: z / embed must be 16-byte aligned");
  CUtensorMap zmap, emap, amap, qmap;
  float* eaug_img = a.ws.tc_aug;
  {
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)a.D, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)a.D * HW * 4};
    cuuint32_t box[3] = {32, TC_DCH, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_cached(enc, &zmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a.z, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQ_REQUIRE(r == CUDA_SUCCESS, VQ_ERR_CUDA, "cuTensorMapEncodeTiled(z) failed: %d", (int)r);
  }
  {   // pair mode: a box is this CTA's half of a 256-code slice
    cuuint64_t dims[2] = {(cuuint64_t)a.D, (cuuint64_t)(g.nb * g.BN)};
    cuuint64_t strides[1] = {(cuuint64_t)a.D * 4};
    cuuint32_t box[2] = {TC_DCH, (cuuint32_t)(pair ? g.BN / 2 : g.BN)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_cached(enc, &emap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.ws.tc_es, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQ_REQUIRE(r == CUDA_SUCCESS, VQ_ERR_CUDA, "cuTensorMapEncodeTiled(embed) failed: %d", (int)r);
  }
  if (a.q) {   // q [B][D][HW] like z; a box is one epilogue warp's 16 pixels x one 32-channel chunk (dense rows of 64 bytes)
    VQ_REQUIRE((((uintptr_t)a.q) & 15) == 0, VQ_ERR_INVALID_ARG, "tensor-core path: q must be 16-byte aligned");
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)a.D, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)a.D * HW * 4};
    cuuint32_t box[3] = {16, TC_DCH, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_cached(enc, &qmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a.q, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQ_REQUIRE(r == CUDA_SUCCESS, VQ_ERR_CUDA, "cuTensorMapEncodeTiled(q) failed: %d", (int)r);
  } else {
    memset(&qmap, 0, sizeof(qmap));
  }
  {   // augmentation image as rows of one 8-code group (64 floats, already in the UMMA no-swizzle layout): only the pair
      // kernel loads it through a tensor map (a tensor load may signal the barrier of the peer CTA)
    cuuint64_t dims[2] = {64, (cuuint64_t)(g.nb * g.BN / 8)};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, (cuuint32_t)(pair ? g.BN / 16 : g.BN / 8)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_cached(enc, &amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)eaug_img, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQ_REQUIRE(r == CUDA_SUCCESS, VQ_ERR_CUDA, "cuTensorMapEncodeTiled(augmentation) failed: %d", (int)r);
  }
  uint32_t* meta = reinterpret_cast<uint32_t*>(a.ws.tc_meta);
  uint32_t* rmax = reinterpret_cast<uint32_t*>(a.ws.tc_ctab);
  const int ktot = g.nb * g.BN;
  vq_tc_prep2_kernel<<<(ktot + TC_PREP2_THREADS - 1) / TC_PREP2_THREADS, TC_PREP2_THREADS, (size_t)a.K * 4, s>>>(
      a.embed, a.ws.e2, a.K, a.D, g.BN, g.nb, a.ws.tc_es, eaug_img, a.ws.tc_perm, rmax, meta);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());

  TcsParams P{};
  P.z = a.z; P.E = a.embed; P.e2 = a.ws.e2; P.eaug_img = eaug_img; P.meta = meta;
  P.perm = a.ws.tc_perm; P.rmax = rmax;
  P.B = a.B; P.D = a.D; P.H = a.H; P.W = a.W; P.HW = HW; P.K = a.K;
  P.BN = g.BN; P.nb = g.nb; P.nD = g.nD; P.nst = g.nst;
  P.bn_shift = 0;
  while ((1 << P.bn_shift) < g.BN) ++P.bn_shift;
  P.w_shift = -1;
  if ((a.W & (a.W - 1)) == 0) { P.w_shift = 0; while ((1 << P.w_shift) < a.W) ++P.w_shift; }
  P.tiles_per_img = HW / TC_TILE;
  P.ntiles = a.B * P.tiles_per_img;
  P.stage_bytes = (uint32_t)g.stage_bytes; P.off_stage = (uint32_t)g.off_stage;
  P.off_aug = (uint32_t)g.off_aug; P.off_aaug = (uint32_t)g.off_aaug;
  P.off_pub = (uint32_t)g.off_pub; P.off_zn = (uint32_t)g.off_zn;
  P.off_ctab = (uint32_t)g.off_ctab; P.off_qst = (uint32_t)g.off_qst; P.off_bar = (uint32_t)g.off_bar;
  P.ids = a.ids; P.ids_nat = a.ids_nat; P.q = a.q; P.loss_acc = a.ws.loss_acc;
  P.counts = a.stats ? a.ws.counts : nullptr;
  P.sums = a.stats ? a.stats + stats_sums_offset(a.K) : nullptr;
  P.sums_rep = a.ws.sums_rep;
  P.nrep = tc_sums_replicas(a.K, a.D);
  P.fb_count = a.ws.misc; P.fb_rows = a.ws.fb_rows;
  P.ids_mode = ids_mode_of(a.flags);
  P.dbg = dbg;

  int grid = sm_count_tc();
  if (pair) {                                  // one 2-CTA cluster per tile pair, at most one CTA per SM
    const int npairs = (P.ntiles + 1) / 2;
    int nclusters = grid / 2;
    if (nclusters > npairs) nclusters = npairs;
    grid = 2 * nclusters;
  } else if (grid > P.ntiles) {
    grid = P.ntiles;
  }
  const bool stats = a.stats != nullptr;
  typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcsParams);
  static const KernFn kerns[8] = {
      vq_assign_tcs_kernel<false, false, false>, vq_assign_tcs_kernel<false, true, false>,
      vq_assign_tcs_kernel<true, false, false>,  vq_assign_tcs_kernel<true, true, false>,
      vq_assign_tcs_kernel<false, false, true>,  vq_assign_tcs_kernel<false, true, true>,
      vq_assign_tcs_kernel<true, false, true>,   vq_assign_tcs_kernel<true, true, true>};
  const int ki = (pair ? 4 : 0) + (dbg ? 2 : 0) + (stats ? 1 : 0);
  KernFn kern = kerns[ki];
  static bool attr_set[kMaxDevices][8] = {};               // the attribute is per device
  const int dev = current_device();
  if (!attr_set[dev][ki]) {
    VQ_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    attr_set[dev][ki] = true;
  }
  const bool prof = profile_begin(s);
  if (pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = g.total;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VQ_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, zmap, emap, amap, qmap, P));
  } else {
    kern<<<grid, TC_THREADS, g.total, s>>>(zmap, emap, amap, qmap, P);
  }
  if (prof) profile_end(s);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

int launch_assign_tc(const FwdArgs& a, cudaStream_t s) { return launch_assign_tc_impl(a, nullptr, s); }

int tc_debug_timing(long long* host_out, int n) {
#ifdef VQ_TC_TIMING
  const int tot = 148 * 16 * 8;
  if (n < tot) return -1;
  if (cudaMemcpyFromSymbol(host_out, g_tc_timing, sizeof(long long) * tot) != cudaSuccess) return -2;
  return tot;
#else
  (void)host_out; (void)n;
  return 0;                      // role timing is not compiled into this build
#endif
}

int tc_debug_ncols(int D, int K) {
  const TcGeom g = tc_geometry(D, K);
  if (g.ok && g.nb * g.BN <= TC_SORT_MAX && (g.nst == 2 || !tcs_geometry(D, K).ok)) return g.nb * g.BN;
  const TcsGeom gs = tcs_geometry(D, K);
  return gs.ok ? gs.nb * gs.BN : 0;
}

}  // namespace vqb200

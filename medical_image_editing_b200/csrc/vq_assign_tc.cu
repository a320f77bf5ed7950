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
  // enough replicas that concurrent CTAs rarely reduce into the same L2 line, capped at 16 MB in total (the replicas
  // are zeroed by vq_prep_kernel and folded by vq_finish_kernel on every call: a few microseconds; with 4 MB -- 8
  // replicas at K = 512, D = 256 -- a skewed code usage (post-ReLU input) serialised the reductions of the hot rows in L2)
  size_t r = ((size_t)16 << 20) / ((size_t)K * D * 4);
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
// Rank by counting: eight threads own one sorted-order key (norm bits, original index); each counts the smaller keys in
// an eighth of the shared-memory copy, three shuffles add the counts -- O(K^2 / threads) compares, no sorting network,
// any number of CTAs (32 codes per CTA: 16 CTAs at K = 512; one thread per code and 2 CTAs took 14 us of mostly
// dependent shared-memory latency).
constexpr int TC_PREP2_THREADS = 256;
constexpr int TC_PREP2_CODES = TC_PREP2_THREADS / 8;       // codes per CTA
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
  // eight threads per code (TC_PREP2_CODES codes per CTA): each counts an eighth of the keys, then they share the row copy
  const int sub = tid & 7;
  const int i = blockIdx.x * TC_PREP2_CODES + (tid >> 3);        // code (i < K) or padding slot (K <= i < ktot)
  if (i == 0 && sub == 0) meta[0] = __float_as_uint(rcap);
  const bool inside = i < ktot;
  const bool real = i < K;
  int rank = 0;
  uint32_t mine = 0u;
  if (real) {
    mine = keys[i];
    const int per = ((K + 31) / 32) * 4;                         // keys per sub-thread, a multiple of 4
    const int j0 = sub * per;
    const int j1 = min(K, j0 + per);
    int j = j0;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    for (; j + 4 <= j1; j += 4) {                                // four keys per 16-byte load
      const uint4 kk = k4[j >> 2];
      rank += (kk.x < mine || (kk.x == mine && j < i)) ? 1 : 0;
      rank += (kk.y < mine || (kk.y == mine && j + 1 < i)) ? 1 : 0;
      rank += (kk.z < mine || (kk.z == mine && j + 2 < i)) ? 1 : 0;
      rank += (kk.w < mine || (kk.w == mine && j + 3 < i)) ? 1 : 0;
    }
    for (; j < j1; ++j) {
      const uint32_t kj = keys[j];
      rank += (kj < mine || (kj == mine && j < i)) ? 1 : 0;
    }
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  rank += __shfl_xor_sync(0xffffffffu, rank, 4);
  if (inside) {
    const int pos = real ? rank : i;          // padding keeps its slot: K .. ktot-1
    if (sub == 0) {
      float a0 = -1e30f, a1 = 0.f, a2 = 0.f;  // padding / excluded codes never win
      if (real) {
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
    }
    // sorted copy of the codebook row (TMA source): the eight threads of the code take every eighth 16-byte piece
    const int dq = D >> 2;
    float4* dst = reinterpret_cast<float4*>(es + (size_t)pos * D);
    const float4* src = reinterpret_cast<const float4*>(E + (size_t)(real ? i : 0) * D);
    int j = sub;
    for (; j + 24 < dq; j += 32) {                // four 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = real ? __ldg(src + j + 8 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) dst[j + 8 * u] = v[u];
    }
    for (; j < dq; j += 8) dst[j] = real ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
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
              float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (act) e4 = lds_v4(eb + (uint32_t)((2 * t) >> 3) * bn128 + (((j & 7) << 4) ^ r7));    // idle lanes load nothing
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
// resident kernel, third generation (emb_dim = 64): decoupled roles, LSU-lean epilogue
// ---------------------------------------------------------------------------------------------
// The kernel above is bound by the shared-memory / LSU data pipe (ncu: 82 % of the wavefront peak in training mode,
// 60 % in eval; profiles/r02b_*): conflicted 4-byte z copies, code rows re-read by every re-rank pass of every lane, q
// leaving in 64-byte runs, and one long serial chain per tile in the output warps.  This kernel keeps the search
// (tcgen05 tf32 scores in TMEM, rigorous bound, exact fp32 re-rank, exhaustive fallback list) and reorganises the rest:
//   warp 11      TMA producer (as above)                       warp 15  MMA issuer (as above)
//   warp 18      |z|^2: lane = 4 adjacent pixels (16-byte shared-memory loads, conflict-free)
//   warps 0..7   scan: thread = (pixel, column group).  16 accumulator columns per tcgen05.ld, the next load in flight while
//                the current one is reduced; every column enters two max-reductions over orthogonal partitions of the
//                thread's columns -- quarter-chunks (8 consecutive columns) and residues (column index mod 8) -- so a
//                column can only be a candidate if its quarter AND its residue are hot: one instruction per score, no
//                per-column mask.  Bounds + records go to the `pub` slot.
//   8 output warps (8-10, 12-14, 16, 17): two software-pipelined jobs per iteration j:
//                merge(j)      thread = pixel (output warps 0..3 / 4..7 take the even / odd tiles): the two column groups'
//                              records -> one candidate cell: winner into the `win` ring (+ code map, histogram); several:
//                              the pixel goes to the straggler queue and its z row is gathered into the queue entry with
//                              cp.async (nobody waits for that load); overflow / non-finite / exploded-code doubt:
//                              fallback list
//                outputs(j-2)  thread = (4 adjacent pixels, 2 channel quads).  z: 16-byte global loads (L2 hits: the TMA
//                              has just brought the tile in), so the shared-memory stage is held only by the tensor core
//                              and the |z|^2 warp; code rows: two 16-byte loads per pixel; q: 16-byte stores, a quarter-warp
//                              writes one full 128-byte line; (z-q)^2; EMA sums (red.v4, 64 contiguous bytes per pixel and
//                              instruction, equal codes among the four pixels summed first)
//   warp 19      stragglers: exact fp32 re-rank (the library's two-chain dot product, two lanes per pixel, two candidate
//                cells per pass) of the queued pixels from the z rows in the queue; writes the winner into the `win` ring
// The output job waits for a tile's `win` slot to be complete (128 arrivals: decided pixels by the merging warps, the
// rest by the straggler warp), two tiles behind the merge: q / loss / sums have a single writer, no ordering between roles
// is needed, and the straggler warp has two tile periods per batch.
constexpr int R3_D = 64;
constexpr int R3_ND = R3_D / TC_DCH;                 // 2 chunks of 32 channels
constexpr int R3_SCAN_WARPS = 8, R3_OUT_WARPS = 8;
constexpr int R3_WARPS = TC_AUX_WARPS + R3_SCAN_WARPS + R3_OUT_WARPS;      // 20 warps: five per scheduler, 96 registers per thread
constexpr int R3_THREADS = 32 * R3_WARPS;
// warp -> role.  Scan warp w (0..7) reads TMEM lane quadrant w & 3.  A scheduler (warp id mod 4) shares its issue slots
// evenly among its ready warps, so a serial role runs at the pace of its scheduler's crowd: the straggler warp shares
// scheduler 3 with the two light single-thread roles (TMA producer, MMA issuer) instead of two output warps
//   scheduler 0: scan 0, 4, output 8, 12, 16      scheduler 1: scan 1, 5, output 9, 13, 17
//   scheduler 2: scan 2, 6, output 10, 14, |z|^2 18      scheduler 3: scan 3, 7, producer 11, MMA 15, stragglers 19
constexpr int R3_W_PROD = 11, R3_W_MMA = 15, R3_W_ZN = 18, R3_W_STRAG = 19;
// output warp index (0..7) of a hardware warp, or -1
__device__ __forceinline__ int r3_out_index(int warp) {
  switch (warp) {
    case 8: return 0; case 9: return 1; case 10: return 2; case 12: return 3;
    case 13: return 4; case 14: return 5; case 16: return 6; case 17: return 7;
    default: return -1;
  }
}
constexpr int R3_ZSTAGE = R3_ND * TC_TILE * 128;     // 32 KB
constexpr int R3_QCAP = 16;                          // straggler queue entries: deferral rate x (gather latency + batch time), with margin
constexpr int R3_LAG = 3;                            // the output job runs this many tiles behind the merge
constexpr int R3_QSTRIDE = 272;                      // bytes per entry: z row (256) + four records; 68 floats: conflict-free 16-byte reads
constexpr int R3_WINS = 4;                           // win ring slots (outputs run two tiles behind the merge)
constexpr int R3_MAXCAND = 16;                       // candidate cells re-scored exactly per pixel (more: fallback list)
constexpr uint32_t R3_WIN_SKIP = 0x80000000u;        // `win` entry of a pixel on the fallback list

struct R3Geom {
  int BN, nb;
  size_t off_emain, off_eaug, off_aaug, off_z, off_pub, off_win, off_zn, off_hist, off_perm, off_ctab, off_queue, off_bar, total;
  bool ok;
};

static R3Geom r3_geometry(int K) {
  R3Geom g{};
  g.ok = false;
  g.BN = 32;
  while (g.BN < K && g.BN < TC_MAXBN) g.BN <<= 1;
  g.nb = (K + g.BN - 1) / g.BN;
  if (g.nb > 2) return g;                       // 512 TMEM columns = two accumulator blocks
  const size_t ktot = (size_t)g.nb * g.BN;
  size_t off = 0;
  g.off_emain = off; off += ktot * R3_ND * 128;
  g.off_eaug = off;  off += align_up(ktot * 32, 1024);
  g.off_aaug = off;  off += 4096;
  g.off_z = off;     off += 2 * (size_t)R3_ZSTAGE;
  g.off_pub = off;   off += 2 * TC_TILE * 16;            // [column group][pixel] x 16 B (one slot: the merge reads it at once)
  g.off_win = off;   off += R3_WINS * TC_TILE * 4;       // [tile & 3][pixel]
  g.off_zn = off;    off += 3 * TC_TILE * 4;             // [tile % 3][pixel]
  g.off_hist = off;  off += align_up((size_t)((K + 1) / 2) * 4, 16);      // two 16-bit counters per word (a CTA sees < 65536 pixels: checked at launch)
  g.off_perm = off;  off += align_up(ktot * 2, 16);
  g.off_ctab = off;  off += align_up((size_t)g.nb * (g.BN / 32) * 8, 16);
  g.off_queue = off; off += R3_QCAP * R3_QSTRIDE + 24 * 8 + 96 + 32 + 16 * 8 + 16 * 4;   // entries, entry barriers, {iteration << 7 | pixel}, {tail, head, warps done}
  g.off_bar = off;   off += 256;
  g.total = off;            // no slack: the kernel checks that the dynamic shared memory starts 1024-byte aligned
  g.ok = g.total <= (size_t)TC_SMEM_LIMIT;
  return g;
}

__device__ __forceinline__ void mbar_arrive_cnt(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
#ifndef VQ_R3_RELAXED_SLEEP
#define VQ_R3_RELAXED_SLEEP 0
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (VQ_R3_RELAXED_SLEEP) mbar_wait_sleep(bar, parity, VQ_R3_RELAXED_SLEEP);
  else mbar_wait(bar, parity);      // try_wait with a suspend-time hint: the hardware parks the warp on the barrier
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0u;
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {      // arrives on `bar` once this thread's copies have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_s32(int* p, int v) {
  asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld that the compiler cannot move the uses of `r` across (the registers are in/out operands)
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Scan state of one thread (pixel, column group), accumulator units a_k = z.e_k - |e_k|^2/2:
//   L     lower bound on the best exact a_k among the columns seen so far = max_c (chunkmax_c - delta_c)
//   Urec  upper bound on the exact a_k of every recorded candidate
//   recA / recB  records of the (at most two) blocks with hot columns: block << 24 | hot quarter-chunks << 8 | hot residues
//                (bit 15-q of the quarter field: quarter q of the thread's columns in that block; bit 7-j: residue j);
//                cnt > 2: overflow (fallback list)
struct R3Scan {
  float L, Urec;
  int cnt;
  uint32_t recA, recB;
};

// 16 columns (half a chunk): two quarter maxima and the eight residue maxima, 16 three-input max instructions
__device__ __forceinline__ void r3_reduce16(const uint32_t (&b)[16], float& q0, float& q1, float (&R)[8]) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(b[j]);
  q0 = fmaxf(fmax3(fmax3(v[0], v[1], v[2]), fmax3(v[3], v[4], v[5]), v[6]), v[7]);
  q1 = fmaxf(fmax3(fmax3(v[8], v[9], v[10]), fmax3(v[11], v[12], v[13]), v[14]), v[15]);
#pragma unroll
  for (int j = 0; j < 8; ++j) R[j] = fmax3(R[j], v[j], v[j + 8]);
}

// all columns of one accumulator block that belong to this thread: chunks cg, cg + 2, ... (32 columns each), read as
// half-chunks with the next tcgen05.ld in flight.  hm[2 i + h]: quarter maxima of half-chunk i, R: residue maxima
__device__ __forceinline__ void r3_block_maxima(uint32_t taddr, int cg, int nchunks, float (&hm)[16], float (&R)[8]) {
#pragma unroll
  for (int q = 0; q < 16; ++q) hm[q] = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) R[j] = -INFINITY;
  uint32_t b0[16], b1[16];
  if (nchunks == 8) {                                   // (warp-uniform) 256-code block: four chunks per thread, no guards
    const uint32_t t0 = taddr + (uint32_t)(32 * cg);
    tmem_ld16(t0, b0);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      tmem_ld_wait16(b0);
      tmem_ld16(t0 + (uint32_t)(64 * (i >> 1) + 16), b1);
      r3_reduce16(b0, hm[2 * i], hm[2 * i + 1], R);
      tmem_ld_wait16(b1);
      if (i + 2 < 8) tmem_ld16(t0 + (uint32_t)(64 * ((i + 2) >> 1)), b0);
      r3_reduce16(b1, hm[2 * i + 2], hm[2 * i + 3], R);
    }
    return;
  }
  const int nmine = (nchunks - cg + 1) >> 1;            // chunks of this thread (warp-uniform), <= 4
  const int nh = 2 * nmine;                             // half-chunks
  // half-chunk i: chunk cg + 2 (i >> 1), half i & 1 -> first column 32 cg + 64 (i >> 1) + 16 (i & 1)
  tmem_ld16(taddr + (uint32_t)(32 * cg), b0);
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    if (i < nh) {
      tmem_ld_wait16(b0);
      tmem_ld16(taddr + (uint32_t)(32 * cg + 64 * (i >> 1) + 16), b1);
      r3_reduce16(b0, hm[2 * i], hm[2 * i + 1], R);
      tmem_ld_wait16(b1);
      if (i + 2 < nh) tmem_ld16(taddr + (uint32_t)(32 * cg + 64 * ((i + 2) >> 1)), b0);
      r3_reduce16(b1, hm[2 * i + 2], hm[2 * i + 3], R);
    }
  }
}

// end of a block: bounds, hot quarters, hot residues, record
__device__ __forceinline__ void r3_block_end(R3Scan& st, const float (&hm)[16], const float (&R)[8], int cg, int nchunks,
                                             int blk, uint32_t ctab_s, float zn) {
  float dl[4];
  float bm = -INFINITY, dmax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cg + 2 * i;
    dl[i] = 0.f;
    if (c < nchunks) {                                    // warp-uniform
      const float2 ab = lds_v2(ctab_s + (uint32_t)(blk * nchunks + c) * 8);
      dl[i] = __fmaf_rn(zn, ab.x, ab.y);
      const float m = fmaxf(fmaxf(hm[4 * i], hm[4 * i + 1]), fmaxf(hm[4 * i + 2], hm[4 * i + 3]));
      st.L = fmaxf(st.L, m - dl[i]);
      bm = fmaxf(bm, m);
      dmax = fmaxf(dmax, dl[i]);
    }
  }
  if (st.Urec < st.L) { st.cnt = 0; st.Urec = -INFINITY; }           // nothing recorded so far can still win
  // "below threshold" sign bits, gathered by short independent funnel-shift chains (one per chunk / residue half)
  uint32_t nq4[4], nr2[2] = {0u, 0u};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float T = st.L - dl[i];
    const float2 nT2 = make_float2(-T, -T);
    const float2 dA = __fadd2_rn(make_float2(hm[4 * i], hm[4 * i + 1]), nT2);
    const float2 dB = __fadd2_rn(make_float2(hm[4 * i + 2], hm[4 * i + 3]), nT2);
    uint32_t n = __float_as_uint(dA.x) >> 31;
    n = __funnelshift_l(__float_as_uint(dA.y), n, 1);
    n = __funnelshift_l(__float_as_uint(dB.x), n, 1);
    nq4[i] = __funnelshift_l(__float_as_uint(dB.y), n, 1);
  }
  {
    const float Tr = st.L - dmax;
    const float2 nT2 = make_float2(-Tr, -Tr);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float2 d2 = __fadd2_rn(make_float2(R[j], R[j + 1]), nT2);
      nr2[j >> 2] = __funnelshift_l(__float_as_uint(d2.x), nr2[j >> 2], 1);
      nr2[j >> 2] = __funnelshift_l(__float_as_uint(d2.y), nr2[j >> 2], 1);
    }
  }
  const uint32_t nq = (nq4[0] << 12) | (nq4[1] << 8) | (nq4[2] << 4) | nq4[3];
  const uint32_t nr = (nr2[0] << 4) | nr2[1];
  const uint32_t qm = ~nq & 0xFFFFu;                      // bit (15-q): quarter q is hot
  const uint32_t rm = ~nr & 0xFFu;                        // bit (7-j): residue j is hot
  const uint32_t rec = ((uint32_t)blk << 24) | (qm << 8) | rm;
  const bool has = qm != 0u;
  const bool s0 = has && st.cnt == 0, s1 = has && st.cnt == 1;
  st.recA = s0 ? rec : st.recA;
  st.recB = s1 ? rec : st.recB;
  st.cnt += has ? 1 : 0;
  st.Urec = has ? fmaxf(st.Urec, bm + dmax) : st.Urec;
}
// number of candidate cells of a record
__device__ __forceinline__ int r3_cells(uint32_t rec) { return __popc(rec & 0x00FFFF00u) * __popc(rec & 0xFFu); }
// sorted-codebook position of cell (quarter q, residue j) of a record of column group cg
__device__ __forceinline__ int r3_cell_pos(uint32_t rec, int cg, int q, int j, int bnsh) {
  return (int)((rec >> 24) << bnsh) + (cg + 2 * (q >> 2)) * 32 + (q & 3) * 8 + j;
}

// Optional event trace (build with -DVQ_R3_TRACE; read with vq_debug_tc_timing / tools/r3_trace.py): clock64 of CTA 0's first
// 64 tiles, [event][tile]
#ifdef VQ_R3_TRACE
__device__ long long g_r3_trace[32 * 64];
#define R3_EV(e, it_)                                                                                  \
  do {                                                                                                 \
    if (blockIdx.x == 0 && lane == 0 && (it_) < 64) g_r3_trace[(e) * 64 + (it_)] = clock64();         \
  } while (0)
#else
#define R3_EV(e, it_) do { } while (0)
#endif

struct R3Params {
  TcParams t;
  uint32_t off_win, off_queue;
  uint32_t zero;          // always 0 (see r3_after_loads)
};
// A shared-memory slot is handed back (mbarrier.arrive) right after it was read with plain loads whose values are only
// used later.  The arrive does not wait for those loads: with the LSU queue backed up (the EMA reductions of training mode)
// the next writer overwrote the slot before the loads had read it (measured: wrong (z-q)^2 in training mode, never in
// eval).  Folding the loaded registers into the barrier address -- through a mask the compiler cannot see is zero --
// makes the arrive wait on the loads' scoreboard.
__device__ __forceinline__ uint32_t r3_after_loads(uint32_t bar, uint32_t loaded_bits, uint32_t zero) { return bar + (loaded_bits & zero); }

// Debug build (-DVQ_R3_CHECK): every wait gives up after a while, records {site, tile, CTA} of the first stuck waiter in
// misc[8..] of the workspace... (g_r3_dbg, read with vq_debug_tc_timing) and lets the kernel run to its end
#ifdef VQ_R3_CHECK
__device__ int g_r3_abort;
__device__ long long g_r3_dbg[32 * 64];
__device__ __forceinline__ void r3_dbg_wait(uint32_t bar, uint32_t parity, int site, int it) {
  for (long long spins = 0;; ++spins) {
    if (mbar_test(bar, parity)) return;
    if (*(volatile int*)&g_r3_abort) return;
    if (spins > 4000000) {
      if (atomicExch(&g_r3_abort, 1) == 0) { g_r3_dbg[0] = site; g_r3_dbg[1] = it; g_r3_dbg[2] = blockIdx.x; g_r3_dbg[3] = threadIdx.x; }
      const int slot = 8 + 4 * (int)(threadIdx.x >> 5);
      if ((threadIdx.x & 31) == 0) { g_r3_dbg[slot] = site; g_r3_dbg[slot + 1] = it; g_r3_dbg[slot + 2] = blockIdx.x; }
      return;
    }
    __nanosleep(20);
  }
}
#define R3_WAIT(bar, par, site, it) r3_dbg_wait(bar, par, site, it)
#define R3_WAIT_SLEEP(bar, par, ns, site, it) r3_dbg_wait(bar, par, site, it)
#define R3_WAIT_RELAXED(bar, par, site, it) r3_dbg_wait(bar, par, site, it)
#else
#define R3_WAIT(bar, par, site, it) mbar_wait(bar, par)
#define R3_WAIT_SLEEP(bar, par, ns, site, it) mbar_wait_sleep(bar, par, ns)
#define R3_WAIT_RELAXED(bar, par, site, it) mbar_wait_relaxed(bar, par)
#endif

// next candidate cell of a queued pixel: records c[0..3] (column groups 0, 0, 1, 1), iterator (ri, qm, rm)
struct R3Cells {
  uint32_t c0, c1, c2, c3, cur, qm, rm;
  int ri;
  __device__ __forceinline__ void init(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    c0 = a; c1 = b; c2 = c; c3 = d; ri = 0; cur = a; qm = (a >> 8) & 0xFFFFu; rm = a & 0xFFu;
  }
  __device__ __forceinline__ int next(int bnsh) {         // precondition: a cell is left
    while (qm == 0u || rm == 0u) {
      ++ri;
      cur = ri == 1 ? c1 : ri == 2 ? c2 : c3;
      qm = (cur >> 8) & 0xFFFFu; rm = cur & 0xFFu;
      if (ri >= 3) break;
    }
    const int q = __clz(qm << 16), j = __clz(rm << 24);
    const int k = r3_cell_pos(cur, ri >> 1, q, j, bnsh);
    rm &= ~(0x80u >> j);
    if (rm == 0u) { qm &= ~(0x8000u >> q); rm = qm ? (cur & 0xFFu) : 0u; }
    return k;
  }
};

template <bool STATS>
__global__ void __launch_bounds__(R3_THREADS, 1)
vq_assign_r3_kernel(const __grid_constant__ CUtensorMap zmap, const __grid_constant__ CUtensorMap emap, const R3Params PP) {
  const TcParams& P = PP.t;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();                    // swizzled TMA / UMMA tiles need 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint64_t* bars = (uint64_t*)(smem + P.off_bar);
  const uint32_t bar0 = sbase + P.off_bar;
  // barrier slots: 0 e_full | 1,2 z_full | 3,4 z_empty | 5,6 tmem_full | 7,8 tmem_empty | 9..11 zn_full |
  //                13..16 win_full | 17..20 win_empty | 21,22 pub_full | 24,25 pub_empty ; slot 23: tmem base
  // (the pub slot is single, but each merging team waits on its own barrier pair: a waiter that skipped every other phase
  // of a shared barrier would alias with the phase before -- parity waits may be at most one phase ahead)
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + 23);
  uint16_t* perm_s = (uint16_t*)(smem + P.off_perm);
  // straggler queue: entries [R3_QCAP] x {z row, four records}, one mbarrier per entry (32 cp.async arrivals), the entries'
  // {iteration << 7 | pixel} words, {tail, head, warps done}
  const uint32_t queue_s = sbase + PP.off_queue;
  const uint32_t qbar_s = queue_s + R3_QCAP * R3_QSTRIDE;
  const uint32_t qitp_s = qbar_s + 24 * 8;
  volatile uint32_t* qctl = (volatile uint32_t*)(smem + PP.off_queue + R3_QCAP * R3_QSTRIDE + 24 * 8 + 96);
  const uint32_t qkey_s = qitp_s + 96 + 32;             // straggler scratch: best key per batch entry [16] x 8 B, |z|^2 per entry [16] x 4 B
  const uint32_t qz2_s = qkey_s + 16 * 8;
  uint32_t* hist = (uint32_t*)(smem + P.off_hist);
  const uint32_t win_s = sbase + PP.off_win;

  constexpr int nD = R3_ND;
  const int ktot = P.nb * P.BN;
  const int bnsh = P.bn_shift;
  const uint32_t bn128 = (uint32_t)P.BN * 128;

  if (threadIdx.x == 32) {
    mbar_init(BAR(0), 1);
    mbar_init(BAR(1), 1); mbar_init(BAR(2), 1);
    mbar_init(BAR(3), 2); mbar_init(BAR(4), 2);           // the |z|^2 warp and the MMA commit
    mbar_init(BAR(5), 1); mbar_init(BAR(6), 1);
    mbar_init(BAR(7), R3_SCAN_WARPS); mbar_init(BAR(8), R3_SCAN_WARPS);
    for (int i = 0; i < 3; ++i) mbar_init(BAR(9 + i), 1);
    for (int i = 0; i < R3_WINS; ++i) { mbar_init(BAR(13 + i), TC_TILE); mbar_init(BAR(17 + i), R3_OUT_WARPS); }
    mbar_init(BAR(21), R3_SCAN_WARPS); mbar_init(BAR(22), R3_SCAN_WARPS);
    mbar_init(BAR(24), R3_OUT_WARPS / 2); mbar_init(BAR(25), R3_OUT_WARPS / 2);
    for (int i = 0; i < R3_QCAP; ++i) mbar_init(qbar_s + 8u * i, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < R3_SCAN_WARPS) {
    const int t = threadIdx.x;
    constexpr int NT = 32 * R3_SCAN_WARPS;
    float4* a = (float4*)(smem + P.off_aaug);
    for (int i = t; i < 256; i += NT) {                   // ones block of the augmentation K-step (see the kernel above)
      const int row = (i >> 3) & 7;
      const float v = row < 3 ? 1.f : 0.f;
      a[i] = make_float4(v, v, v, v);
    }
    for (int k = t; k < ktot; k += NT) perm_s[k] = (uint16_t)P.perm[k];
    {
      const float c1 = 0.001953125f * 1.03f;
      const float c2 = (float)(R3_D + 16) * 4.76837158e-7f;
      for (int c = t; c < P.nb * (P.BN >> 5); c += NT) {
        const float rm = __uint_as_float(P.rmax[c]);
        ((float*)(smem + P.off_ctab))[2 * c] = 0.5f * (c1 + c2) * rm;
        ((float*)(smem + P.off_ctab))[2 * c + 1] = 0.5f * c2 * rm * rm + 1e-30f;
      }
    }
    if (t < 8) qctl[t] = 0u;
    if (STATS) for (int i = t; i < (P.K + 1) / 2; i += NT) hist[i] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == R3_W_PROD) {
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
        const int s = it & 1, ph = (it >> 1) & 1;
        R3_WAIT_SLEEP(BAR(3 + s), ph ^ 1, 64, 0, it);
        R3_EV(0, it);
        mbar_expect_tx(BAR(1 + s), R3_ZSTAGE);
        const int b = tile / P.tiles_per_img, pt = tile % P.tiles_per_img;
        for (int c = 0; c < nD; ++c)
          for (int grp = 0; grp < 4; ++grp)
            tma_load_3d(sbase + P.off_z + s * R3_ZSTAGE + c * (TC_TILE * 128) + grp * 4096, &zmap, BAR(1 + s),
                        pt * TC_TILE + grp * 32, c * TC_DCH, b);
      }
    }
  } else if (warp == R3_W_MMA) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(P.BN);
      mbar_wait(BAR(0), 0);
      int g = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        R3_WAIT_SLEEP(BAR(1 + s), ph, 32, 1, it);
        R3_EV(1, it);
        tc_fence_after();
        const uint32_t zaddr = sbase + P.off_z + s * R3_ZSTAGE;
        for (int blk = 0; blk < P.nb; ++blk, ++g) {
          const int a = g & 1, aph = (g >> 1) & 1;
          R3_WAIT_SLEEP(BAR(7 + a), aph ^ 1, 32, 2, it);
          R3_EV(2 + blk, it);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)a * TC_MAXBN;
          uint32_t acc = 0;
#pragma unroll
          for (int c = 0; c < nD; ++c) {
            const uint32_t eaddr = sbase + P.off_emain + (uint32_t)(blk * nD + c) * P.BN * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = make_desc(zaddr + c * (TC_TILE * 128) + ks * 1024, 4096, 512, 1);
              const uint64_t bd = make_desc(eaddr + ks * 32, 16, 1024, 2);
              umma_tf32(d_tmem, ad, bd, idesc, acc);
              acc = 1;
            }
          }
          {
            const uint64_t ad = make_desc(sbase + P.off_aaug, 1024, 512, 1);
            const uint64_t bd = make_desc(sbase + P.off_eaug + (uint32_t)blk * P.BN * 32, 128, 256, 0);
            umma_tf32(d_tmem, ad, bd, idesc, acc);
          }
          umma_commit(BAR(5 + a));
        }
        umma_commit(BAR(3 + s));               // the tensor core is done reading this z stage
      }
    }
  } else if (warp == R3_W_ZN) {
    // ===================================== |z|^2 worker =====================================
    // lane = pixels 4 lane .. 4 lane + 3, one 16-byte load per channel.
    // Per pixel the same two ascending-d fma chains (even / odd channel quads) as everywhere in this library.
    const uint32_t zrow0 = sbase + P.off_z + (uint32_t)(lane >> 3) * 4096;
    uint32_t zx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zx[i] = (uint32_t)i * 128 + (uint32_t)(((lane & 7) ^ (i << 1)) << 4);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      R3_WAIT(BAR(1 + s), ph, 3, it);
      uint32_t zc = zrow0 + s * R3_ZSTAGE;
      float2 a01 = make_float2(0.f, 0.f), a23 = a01, b01 = a01, b23 = a01;
#pragma unroll
      for (int c = 0; c < nD; ++c) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = lds_v4(zc + jj * 512 + zx[i]);
            if ((jj & 1) == 0) {
              a01 = __ffma2_rn(make_float2(v.x, v.y), make_float2(v.x, v.y), a01);
              a23 = __ffma2_rn(make_float2(v.z, v.w), make_float2(v.z, v.w), a23);
            } else {
              b01 = __ffma2_rn(make_float2(v.x, v.y), make_float2(v.x, v.y), b01);
              b23 = __ffma2_rn(make_float2(v.z, v.w), make_float2(v.z, v.w), b23);
            }
          }
        }
        zc += 16384;
      }
      a01 = __fadd2_rn(a01, b01);
      a23 = __fadd2_rn(a23, b23);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + P.off_zn + (uint32_t)((it % 3) * TC_TILE + 4 * lane) * 4),
                   "f"(a01.x), "f"(a01.y), "f"(a23.x), "f"(a23.y) : "memory");
      __syncwarp();
      R3_EV(4, it);
      if (lane == 0) { mbar_arrive(BAR(9 + (it % 3))); mbar_arrive(BAR(3 + s)); }
    }
  } else if (warp < R3_SCAN_WARPS) {
    // ===================================== scan warps =======================================
    // thread = (pixel, column group): maxima of its columns of both accumulator blocks, bounds, records; the results go to
    // the `pub` slot (the output warps merge the two column groups and decide)
    const int quad = warp & 3, cg = warp >> 2;
    const int p = quad * 32 + lane;                       // pixel within the tile == TMEM lane
    const uint32_t ctab_s = sbase + P.off_ctab;
    const uint32_t pub_s = sbase + P.off_pub + (uint32_t)(cg * TC_TILE + p) * 16;
    const int nchunks = P.BN >> 5;
    const float rminbig = __uint_as_float(P.meta[1]);     // +inf when no code is excluded
    int g = 0;
    for (int it = 0; it < my_tiles; ++it) {
      R3Scan st;
      st.L = -INFINITY; st.Urec = -INFINITY; st.cnt = 0; st.recA = 0; st.recB = 0;
      // |z|^2 of this tile: ready long before the first accumulator block (same load, no tensor-core work in between)
      R3_WAIT(BAR(9 + (it % 3)), (it / 3) & 1, 4, it);
      const float z2 = lds_f32(sbase + P.off_zn + (uint32_t)((it % 3) * TC_TILE + p) * 4);
      const float zn = sqrtf(z2) * 1.00001f;
      for (int blk = 0; blk < P.nb; ++blk, ++g) {
        const int a = g & 1, aph = (g >> 1) & 1;
        float hm[16], R[8];
        if (warp == 0) R3_EV(5 + 3 * blk, it);
        R3_WAIT(BAR(5 + a), aph, 5, it);
        if (warp == 0) R3_EV(6 + 3 * blk, it);
        tc_fence_after();
        if (cg < nchunks) r3_block_maxima(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)a * TC_MAXBN, cg, nchunks, hm, R);
        tc_fence_before();
        __syncwarp();
        if (warp == 0) R3_EV(7 + 3 * blk, it);
        if (lane == 0) mbar_arrive(BAR(7 + a));           // the accumulator block may be overwritten
        if (cg < nchunks) r3_block_end(st, hm, R, cg, nchunks, blk, ctab_s, zn);
      }
      // flags: 1 overflow, 2 non-finite |z|^2, 4 an excluded ("big") code could still beat this group's lower bound
      // (s_k <= r_k (2|z| - r_k), decreasing in r_k >= |z|)
      uint32_t flags = st.cnt > 2 ? 1u : 0u;
      if (!(z2 <= 3.0e38f)) flags |= 2u;
      if (rminbig < 3.0e38f) {
        const float lbest = 2.f * (st.L < 0.f ? st.L * 1.0009765625f : st.L);          // lower bound on the best exact 2 a_k
        const float bigub = rminbig * (2.f * zn - rminbig);
        if (!((rminbig >= zn) && (bigub + 1e-5f * (fabsf(bigub) + fabsf(lbest)) < lbest))) flags |= 4u;
      }
      const uint32_t w1 = f32_up16(st.Urec) | flags;
      if (it > 0) R3_WAIT(BAR(24 + ((it - 1) & 1)), ((it - 1) >> 1) & 1, 6, it);     // the merging warps have read the slot (tile it - 1)
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(pub_s), "r"(__float_as_uint(st.L)), "r"(w1),
                   "r"(st.cnt >= 1 ? st.recA : 0u), "r"(st.cnt >= 2 ? st.recB : 0u) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(21 + (it & 1)));
      if (warp == 0) R3_EV(14, it);
    }
  } else if (r3_out_index(warp) >= 0) {
    // ===================================== output warps =====================================
    // iteration j: merge(j) (my team's tiles only), then outputs(j - 2)
    const int ow = r3_out_index(warp);
    // outputs: pixels 4 pq .. 4 pq + 3 of the tile, channel quads co and co + 8 (channels 4 co .. + 3 and 32 + 4 co .. + 3);
    // a quarter-warp = eight adjacent pixel quads of one channel = one 128-byte line of z / q
    const int pq = 8 * (ow & 3) + (lane & 7), co = 4 * (ow >> 2) + (lane >> 3);
    const uint32_t emain = sbase + P.off_emain;
    float* sums_mine = nullptr;
    if (STATS) {
      const int rep = (int)(blockIdx.x % (unsigned)P.nrep);
      sums_mine = (rep == 0 ? P.sums : P.sums_rep + (size_t)(rep - 1) * P.K * R3_D) + 4 * co;
    }
    const size_t hw = (size_t)P.HW;
    const size_t off_out = (size_t)(4 * co) * hw + 4 * pq;         // my first channel row / pixel inside an image's tile
    float2 lsA = make_float2(0.f, 0.f), lsB = lsA;
    int tb = (int)blockIdx.x / P.tiles_per_img, tpt = (int)blockIdx.x % P.tiles_per_img;   // tile j
    int b1 = 0, p01 = 0, b2 = 0, p02 = 0, b3 = 0, p03 = 0;                                 // tiles j - 1, j - 2, j - 3
    mbar_wait(BAR(0), 0);                                 // codebook resident (read below with plain loads)
    for (int it = 0; it < my_tiles + R3_LAG; ++it) {
      const int b = tb, p0 = tpt * TC_TILE;
      tpt += (int)gridDim.x;
      while (tpt >= P.tiles_per_img) { tpt -= P.tiles_per_img; ++tb; }
      // ---- merge(it) ------------------------------------------------------------------------------------
      if (it < my_tiles && (ow >> 2) == (it & 1)) {
        const int ws = it & (R3_WINS - 1), wph = (it / R3_WINS) & 1;
        const int p = 32 * (ow & 3) + lane;
        if ((ow & 3) == 0) R3_EV(11, it);
        R3_WAIT_RELAXED(BAR(21 + (it & 1)), (it >> 1) & 1, 7, it);      // scan results of this tile
        if ((ow & 3) == 0) R3_EV(12, it);
        uint32_t pw[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pw[i][0]), "=r"(pw[i][1]), "=r"(pw[i][2]), "=r"(pw[i][3])
                       : "r"(sbase + P.off_pub + (uint32_t)(i * TC_TILE + p) * 16) : "memory");
        {
          const uint32_t lb = __reduce_or_sync(0xffffffffu, pw[0][0] | pw[1][0] | pw[0][3] | pw[1][3]);
          if (lane == 0) mbar_arrive(r3_after_loads(BAR(24 + (it & 1)), lb, PP.zero));      // the pub slot may be rewritten (tile it + 1)
        }
        const float Lg = fmaxf(__uint_as_float(pw[0][0]), __uint_as_float(pw[1][0]));
        // flags of the scan threads (low bits of word 1): 1 overflow, 2 non-finite |z|^2, 4 an excluded ("big") code could
        // still beat this group's own lower bound -- safe as soon as one group rules it out (its bound is <= the merged one)
        bool ovf = false;
        const bool bad = ((pw[0][1] | pw[1][1]) & 2u) != 0u;
        const bool big_unsafe = (pw[0][1] & pw[1][1] & 4u) != 0u;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (__uint_as_float(pw[i][1] & 0xFFFF0000u) >= Lg) ovf |= (pw[i][1] & 1u) != 0;
          else { pw[i][2] = 0; pw[i][3] = 0; }
        }
        const uint32_t c0 = pw[0][2], c1 = pw[0][3], c2 = pw[1][2], c3 = pw[1][3];     // records of column group 0, 0, 1, 1
        const int total = r3_cells(c0) + r3_cells(c1) + r3_cells(c2) + r3_cells(c3);
        bool fb = bad || ovf || total == 0 || big_unsafe || total > R3_MAXCAND;
        bool defer = !fb && total > 1;
        {   // at most 8 queued pixels per warp and tile (the queue holds R3_QCAP): the rest goes to the fallback list
          const uint32_t dm = __ballot_sync(0xffffffffu, defer);
          if (defer && __popc(dm & ((1u << lane) - 1u)) >= 8) { defer = false; fb = true; }
        }
        const int pp = p0 + p;
        R3_WAIT_RELAXED(BAR(17 + ws), wph ^ 1, 8, it);    // every output warp has read this win slot (tile it - 4)
        uint32_t wv1 = R3_WIN_SKIP;
        if (fb) {
          const int fslot = atomicAdd(P.fb_count, 1);
          P.fb_rows[fslot] = b * P.HW + pp;
        } else if (!defer) {
          const uint32_t rec = c0 | c1 | c2 | c3;         // exactly one of them is non-zero
          const int cgo = (c0 | c1) ? 0 : 1;
          const int q = __clz((rec & 0x00FFFF00u) << 8), j = __clz(rec << 24);
          const int w = r3_cell_pos(rec, cgo, q, j, bnsh);
          uint32_t worig;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(worig) : "r"(sbase + P.off_perm + (uint32_t)w * 2));
          wv1 = (uint32_t)w | (worig << 16);
          int h, wc;
          if (P.w_shift >= 0) { h = pp >> P.w_shift; wc = pp & (P.W - 1); }
          else { h = pp / P.W; wc = pp - h * P.W; }
          const size_t nb_ = (size_t)b * hw;
          if (P.ids) store_id(P.ids + nb_, pp, h, wc, P.H, (int)worig, P.ids_mode);
          if (P.ids_nat) P.ids_nat[nb_ + pp] = (int)worig;
          if (STATS) atomicAdd(&hist[worig >> 1], (worig & 1u) ? 65536u : 1u);
        }
        if (!defer) asm volatile("st.shared.u32 [%0], %1;" ::"r"(win_s + (uint32_t)(ws * TC_TILE + p) * 4), "r"(wv1) : "memory");
        // ---- queue the pixels with several candidate cells; their z rows are gathered with cp.async -----------
        const uint32_t dmask = __ballot_sync(0xffffffffu, defer);
        if (dmask) {
          const int n = __popc(dmask);
          uint32_t base = 0;
          if (lane == 0) {
            base = atomicAdd((uint32_t*)&qctl[0], (uint32_t)n);
            while ((int)(base + (uint32_t)n - qctl[1]) > R3_QCAP) {                    // space in the ring
#ifdef VQ_R3_CHECK
              if (*(volatile int*)&g_r3_abort) break;
#endif
              __nanosleep(100);
            }
          }
          base = __shfl_sync(0xffffffffu, base, 0);
          if (defer) {
            const uint32_t slot = (base + (uint32_t)__popc(dmask & ((1u << lane) - 1u))) % R3_QCAP;
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(queue_s + slot * R3_QSTRIDE + 256), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(qitp_s + slot * 4), "r"(((uint32_t)it << 7) | (uint32_t)p) : "memory");
            __threadfence_block();
          }
          __syncwarp();
          const float* zimg = P.z + (size_t)b * R3_D * hw + (size_t)lane * hw;       // channel `lane` of the image
          uint32_t rank = 0;
          for (uint32_t m = dmask; m; m &= m - 1u, ++rank) {
            const int src = __ffs(m) - 1;
            const int ppm = __shfl_sync(0xffffffffu, pp, src);
            const uint32_t slot = (base + rank) % R3_QCAP;
            const uint32_t dst = queue_s + slot * R3_QSTRIDE + (uint32_t)lane * 4;
            cp_async4(dst, zimg + ppm);
            cp_async4(dst + 128, zimg + 32 * hw + ppm);
            cp_async_arrive(qbar_s + slot * 8);
          }
        }
        __syncwarp();
        const int ndec = 32 - __popc(dmask);
        if (lane == 0 && ndec > 0) mbar_arrive_cnt(BAR(13 + ws), (uint32_t)ndec);
        if ((ow & 3) == 0) R3_EV(13, it);
      }
      // ---- outputs(it - R3_LAG) ---------------------------------------------------------------------------
      static_assert(R3_LAG == 3 && R3_LAG < R3_WINS, "tile bookkeeping below");
      if (it >= R3_LAG) {
        const int t = it - R3_LAG, ws = t & (R3_WINS - 1), wph = (t / R3_WINS) & 1;
        if (ow == 0) R3_EV(15, t);
        const size_t img = (size_t)b3 * R3_D * hw + p03 + off_out;
        float4 z4[8];                                     // z4[i]: channel 4 co + i (i < 4), 32 + 4 co + i - 4 (i >= 4) of my four pixels
        {
          const float* zg = P.z + img;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            z4[i] = __ldg(reinterpret_cast<const float4*>(zg + (size_t)i * hw));
            z4[4 + i] = __ldg(reinterpret_cast<const float4*>(zg + (size_t)(32 + i) * hw));
          }
        }
        R3_WAIT_RELAXED(BAR(13 + ws), wph, 9, t);         // winners of the tile (decided by the merge, or by the straggler warp)
        if (ow == 0) R3_EV(18, t);
        uint32_t wv[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wv[0]), "=r"(wv[1]), "=r"(wv[2]), "=r"(wv[3])
                     : "r"(win_s + (uint32_t)(ws * TC_TILE + 4 * pq) * 4) : "memory");
        {
          const uint32_t lb = __reduce_or_sync(0xffffffffu, wv[0] | wv[1] | wv[2] | wv[3]);
          if (lane == 0) mbar_arrive(r3_after_loads(BAR(17 + ws), lb, PP.zero));     // the slot may be rewritten (tile t + 4)
        }
        float4 ea[4], eb[4];                              // code row of pixel x: channels 4 co .. + 3 and 32 + 4 co .. + 3
        float msk[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const bool skip = (wv[x] & R3_WIN_SKIP) != 0u;
          uint32_t w = skip ? 0u : (wv[x] & 0xFFFFu);
#ifdef VQ_R3_CHECK
          if (w >= (uint32_t)ktot || (!skip && ((wv[x] >> 16) & 0x7FFFu) >= (uint32_t)P.K)) { atomicAdd(P.fb_count + 2, 1); P.fb_count[3] = (int)wv[x]; P.fb_count[4] = t; w = 0u; }
#endif
          msk[x] = skip ? 0.f : 1.f;
          const uint32_t kb = w >> bnsh, row = w & (uint32_t)(P.BN - 1);
          const uint32_t er = emain + kb * (uint32_t)nD * bn128 + row * 128 + ((((uint32_t)co) ^ (row & 7)) << 4);
          ea[x] = lds_v4(er);
          eb[x] = lds_v4(er + bn128);
        }
        // q: one 16-byte store per channel, a quarter-warp writes one full 128-byte line; (z - q)^2
        float* qo = P.q + img;
        float2 l01 = make_float2(0.f, 0.f), l23 = l01;
        const float2 m1 = make_float2(-1.f, -1.f);
#define R3_CH(i, row_, F, C)                                                                             \
        {                                                                                                \
          const float4 qv = make_float4(F[0].C, F[1].C, F[2].C, F[3].C);                                 \
          __stcs(reinterpret_cast<float4*>(qo + (size_t)(row_) * hw), qv);                               \
          const float2 d01 = __ffma2_rn(make_float2(qv.x, qv.y), m1, make_float2(z4[i].x, z4[i].y));     \
          const float2 d23 = __ffma2_rn(make_float2(qv.z, qv.w), m1, make_float2(z4[i].z, z4[i].w));     \
          l01 = __ffma2_rn(d01, d01, l01);                                                               \
          l23 = __ffma2_rn(d23, d23, l23);                                                               \
        }
        R3_CH(0, 0, ea, x) R3_CH(1, 1, ea, y) R3_CH(2, 2, ea, z) R3_CH(3, 3, ea, w)
        R3_CH(4, 32, eb, x) R3_CH(5, 33, eb, y) R3_CH(6, 34, eb, z) R3_CH(7, 35, eb, w)
#undef R3_CH
        lsA = __ffma2_rn(l01, make_float2(msk[0], msk[1]), lsA);
        lsB = __ffma2_rn(l23, make_float2(msk[2], msk[3]), lsB);
        // EMA sums: equal codes among the four pixels are added up first
        if (STATS) {
          uint32_t wo[4];
          bool live[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) { wo[x] = (wv[x] >> 16) & 0x7FFFu; live[x] = (wv[x] & R3_WIN_SKIP) == 0u; }
          if (live[0] && live[1] && wo[0] == wo[1]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z4[i].x += z4[i].y;
            live[1] = false;
          }
          if (live[2] && live[3] && wo[2] == wo[3]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z4[i].z += z4[i].w;
            live[3] = false;
          }
          if (live[0] && live[2] && wo[0] == wo[2]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z4[i].x += z4[i].z;
            live[2] = false;
          }
#define R3_RED(x, C)                                                                                              \
          if (live[x]) {                                                                                         \
            float* so = sums_mine + (size_t)wo[x] * R3_D;                                                        \
            red_add_v4(so, z4[0].C, z4[1].C, z4[2].C, z4[3].C);                                                  \
            red_add_v4(so + 32, z4[4].C, z4[5].C, z4[6].C, z4[7].C);                                             \
          }
          R3_RED(0, x) R3_RED(1, y) R3_RED(2, z) R3_RED(3, w)
#undef R3_RED
        }
        if (ow == 0) R3_EV(19, t);
      }
      b3 = b2; p03 = p02; b2 = b1; p02 = p01; b1 = b; p01 = p0;
    }
    __syncwarp();
    if (lane == 0) { __threadfence_block(); atomicAdd((uint32_t*)&qctl[2], 1u); }     // no more queue entries from this warp
    float lsum = (lsA.x + lsA.y) + (lsB.x + lsB.y);
    lsum = warp_sum(lsum);
    if (lane == 0 && P.loss_acc && lsum != 0.f) atomicAdd(P.loss_acc, (double)lsum);
  } else if (warp == R3_W_STRAG) {
    // ===================================== straggler warp ===================================
    // Exact re-rank of the queued pixels.  A batch = the leading complete entries of the queue (<= 16).  Every lane pair
    // (lp = lane & 15, channel-quad parity hf = lane >> 4) re-scores ONE (entry, candidate cell) per pass -- the cells of all
    // entries of the batch are laid end to end, so a typical batch (6 entries x 2..3 cells) is one pass -- with the library's
    // dot product (two ascending-d fma chains over the even / odd channel quads, joined by one addition); the best key of
    // an entry is collected with a shared-memory atomicMax.
    const int lp = lane & 15, hf = lane >> 4;
    const uint32_t emain = sbase + P.off_emain;
    const uint32_t perm_a = sbase + P.off_perm;
    const size_t hw = (size_t)P.HW;
    mbar_wait(BAR(0), 0);
    uint32_t head = 0;
    while (true) {
      // entries head .. head + n - 1 have landed (their barriers, in order)
      const uint32_t idx = head + (uint32_t)lp;
      const uint32_t slot = idx % R3_QCAP;
      const bool ready = mbar_test(qbar_s + slot * 8, (idx / R3_QCAP) & 1u);
      const uint32_t rmask = __ballot_sync(0xffffffffu, ready) & 0xFFFFu;
      const int n = min(__ffs(~rmask) - 1, R3_QCAP);      // leading run of complete entries
      if (n == 0) {
        if (qctl[2] == (uint32_t)R3_OUT_WARPS && qctl[0] == head) break;      // every merging warp is done and the queue is empty
        __nanosleep(100);
        continue;
      }
      // ---- setup: lane pair lp owns entry lp: records, number of cells, |z|^2 ---------------------------
      const bool own = lp < n;
      const uint32_t ent = queue_s + slot * R3_QSTRIDE;
      uint32_t itp = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      float z2 = 0.f;
      if (own) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(itp) : "r"(qitp_s + slot * 4) : "memory");
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3) : "r"(ent + 256) : "memory");
#pragma unroll
        for (int t = 0; t < 8; ++t) {                     // |z|^2 = A + B: lane hf holds the channels of chain hf, ascending d
          const float4 v = lds_v4(ent + (uint32_t)(2 * t + hf) * 16);
          z2 = __fmaf_rn(v.x, v.x, z2); z2 = __fmaf_rn(v.y, v.y, z2); z2 = __fmaf_rn(v.z, v.z, z2); z2 = __fmaf_rn(v.w, v.w, z2);
        }
      }
      z2 = __fadd_rn(z2, __shfl_xor_sync(0xffffffffu, z2, 16));
      const int it = (int)(itp >> 7), p = (int)(itp & 127u);
      const int it_lead = __shfl_sync(0xffffffffu, it, 0);
      R3_EV(20, it_lead);
      const int cnt = own ? r3_cells(c0) + r3_cells(c1) + r3_cells(c2) + r3_cells(c3) : 0;
      int pre = cnt;                                      // inclusive prefix sum over the entries (lanes 0..15; both halves alike)
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, o, 16);
        if (lp >= o) pre += v;
      }
      const int total = __shfl_sync(0xffffffffu, pre, 15, 16);
      const int pre_ex = own ? pre - cnt : 0x7FFFFFFF;    // first flat cell index of my entry (entries beyond the batch: never matched)
      if (own && hf == 0) {
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(qkey_s + (uint32_t)lp * 8), "r"(0u), "r"(0u) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(qz2_s + (uint32_t)lp * 4), "f"(z2) : "memory");
      }
      __syncwarp();
      R3_EV(22, it_lead);
      int niter = 0;
      for (int f0 = 0; f0 < total; f0 += 16) {
        ++niter;
        const int fidx = f0 + lp;
        const bool on = fidx < total;
        // entry of flat cell fidx: the last entry whose first cell index is <= fidx (binary search over the lanes' pre_ex)
        int ee = 0;
#pragma unroll
        for (int st = 8; st >= 1; st >>= 1) {
          const int pc = __shfl_sync(0xffffffffu, pre_ex, ee + st, 16);
          if (pc <= fidx) ee += st;
        }
        const int pe = __shfl_sync(0xffffffffu, pre_ex, ee, 16);
        const uint32_t slot_e = (head + (uint32_t)ee) % R3_QCAP;
        const uint32_t ent_e = queue_s + slot_e * R3_QSTRIDE;
        float dot_o = 0.f, e2k_o = 0.f, z2e_o = 0.f;
        uint32_t korig_o = 0u;
        int k_o = 0;
        if (on) {
          uint32_t r0, r1, r2, r3;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(ent_e + 256) : "memory");
          // the j-th cell of the entry: record, then (quarter, residue) = (j / residues, j % residues) among the set bits
          int j = fidx - pe, ri = 0;
          uint32_t rec = r0;
          int cr = r3_cells(r0);
          if (j >= cr) { j -= cr; rec = r1; ri = 1; cr = r3_cells(r1);
            if (j >= cr) { j -= cr; rec = r2; ri = 2; cr = r3_cells(r2);
              if (j >= cr) { j -= cr; rec = r3; ri = 3; } } }
          const uint32_t qm = __brev((rec >> 8) & 0xFFFFu) >> 16, rm = __brev(rec & 0xFFu) >> 24;     // bit q / bit j set: hot
          const int nr = __popc(rm);
          const int qi = j / nr, rj = j - qi * nr;
          const int q = (int)__fns(qm, 0, qi + 1), jr = (int)__fns(rm, 0, rj + 1);
          const int k = r3_cell_pos(rec, ri >> 1, q, jr, bnsh);
          const int kb = k >> bnsh, row = k & (P.BN - 1);
          const uint32_t eb = emain + (uint32_t)(kb * nD) * bn128 + (uint32_t)row * 128, r7 = (uint32_t)(row & 7) << 4;
          uint32_t korig;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(korig) : "r"(perm_a + (uint32_t)k * 2));
          const float e2k = lds_f32(sbase + P.off_eaug + (uint32_t)(((kb << bnsh) << 5) + (row >> 3) * 256 + (row & 7) * 16) + 12);
          const float z2e = lds_f32(qz2_s + (uint32_t)ee * 4);
          float dot = 0.f;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const uint32_t jq = (uint32_t)(2 * t) + (uint32_t)hf;
            const float4 zv = lds_v4(ent_e + jq * 16);
            const float4 e4 = lds_v4(eb + (uint32_t)((2 * t) >> 3) * bn128 + (((jq & 7) << 4) ^ r7));
            dot = __fmaf_rn(zv.x, e4.x, dot);
            dot = __fmaf_rn(zv.y, e4.y, dot);
            dot = __fmaf_rn(zv.z, e4.z, dot);
            dot = __fmaf_rn(zv.w, e4.w, dot);
          }
          e2k_o = e2k; z2e_o = z2e; korig_o = korig; k_o = k;
          dot_o = dot;
        }
        dot_o = __fadd_rn(dot_o, __shfl_xor_sync(0xffffffffu, dot_o, 16));      // A + B (commutative: same bits in both lanes)
        if (on && hf == 0) {
          const unsigned long long kc = ((unsigned long long)f32_orderable(ref_score(dot_o, e2k_o, z2e_o)) << 32) |
                                        ((unsigned long long)(0xFFFFu - korig_o) << 16) | (unsigned long long)k_o;
          atomicMax((unsigned long long*)(smem + (qkey_s - sbase) + (size_t)ee * 8), kc);        // ties: lowest ORIGINAL index
        }
        __syncwarp();
      }
      R3_EV(21, it_lead);
#ifdef VQ_R3_TRACE
      if (blockIdx.x == 0 && lane == 0 && it_lead < 64) { g_r3_trace[23 * 64 + it_lead] = niter; g_r3_trace[24 * 64 + it_lead] = n; }
#endif
      // ---- winners: win ring, code map, histogram ------------------------------------------------------
      const int ws = it & (R3_WINS - 1);
      if (own && hf == 0) {
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int b = tile / P.tiles_per_img, pp = (tile % P.tiles_per_img) * TC_TILE + p;
        uint32_t klo, khi;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(klo), "=r"(khi) : "r"(qkey_s + (uint32_t)lp * 8) : "memory");
        const int w = (int)(klo & 0xFFFFu);
        const uint32_t worig = 0xFFFFu - (klo >> 16);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(win_s + (uint32_t)(ws * TC_TILE + p) * 4), "r"((uint32_t)w | (worig << 16)) : "memory");
        int h, wc;
        if (P.w_shift >= 0) { h = pp >> P.w_shift; wc = pp & (P.W - 1); }
        else { h = pp / P.W; wc = pp - h * P.W; }
        const size_t nb_ = (size_t)b * hw;
        if (P.ids) store_id(P.ids + nb_, pp, h, wc, P.H, (int)worig, P.ids_mode);
        if (P.ids_nat) P.ids_nat[nb_ + pp] = (int)worig;
        if (STATS) atomicAdd(&hist[worig >> 1], (worig & 1u) ? 65536u : 1u);
      }
      uint32_t am[R3_WINS];
#pragma unroll
      for (int i = 0; i < R3_WINS; ++i) am[i] = __ballot_sync(0xffffffffu, own && hf == 0 && ws == i);
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
#pragma unroll
        for (int i = 0; i < R3_WINS; ++i)
          if (am[i]) mbar_arrive_cnt(BAR(13 + i), (uint32_t)__popc(am[i]));
        qctl[1] = head + (uint32_t)n;                     // the entries may be overwritten
      }
      head += (uint32_t)n;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (STATS) {
    for (int i = threadIdx.x; i < (P.K + 1) / 2; i += R3_THREADS) {
      const uint32_t c = hist[i];
      if (c & 0xFFFFu) red_add_s32(&P.counts[2 * i], (int)(c & 0xFFFFu));
      if (c >> 16) red_add_s32(&P.counts[2 * i + 1], (int)(c >> 16));
    }
  }
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
constexpr int TCS_MAX_ND = 16;      // D <= 512 (VQGAN: emb_dim 512, vqgan.py:389-390)
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
        }
      }
      if (STATS) {
        // histogram: the pixels of the warp that chose the same code share one atomic (a million single increments on
        // 512 addresses serialise in L2 -- worst when a few codes take most of the pixels)
        const int key = (hf == 0 && !fb) ? (int)worig : -1 - lane;
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs(grp) - 1) atomicAdd(&P.counts[key], __popc(grp));
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

// third-generation resident kernel (emb_dim 64, 257..512 codes): opt-in with VQ_B200_R3=1 in the environment.  Measured on
// B200: 15 % faster than the default kernel on clustered input (few ambiguous pixels), slower on Gaussian /
// post-ReLU input, where ~5 % of the pixels need the exact re-rank and its single straggler warp sets the pace
static bool r3_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VQ_B200_R3"); on = (e && e[0] == '1') ? 1 : 0; }
  return on != 0;
}
static bool r3_supported(const FwdArgs& a) {
  // the per-CTA histogram packs two 16-bit counters per word: a CTA must see fewer than 65536 pixels
  const long long ntiles = (long long)a.B * ((long long)a.H * a.W / TC_TILE);
  const int sms = device_sm_count() > 0 ? device_sm_count() : 1;
  if ((ntiles + sms - 1) / sms * TC_TILE >= 65536) return false;
  return a.D == R3_D && r3_geometry(a.K).ok && r3_geometry(a.K).BN == TC_MAXBN && r3_geometry(a.K).nb == 2 && (a.H * a.W) % TC_TILE == 0 && r3_geometry(a.K).nb * r3_geometry(a.K).BN <= TC_SORT_MAX;
}

static int launch_assign_r3(const FwdArgs& a, cudaStream_t s) {
  const int HW = a.H * a.W;
  const R3Geom g = r3_geometry(a.K);
  VQ_REQUIRE(g.ok && a.D == R3_D && HW % TC_TILE == 0, VQ_ERR_UNSUPPORTED, "tensor-core path (resident, emb_dim 64): unsupported shape");
  VQ_REQUIRE(a.q != nullptr, VQ_ERR_INVALID_ARG, "tensor-core path: q must not be null");
  EncodeTiledFn enc = get_encode_fn();
  VQ_REQUIRE(enc != nullptr, VQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  VQ_REQUIRE((((uintptr_t)a.z) & 15) == 0 && (((uintptr_t)a.embed) & 15) == 0 && (((uintptr_t)a.q) & 15) == 0, VQ_ERR_INVALID_ARG,
             "tensor-core path: z / embed / q must be 16-byte aligned");
  CUtensorMap zmap, emap;
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
  {
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
  vq_tc_prep2_kernel<<<(ktot + TC_PREP2_CODES - 1) / TC_PREP2_CODES, TC_PREP2_THREADS, (size_t)a.K * 4, s>>>(
      a.embed, a.ws.e2, a.K, a.D, g.BN, g.nb, a.ws.tc_es, eaug_img, a.ws.tc_perm, rmax, meta);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());

  R3Params PP{};
  TcParams& P = PP.t;
  P.z = a.z; P.E = a.embed; P.e2 = a.ws.e2; P.eaug_img = eaug_img; P.meta = meta;
  P.perm = a.ws.tc_perm; P.rmax = rmax;
  P.B = a.B; P.D = a.D; P.H = a.H; P.W = a.W; P.HW = HW; P.K = a.K;
  P.BN = g.BN; P.nb = g.nb; P.nD = R3_ND; P.nst = 2;
  P.bn_shift = 0;
  while ((1 << P.bn_shift) < g.BN) ++P.bn_shift;
  P.w_shift = -1;
  if ((a.W & (a.W - 1)) == 0) { P.w_shift = 0; while ((1 << P.w_shift) < a.W) ++P.w_shift; }
  P.tiles_per_img = HW / TC_TILE;
  P.ntiles = a.B * P.tiles_per_img;
  P.off_emain = (uint32_t)g.off_emain; P.off_eaug = (uint32_t)g.off_eaug; P.off_aaug = (uint32_t)g.off_aaug;
  P.off_z = (uint32_t)g.off_z; P.off_pub = (uint32_t)g.off_pub; P.off_zn = (uint32_t)g.off_zn;
  P.off_hist = (uint32_t)g.off_hist; P.off_perm = (uint32_t)g.off_perm; P.off_ctab = (uint32_t)g.off_ctab;
  P.off_bar = (uint32_t)g.off_bar;
  PP.off_win = (uint32_t)g.off_win; PP.off_queue = (uint32_t)g.off_queue; PP.zero = 0u;
  P.ids = a.ids; P.ids_nat = a.ids_nat; P.q = a.q; P.loss_acc = a.ws.loss_acc;
  P.counts = a.stats ? a.ws.counts : nullptr;
  P.sums = a.stats ? a.stats + stats_sums_offset(a.K) : nullptr;
  P.sums_rep = a.ws.sums_rep;
  P.nrep = tc_sums_replicas(a.K, a.D);
  P.fb_count = a.ws.misc; P.fb_rows = a.ws.fb_rows;
  P.ids_mode = ids_mode_of(a.flags);
  P.dbg = nullptr;

  int grid = sm_count_tc();
  if (grid > P.ntiles) grid = P.ntiles;
  const bool stats = a.stats != nullptr;
  typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const R3Params);
  KernFn kern = stats ? vq_assign_r3_kernel<true> : vq_assign_r3_kernel<false>;
  static bool attr_set[kMaxDevices][2] = {};
  const int dev = current_device();
  if (!attr_set[dev][stats ? 1 : 0]) {
    VQ_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    attr_set[dev][stats ? 1 : 0] = true;
  }
  const bool prof = profile_begin(s);
  kern<<<grid, R3_THREADS, g.total, s>>>(zmap, emap, PP);
  if (prof) profile_end(s);
  count_launch();
  VQ_CUDA_CHECK(cudaGetLastError());
  return VQ_OK;
}

int launch_assign_tc_impl(const FwdArgs& a, float* dbg, cudaStream_t s) {
  const int HW = a.H * a.W;
  if (!dbg && r3_enabled() && r3_supported(a)) return launch_assign_r3(a, s);
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
  vq_tc_prep2_kernel<<<(ktot + TC_PREP2_CODES - 1) / TC_PREP2_CODES, TC_PREP2_THREADS, (size_t)a.K * 4, s>>>(
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
  vq_tc_prep2_kernel<<<(ktot + TC_PREP2_CODES - 1) / TC_PREP2_CODES, TC_PREP2_THREADS, (size_t)a.K * 4, s>>>(
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
#ifdef VQ_R3_CHECK
  {
    const int tot3 = 32 * 64;
    if (n < tot3) return -1;
    if (cudaMemcpyFromSymbol(host_out, g_r3_dbg, sizeof(long long) * tot3) != cudaSuccess) return -2;
    return tot3;
  }
#endif
#ifdef VQ_R3_TRACE
  const int tot3 = 32 * 64;
  if (n < tot3) return -1;
  if (cudaMemcpyFromSymbol(host_out, g_r3_trace, sizeof(long long) * tot3) != cudaSuccess) return -2;
  return tot3;
#endif
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

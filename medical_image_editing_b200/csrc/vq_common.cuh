// vq_common.cuh -- shared declarations for the B200 VQ bottleneck kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/vq_b200.h"

namespace vqb200 {

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message, negative return codes; nothing throws across the C-ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define VQ_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::vqb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                          __FILE__, __LINE__);                                           \
      return VQ_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define VQ_REQUIRE(cond, code, ...)                                                      \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::vqb200::set_error(__VA_ARGS__);                                                  \
      return (code);                                                                     \
    }                                                                                    \
  } while (0)

// per-device launch state: one process may drive several GPUs (device index = cudaGetDevice() at the call)
constexpr int kMaxDevices = 64;
int current_device();                 // cudaGetDevice(), clamped to [0, kMaxDevices)
int device_sm_count();                // multiprocessors of the current device (cached per device)
int tuning_knob(const char* name, int dflt);   // integer environment switch, read at every call (A/B timing of launch shapes)

// measurement hooks (vq_launch_count / vq_profile_*)
void count_launch(int n = 1);
long long launch_count();
void profile_enable(bool on);
bool profile_begin(cudaStream_t s);   // records the start event if profiling is on
void profile_end(cudaStream_t s);     // records the stop event
int profile_read(double* total_ms, int* launches);

// ---------------------------------------------------------------------------------------------
// workspace layout (caller-owned memory, carved here; every region 256-byte aligned)
// ---------------------------------------------------------------------------------------------
constexpr int kCodePad = 256;  // codebook padded to a multiple of this many codes

struct Workspace {
  float* e2;         // [Kpad]   |e_k|^2, +inf for padding codes
  float* et;         // [D][Kpad] transposed codebook (0 for padding codes) -- SIMT search operand
  double* loss_acc;  // [1]      sum (z-q)^2
  int* counts;       // [K]      exact int32 histogram
  int* misc;         // [8]      misc[0] = number of rows routed to the exact fallback search
  int* fb_rows;      // [N]      rows routed to the exact fallback search (tensor-core path)
  // tensor-core path (vq_assign_tc.cu): codebook sorted by norm, augmentation image, permutation, bounds
  float* tc_es;      // [Kpad][D] codebook rows in ascending-norm order (zero rows for padding)
  float* tc_aug;     // [Kpad*8]  shared-memory image of the augmentation columns (-|e|^2/2 split in tf32)
  int* tc_perm;      // [Kpad]    sorted position -> original code index
  float* tc_ctab;    // [Kpad/32][2] per 32-code chunk: error-bound coefficients (A, B)
  float* tc_meta;    // [16]     per-call scalars (r_minbig, ...)
  float* sums_rep;   // [R-1][K*D] extra replicas of the per-code sums (R = tc_sums_replicas(K, D)); summed by finish
  size_t bytes;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline size_t stats_sums_offset(int K) { return align_up((size_t)2 * K, 4); }
inline int pad_codes(int K) { return (int)align_up((size_t)K, kCodePad); }

int tc_sums_replicas(int K, int D);                              // vq_assign_tc.cu

inline Workspace carve_workspace(void* base, int64_t N, int K, int D) {
  Workspace w;
  const int Kpad = pad_codes(K);
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
  w.e2 = (float*)take(sizeof(float) * Kpad);
  w.et = (float*)take(sizeof(float) * (size_t)D * Kpad);
  w.loss_acc = (double*)take(sizeof(double));
  w.counts = (int*)take(sizeof(int) * K);
  w.misc = (int*)take(sizeof(int) * 8);
  w.fb_rows = (int*)take(sizeof(int) * (size_t)(N > 0 ? N : 1));
  w.tc_es = (float*)take(sizeof(float) * (size_t)Kpad * D);
  w.tc_aug = (float*)take(sizeof(float) * (size_t)Kpad * 8);
  w.tc_perm = (int*)take(sizeof(int) * (size_t)Kpad);
  w.tc_ctab = (float*)take(sizeof(float) * (size_t)(Kpad / 32) * 2);
  w.tc_meta = (float*)take(sizeof(float) * 16);
  w.sums_rep = (float*)take(sizeof(float) * (size_t)(tc_sums_replicas(K, D) - 1) * K * D + 16);
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------------
// kernel launchers (each only enqueues on `stream`; returns a VQ_* code)
// ---------------------------------------------------------------------------------------------
struct FwdArgs {
  const float* z; int B, D, H, W;
  const float* embed; int K;
  int64_t* ids; int32_t* ids_nat; float* q; float* loss; float* stats; float* snapshot;
  Workspace ws;
  int flags = 0;
};

int launch_prep(const FwdArgs& a, bool tc_path, cudaStream_t s);
int launch_assign_simt(const FwdArgs& a, bool fallback_list_mode, cudaStream_t s);
int launch_assign_tc(const FwdArgs& a, cudaStream_t s);          // vq_assign_tc.cu
int launch_assign_small(const FwdArgs& a, cudaStream_t s);       // vq_assign_small.cu: K <= 32, K*D <= 1024
bool small_path_supported(int B, int D, int H, int W, int K);    // vq_assign_small.cu
bool tc_path_supported(int B, int D, int H, int W, int K);       // vq_assign_tc.cu
int launch_assign_tc_impl(const FwdArgs& a, float* dbg, cudaStream_t s);
int tc_debug_ncols(int D, int K);
int tc_debug_timing(long long* host_out, int n);
int launch_fallback_rows(const FwdArgs& a, cudaStream_t s);
int launch_onehot(const void* labels, int label_bytes, int64_t B, int64_t HW, int C, float* out, cudaStream_t s);
int launch_finish(const FwdArgs& a, bool tc_path, cudaStream_t s);
int launch_ema(float* cluster_size, float* embed_avg, long long avg_sd, long long avg_sk, float* embed,
               const float* stats, int K, int D, double momentum, double eps, float count_scale, float sum_scale,
               float* scratch, cudaStream_t s);
int launch_bwd(const float* g_q, const float* g_loss, const float* z, const int32_t* ids_nat,
               const float* snap, float* g_z, int B, int D, int H, int W, int K, cudaStream_t s);
int launch_lookup(const int64_t* ids, int64_t n, const float* embed, int K, int D, float* out, int layout,
                  int B, int A, int C, int* status, cudaStream_t s);
// vq_embed_loss.cu: cross-view cluster loss of EmbeddingLoss (functions/embed_loss.py)
size_t embed_loss_work_bytes(int B, int K);
int launch_embed_loss_fwd(const float* z, const int32_t* labels, const float* embed, int B, int D, int H, int W, int K,
                          float* loss, float* weights, void* work, cudaStream_t s);
int launch_embed_loss_bwd(const float* g_loss, const float* z, const int32_t* labels, const float* embed,
                          const float* weights, float* g_z, int B, int D, int H, int W, int K, cudaStream_t s);

// vq_norm_relu.cu: InstanceNorm2d + ReLU in front of the quantiser
int launch_norm_relu_fwd(const float* x, float* z, float* stats, long long planes, long long HW, float eps, cudaStream_t s);
int launch_norm_relu_bwd(const float* g_z, const float* x, const float* stats, float* g_x, long long planes, long long HW,
                         cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
// ids_mode = (flags / VQ_FLAG_IDS_NATURAL) & 3: bit 0 natural (b, h, w) order, bit 1 one-based
inline int ids_mode_of(int flags) { return (flags / VQ_FLAG_IDS_NATURAL) & 3; }
// the int64 code map entry of pixel p = h * W + w of an image whose map starts at `img`: reference layout (b, w, h) and
// 0-based by default (vq_module.py:171,178), or the (b, h, w) / 1-based form its callers build next (vqwnet.py:110-111)
__device__ __forceinline__ void store_id(int64_t* img, long long p, int h, int w, int H, int id, int ids_mode) {
  img[(ids_mode & 1) ? p : (long long)w * H + h] = (int64_t)(id + ((ids_mode >> 1) & 1));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The reference's score for code k and query z (vq_module.py:54-57):
//   s = fl( fl(2*dot) - |e|^2 ) - |z|^2     (2*dot is exact in binary fp)
// `dot` is, in every kernel of this library (CUDA-core search, exhaustive fallback, tensor-core re-rank), the same
// fp32 value: dot = fl(A + B) with A (B) the ascending-d fma chain over the channels whose quad index d >> 2 is even
// (odd).  Two chains are what the two lanes of a pixel in the tensor-core kernels compute from their own registers;
// identical summation order everywhere keeps the search paths bit-identical to each other.  |z|^2 = fl(A' + B') is
// split over the same two chains (the lanes of a pixel hold exactly those channels); |e|^2 is a single ascending-d chain.
__device__ __forceinline__ float ref_score(float dot, float e2, float z2) {
  return __fsub_rn(__fmaf_rn(2.0f, dot, -e2), z2);
}

}  // namespace vqb200

/*
 * vq_b200.h -- C-ABI of the B200-native vector-quantisation bottleneck.
 *
 * Drop-in boundary for the quantiser of Kaz-K/medical-image-editing
 * (reference: src/networks/vq/vq_module.py, src/networks/vq/grad_approximation.py).
 * The reference has no FFI of its own (it is pure PyTorch); these entry points are what a
 * ctypes binding inside `src/functions/vq_function.py` binds.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says HOST;
 *   - the caller owns all memory, including the workspace (size from vq_workspace_bytes);
 *   - the library never allocates device memory, never synchronises, never touches the
 *     default stream: every call only enqueues kernels on `stream`;
 *   - return value 0 = success, negative = error; the message is in vq_last_error()
 *     (thread-local).  No C++ exception crosses the boundary;
 *   - fp32 everywhere (reference runs without AMP, run_vqwnet.py:112);
 *   - there is no CPU fallback.
 */
#ifndef VQ_B200_H_
#define VQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t without pulling in the CUDA headers. */
typedef void* vq_stream_t;

/* error codes */
#define VQ_OK                 0
#define VQ_ERR_INVALID_ARG   -1
#define VQ_ERR_WORKSPACE     -2
#define VQ_ERR_CUDA          -3
#define VQ_ERR_UNSUPPORTED   -4

/* vq_assign_fwd flags */
#define VQ_FLAG_FORCE_SIMT    1   /* use the exact fp32 CUDA-core search (no tensor cores)   */
#define VQ_FLAG_NO_STATS      4   /* vq_assign_path only: the call will pass stats == NULL   */
#define VQ_FLAG_FORCE_TC      2   /* fail instead of silently choosing the CUDA-core search */
#define VQ_FLAG_PAIR          8   /* streamed-codebook tensor-core kernel: 2-CTA clusters (cta_group::2, each CTA holds
                                     half of every codebook slice) instead of one CTA per tile; same results, not
                                     faster on B200 -- kept for A/B timing */

#define VQ_FLAG_IDS_NATURAL  16   /* `ids` (int64) is written in natural (b, h, w) order instead of the reference's
                                     (b, w, h) order (vq_module.py:171,178): what every caller builds next with
                                     transpose(ids, 1, 2) (vqwnet.py:110, unet_encoder.py:115)                 */
#define VQ_FLAG_IDS_ONE_BASED 32  /* `ids` (int64) holds code + 1: the callers' `ids += 1` (vqwnet.py:111) folded
                                     into the epilogue; `ids_nat` (int32, for vq_bwd) stays 0-based            */

/* vq_lookup layouts */
#define VQ_LAYOUT_ROWS        0   /* out[n, D]  (F.embedding layout, vq_module.py:203-206)  */
#define VQ_LAYOUT_NCHW_T      1   /* ids [B,A,C] -> out [B,D,C,A]  contiguous: the tensor the
                                     callers build with lookup(ids).transpose(1,-1)
                                     (unet_encoder.py:120-123, vqwnet.py:158)               */

/* Library / ABI version (major*1000 + minor). */
int vq_version(void);

/* Last error message of the calling thread ("" if none). */
const char* vq_last_error(void);

/* Which search kernel vq_assign_fwd would use for this shape: 0 = fp32 CUDA-core (register-tiled),
 * 1 = tcgen05 tensor-core + exact fp32 re-rank, 2 = small-codebook fp32 kernel (inference calls, i.e. stats == NULL --
 * tell vq_assign_path with VQ_FLAG_NO_STATS -- with K <= 16 and K*D <= 256: the reference's real dictionaries,
 * run_recon.py:27-48). */
int vq_assign_path(int B, int D, int H, int W, int K, int flags);

/* Workspace bytes needed by vq_assign_fwd for N = B*H*W vectors. */
size_t vq_workspace_bytes(int64_t N, int K, int D);

/* Floats in the packed statistics buffer: [cnt_hi K | cnt_lo K | pad | sums K*D]
 * (count = cnt_hi*4096 + cnt_lo, both halves stay exactly representable in fp32 under an
 * all-reduce(sum) over <= 4096 ranks; sums start at vq_stats_sums_offset(K) floats, a multiple
 * of 4 so that rows of sums are 16-byte aligned whenever D % 4 == 0). */
size_t vq_stats_floats(int K, int D);
size_t vq_stats_sums_offset(int K);

/*
 * Nearest-code assignment, gather, commitment loss and (optionally) EMA statistics in one
 * pass over z.  Replaces VQModule._quantize + F.mse_loss (vq_module.py:159-186):
 *   flatten / k_nearest_neighbor(l2,k=1) / one_hot / lookup / one_hot.sum / flatten.T@one_hot.
 *
 *   z        [B,D,H,W] contiguous NCHW
 *   embed    [K,D]     the codebook (read-only here; q is gathered from THIS codebook,
 *                      i.e. pre-EMA-update, as vq_module.py:179 precedes :199)
 *   ids      [B,W,H]   int64, the reference's layout: ids[b,i,j] = code of pixel (h=j,w=i)
 *                      (vq_module.py:171,178 -- callers do transpose(ids,1,2))
 *   ids_nat  [B,H,W]   int32, natural order (kept for the backward pass); may be NULL
 *   q        [B,D,H,W] contiguous NCHW quantised output; may be NULL
 *   loss     scalar    mean((z-q)^2) over B*D*H*W  (F.mse_loss, vq_module.py:163); may be NULL
 *   stats    packed [cnt_hi K | cnt_lo K | pad | sums K*D] (see vq_stats_floats), sums[k*D+d] =
 *            sum of z over pixels assigned to k (== embed_sum[d,k], vq_module.py:185);
 *            NULL in eval mode
 *   embed_snapshot [K,D] copy of the codebook used for this call (backward needs the
 *            pre-update codebook); may be NULL
 */
int vq_assign_fwd(const float* z, int B, int D, int H, int W,
                  const float* embed, int K,
                  int64_t* ids, int32_t* ids_nat, float* q, float* loss,
                  float* stats, float* embed_snapshot,
                  void* workspace, size_t workspace_bytes, int flags, vq_stream_t stream);

/*
 * EMA codebook update from (possibly all-reduced) packed statistics.  Replaces
 * vq_module.py:194-199 (exponential_moving_average_ x2, Laplace smoothing, embed refresh).
 *   cluster_size [K], embed_avg [D,K], embed [K,D]: the module's three buffers, updated in place;
 *   embed_avg is addressed as embed_avg[d*avg_stride_d + k*avg_stride_k] (element strides): the
 *   reference builds it with `embed.T.clone()` (vq_module.py:156), which keeps the transposed
 *   strides (1, D), while a checkpoint round-trip may leave it contiguous (K, 1)
 *   momentum / eps are the Python doubles of the module (the reference forms 1-momentum and
 *   dict_size*eps in double before they meet the fp32 tensors);
 *   count_scale / sum_scale: multiply counts / sums before the update (1 for a single rank
 *   or global-batch semantics, 1/world_size for the reference's mean semantics)
 *   scratch: >= 16 bytes of device memory
 */
int vq_ema_update(float* cluster_size, float* embed_avg, int64_t avg_stride_d, int64_t avg_stride_k,
                  float* embed, const float* stats, int K, int D, double momentum, double eps,
                  float count_scale, float sum_scale, void* scratch, vq_stream_t stream);

/*
 * Backward of the whole module: straight-through estimator + commitment loss
 * (grad_approximation.py:7-29, F.mse_loss backward):
 *   g_z = g_q + g_loss * 2 * (z - embed_snapshot[ids]) / (B*D*H*W)
 *   g_q may be NULL (treated as 0); g_loss is a device scalar, may be NULL (treated as 0).
 */
int vq_bwd(const float* g_q, const float* g_loss, const float* z, const int32_t* ids_nat,
           const float* embed_snapshot, float* g_z,
           int B, int D, int H, int W, int K, vq_stream_t stream);

/*
 * Codebook gather.  Replaces VQModule.lookup = F.embedding(ids, embed) (vq_module.py:203-206).
 *   ids int64, n elements (for VQ_LAYOUT_NCHW_T: n = B*A*C with dims given by B,A,C);
 *   returns VQ_ERR_INVALID_ARG if K<=0; ids are range-checked on the device and an
 *   out-of-range id sets *status (device int, may be NULL) to 1 and writes zeros.
 */
int vq_lookup(const int64_t* ids, int64_t n, const float* embed, int K, int D,
              float* out, int layout, int B, int A, int C, int* status, vq_stream_t stream);

/*
 * Cross-view cluster loss of EmbeddingLoss._calc_cross_loss (src/functions/embed_loss.py:46-66), the consumer of the
 * quantiser's outputs in the stage-1 trainers (single_window_trainer.py:91-105); SURVEY section 8(f), rank 1.
 *   loss = mean over the (b, k) that occur of  sum_{loc: labels[b, loc] == k + 1} |z[b, :, loc] - embed[k]|^2
 *                                              / (count[b, k] + 1e-6)
 *   labels  int32 [B][H*W]: 0 = no class at this location, k + 1 = class k -- the integer map the reference one-hot
 *           encodes and strips of class 0 (single_window_trainer.py:91-99) before it multiplies by it;
 *   embed   [K][D] row-major (= codebook^T, VQModule.embed; the loss detaches it, embed_loss.py:51);
 *   weights out, float [B*K]: 1 / ((count + 1e-6) * #present) for vq_embed_loss_bwd (0 where (b, k) does not occur);
 *   work    device scratch of vq_embed_loss_work_bytes(B, K) bytes, 256-byte aligned;
 *   vq_embed_loss_bwd: g_z = g_loss * 2 * weights[b, k] * (z - embed[k]) at labelled locations, 0 elsewhere.
 */
size_t vq_embed_loss_work_bytes(int B, int K);
int vq_embed_loss_fwd(const float* z, const int32_t* labels, const float* embed, int B, int D, int H, int W, int K,
                      float* loss, float* weights, void* work, size_t work_bytes, vq_stream_t stream);
int vq_embed_loss_bwd(const float* g_loss, const float* z, const int32_t* labels, const float* embed,
                      const float* weights, float* g_z, int B, int D, int H, int W, int K, vq_stream_t stream);

/*
 * One-hot of an integer label map, channel-major (replaces functions/onehot.py:5-20, the
 * `OneHotEncoder` the stage-1 trainers apply to the code map, single_window_trainer.py:91-99):
 *   labels  int32 or int64 [B, HW] (label_bytes = 4 or 8);  out float [B, C, HW], every element written:
 *   out[b, c, p] = (labels[b, p] == c).  A label outside [0, C) yields an all-zero column.
 */
int vq_onehot(const void* labels, int label_bytes, int64_t B, int64_t HW, int C, float* out, vq_stream_t stream);

/*
 * The two layers in front of the quantiser: `nn.InstanceNorm2d(C)` (no affine parameters, no running
 * statistics, biased variance, eps inside the square root) followed by `nn.ReLU` -- the end of `up_conv1_1.double_conv`
 * (reference blocks.py:39-50), whose output is the z the quantiser reads (vqwnet.py:104-109).
 *   x      [B, C, H, W] fp32 NCHW (the convolution output);  z [B, C, H, W]: relu((x - mean) * rstd), written once;
 *   stats  float [B*C][2] = {mean, rstd} per plane, written by the forward for the backward (may be NULL in inference);
 *   vq_norm_relu_bwd: g_x from g_z, the saved x and stats (the relu mask is recomputed from x exactly as the forward did).
 */
int vq_norm_relu_fwd(const float* x, float* z, float* stats, int B, int C, int H, int W, float eps, vq_stream_t stream);
int vq_norm_relu_bwd(const float* g_z, const float* x, const float* stats, float* g_x, int B, int C, int H, int W,
                     vq_stream_t stream);

/*
 * Measurement hooks (used by bench.py only; they do not change results).
 *   vq_launch_count     number of kernels this library has launched in this process so far.
 *   vq_profile_enable   when on, vq_assign_fwd brackets its dominant kernel (the nearest-code
 *                       search) with CUDA events recorded on the caller's stream.
 *   vq_profile_read     HOST outputs: total milliseconds and number of bracketed launches since
 *                       the last read; synchronises on the recorded events, then resets.
 */
int64_t vq_launch_count(void);
/* Debug: raw tensor-core accumulators (z.e - |e|^2/2, tf32) of the search kernel.
 * out is [B*H*W][vq_debug_tc_ncols(D,K)] floats; returns VQ_ERR_UNSUPPORTED when the shape has no
 * tensor-core path.  vq_debug_fallback_rows: number of rows the last tensor-core vq_assign_fwd on this
 * workspace routed to the exhaustive fp32 search (reads 4 bytes device->host, synchronises). */
int vq_debug_tc_ncols(int D, int K);
int vq_debug_tc_scores(const float* z, int B, int D, int H, int W, const float* embed, int K, float* out,
                       void* workspace, size_t workspace_bytes, vq_stream_t stream);
int vq_debug_fallback_rows(const void* workspace, int64_t N, int K, int D, vq_stream_t stream);
/* Phase timing of the tensor-core epilogue (only in builds with -DVQ_TC_TIMING; returns 0 otherwise): HOST buffer of
 * 148*16*8 int64 clock sums [cta][scan warps 0-7, output warps 8-15][slot]. */
int vq_debug_tc_timing(long long* host_out, int n);
int vq_profile_enable(int on);
int vq_profile_read(double* total_ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* VQ_B200_H_ */

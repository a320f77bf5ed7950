// vq_capi.cu -- the C-ABI (include/vq_b200.h).  Argument checking + kernel sequencing only.
#include <string.h>

#include "vq_common.cuh"

using namespace vqb200;

extern "C" {

int vq_version(void) { return 1000; }

const char* vq_last_error(void) { return get_error(); }

size_t vq_stats_floats(int K, int D) { return stats_sums_offset(K) + (size_t)K * (size_t)D; }

size_t vq_stats_sums_offset(int K) { return stats_sums_offset(K); }

size_t vq_workspace_bytes(int64_t N, int K, int D) {
  if (N < 0 || K <= 0 || D <= 0) return 0;
  return carve_workspace(nullptr, N, K, D).bytes;
}

int vq_assign_path(int B, int D, int H, int W, int K, int flags) {
  if (flags & VQ_FLAG_FORCE_SIMT) return 0;
  if ((flags & VQ_FLAG_NO_STATS) && !(flags & VQ_FLAG_FORCE_TC) && small_path_supported(B, D, H, W, K)) return 2;
  return tc_path_supported(B, D, H, W, K) ? 1 : 0;
}

int vq_assign_fwd(const float* z, int B, int D, int H, int W, const float* embed, int K, int64_t* ids,
                  int32_t* ids_nat, float* q, float* loss, float* stats, float* embed_snapshot, void* workspace,
                  size_t workspace_bytes, int flags, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(B >= 0 && D > 0 && H >= 0 && W >= 0 && K > 0, VQ_ERR_INVALID_ARG,
             "vq_assign_fwd: bad shape B=%d D=%d H=%d W=%d K=%d", B, D, H, W, K);
  VQ_REQUIRE(embed != nullptr && workspace != nullptr, VQ_ERR_INVALID_ARG, "vq_assign_fwd: null embed/workspace");
  VQ_REQUIRE(((uintptr_t)workspace & 255) == 0, VQ_ERR_INVALID_ARG, "vq_assign_fwd: workspace must be 256-byte aligned");
  const int64_t N = (int64_t)B * H * W;
  VQ_REQUIRE(N < (int64_t)1 << 31, VQ_ERR_UNSUPPORTED, "vq_assign_fwd: B*H*W must be < 2^31");
  VQ_REQUIRE((int64_t)D * H * W < (int64_t)1 << 31, VQ_ERR_UNSUPPORTED, "vq_assign_fwd: D*H*W must be < 2^31");
  const size_t need = vq_workspace_bytes(N, K, D);
  VQ_REQUIRE(workspace_bytes >= need, VQ_ERR_WORKSPACE, "vq_assign_fwd: workspace too small (%zu < %zu)",
             workspace_bytes, need);
  VQ_REQUIRE(N == 0 || z != nullptr, VQ_ERR_INVALID_ARG, "vq_assign_fwd: null z");
  if (stats) VQ_REQUIRE(((uintptr_t)stats & 15) == 0, VQ_ERR_INVALID_ARG, "vq_assign_fwd: stats must be 16-byte aligned");

  FwdArgs a;
  a.z = z; a.B = B; a.D = D; a.H = H; a.W = W; a.embed = embed; a.K = K;
  a.ids = ids; a.ids_nat = ids_nat; a.q = q; a.loss = loss; a.stats = stats; a.snapshot = embed_snapshot;
  a.ws = carve_workspace(workspace, N, K, D);
  a.flags = flags;
  cudaStream_t s = (cudaStream_t)stream;

  const bool use_small = stats == nullptr && !(flags & (VQ_FLAG_FORCE_SIMT | VQ_FLAG_FORCE_TC)) &&
                         small_path_supported(B, D, H, W, K);       // inference calls on small codebooks
  bool use_tc = !use_small && !(flags & VQ_FLAG_FORCE_SIMT) && q != nullptr && tc_path_supported(B, D, H, W, K);
  if ((flags & VQ_FLAG_FORCE_TC) && !use_tc) {
    set_error("vq_assign_fwd: tensor-core path unsupported for B=%d D=%d H=%d W=%d K=%d", B, D, H, W, K);
    return VQ_ERR_UNSUPPORTED;
  }
  int rc = launch_prep(a, use_tc, s);
  if (rc) return rc;
  if (N > 0) {
    if (use_tc) {
      rc = launch_assign_tc(a, s);           // tcgen05 approximate search + exact fp32 re-rank
      if (rc) return rc;
      rc = launch_fallback_rows(a, s);       // exhaustive fp32 search of the rows the bound could not decide
      if (rc) return rc;
    } else if (use_small) {
      rc = launch_assign_small(a, s);          // small codebooks: thread = pixel, exact fp32, HBM speed
      if (rc) return rc;
    } else {
      rc = launch_assign_simt(a, false, s);
      if (rc) return rc;
    }
  }
  return launch_finish(a, use_tc, s);
}

int vq_ema_update(float* cluster_size, float* embed_avg, int64_t avg_stride_d, int64_t avg_stride_k, float* embed,
                  const float* stats, int K, int D, double momentum, double eps, float count_scale, float sum_scale,
                  void* scratch, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(K > 0 && D > 0, VQ_ERR_INVALID_ARG, "vq_ema_update: bad shape K=%d D=%d", K, D);
  VQ_REQUIRE(cluster_size && embed_avg && embed && stats && scratch, VQ_ERR_INVALID_ARG, "vq_ema_update: null pointer");
  VQ_REQUIRE(avg_stride_d >= 0 && avg_stride_k >= 0, VQ_ERR_INVALID_ARG, "vq_ema_update: negative embed_avg stride");
  return launch_ema(cluster_size, embed_avg, avg_stride_d, avg_stride_k, embed, stats, K, D, momentum, eps,
                    count_scale, sum_scale, (float*)scratch, (cudaStream_t)stream);
}

int vq_bwd(const float* g_q, const float* g_loss, const float* z, const int32_t* ids_nat, const float* embed_snapshot,
           float* g_z, int B, int D, int H, int W, int K, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(B >= 0 && D > 0 && H >= 0 && W >= 0 && K > 0, VQ_ERR_INVALID_ARG, "vq_bwd: bad shape");
  if ((int64_t)B * H * W == 0) return VQ_OK;
  VQ_REQUIRE(z && ids_nat && embed_snapshot && g_z, VQ_ERR_INVALID_ARG, "vq_bwd: null pointer");
  return launch_bwd(g_q, g_loss, z, ids_nat, embed_snapshot, g_z, B, D, H, W, K, (cudaStream_t)stream);
}

int vq_onehot(const void* labels, int label_bytes, int64_t B, int64_t HW, int C, float* out, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(B >= 0 && HW >= 0 && C > 0, VQ_ERR_INVALID_ARG, "vq_onehot: bad shape B=%lld HW=%lld C=%d", (long long)B, (long long)HW, C);
  VQ_REQUIRE(label_bytes == 4 || label_bytes == 8, VQ_ERR_INVALID_ARG, "vq_onehot: labels must be int32 or int64");
  if (B * HW == 0) return VQ_OK;
  VQ_REQUIRE(labels && out, VQ_ERR_INVALID_ARG, "vq_onehot: null pointer");
  return launch_onehot(labels, label_bytes, B, HW, C, out, (cudaStream_t)stream);
}

size_t vq_embed_loss_work_bytes(int B, int K) {
  if (B <= 0 || K <= 0) return 0;
  return embed_loss_work_bytes(B, K);
}

int vq_embed_loss_fwd(const float* z, const int32_t* labels, const float* embed, int B, int D, int H, int W, int K,
                      float* loss, float* weights, void* work, size_t work_bytes, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(B > 0 && D > 0 && H >= 0 && W >= 0 && K > 0, VQ_ERR_INVALID_ARG,
             "vq_embed_loss_fwd: bad shape B=%d D=%d H=%d W=%d K=%d", B, D, H, W, K);
  VQ_REQUIRE((int64_t)B * K < (int64_t)1 << 31 && (int64_t)B * H * W < (int64_t)1 << 31, VQ_ERR_UNSUPPORTED,
             "vq_embed_loss_fwd: B*K and B*H*W must be < 2^31");
  VQ_REQUIRE(embed && loss && weights && work, VQ_ERR_INVALID_ARG, "vq_embed_loss_fwd: null pointer");
  VQ_REQUIRE((int64_t)B * H * W == 0 || (z && labels), VQ_ERR_INVALID_ARG, "vq_embed_loss_fwd: null z/labels");
  VQ_REQUIRE(((uintptr_t)work & 255) == 0, VQ_ERR_INVALID_ARG, "vq_embed_loss_fwd: work must be 256-byte aligned");
  VQ_REQUIRE(work_bytes >= embed_loss_work_bytes(B, K), VQ_ERR_WORKSPACE, "vq_embed_loss_fwd: work too small (%zu < %zu)",
             work_bytes, embed_loss_work_bytes(B, K));
  return launch_embed_loss_fwd(z, labels, embed, B, D, H, W, K, loss, weights, work, (cudaStream_t)stream);
}

int vq_embed_loss_bwd(const float* g_loss, const float* z, const int32_t* labels, const float* embed,
                      const float* weights, float* g_z, int B, int D, int H, int W, int K, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(B > 0 && D > 0 && H >= 0 && W >= 0 && K > 0, VQ_ERR_INVALID_ARG, "vq_embed_loss_bwd: bad shape");
  if ((int64_t)B * H * W == 0) return VQ_OK;
  VQ_REQUIRE(g_loss && z && labels && embed && weights && g_z, VQ_ERR_INVALID_ARG, "vq_embed_loss_bwd: null pointer");
  return launch_embed_loss_bwd(g_loss, z, labels, embed, weights, g_z, B, D, H, W, K, (cudaStream_t)stream);
}

int vq_lookup(const int64_t* ids, int64_t n, const float* embed, int K, int D, float* out, int layout, int B, int A,
              int C, int* status, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(K > 0 && D > 0 && n >= 0, VQ_ERR_INVALID_ARG, "vq_lookup: bad shape n=%lld K=%d D=%d", (long long)n, K, D);
  VQ_REQUIRE(layout == VQ_LAYOUT_ROWS || layout == VQ_LAYOUT_NCHW_T, VQ_ERR_INVALID_ARG, "vq_lookup: bad layout %d", layout);
  if (n == 0) return VQ_OK;
  VQ_REQUIRE(ids && embed && out, VQ_ERR_INVALID_ARG, "vq_lookup: null pointer");
  if (layout == VQ_LAYOUT_NCHW_T)
    VQ_REQUIRE((int64_t)B * A * C == n && B <= 65535 && (C + 31) / 32 <= 65535, VQ_ERR_INVALID_ARG,
               "vq_lookup: B*A*C != n (or grid too large)");
  return launch_lookup(ids, n, embed, K, D, out, layout, B, A, C, status, (cudaStream_t)stream);
}

int vq_norm_relu_fwd(const float* x, float* z, float* stats, int B, int C, int H, int W, float eps, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0 && eps >= 0.f, VQ_ERR_INVALID_ARG, "vq_norm_relu_fwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  const long long planes = (long long)B * C, HW = (long long)H * W;
  if (planes == 0 || HW == 0) return VQ_OK;
  VQ_REQUIRE(x && z, VQ_ERR_INVALID_ARG, "vq_norm_relu_fwd: null pointer");
  return launch_norm_relu_fwd(x, z, stats, planes, HW, eps, (cudaStream_t)stream);
}

int vq_norm_relu_bwd(const float* g_z, const float* x, const float* stats, float* g_x, int B, int C, int H, int W,
                     vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, VQ_ERR_INVALID_ARG, "vq_norm_relu_bwd: bad shape");
  const long long planes = (long long)B * C, HW = (long long)H * W;
  if (planes == 0 || HW == 0) return VQ_OK;
  VQ_REQUIRE(g_z && x && stats && g_x, VQ_ERR_INVALID_ARG, "vq_norm_relu_bwd: null pointer");
  return launch_norm_relu_bwd(g_z, x, stats, g_x, planes, HW, (cudaStream_t)stream);
}

int vq_debug_tc_ncols(int D, int K) { return tc_debug_ncols(D, K); }

int vq_debug_tc_timing(long long* host_out, int n) { return tc_debug_timing(host_out, n); }

int vq_debug_tc_scores(const float* z, int B, int D, int H, int W, const float* embed, int K, float* out,
                       void* workspace, size_t workspace_bytes, vq_stream_t stream) {
  set_error("");
  VQ_REQUIRE(z && embed && out && workspace, VQ_ERR_INVALID_ARG, "vq_debug_tc_scores: null pointer");
  VQ_REQUIRE(tc_path_supported(B, D, H, W, K), VQ_ERR_UNSUPPORTED, "vq_debug_tc_scores: no tensor-core path for this shape");
  const int64_t N = (int64_t)B * H * W;
  VQ_REQUIRE(workspace_bytes >= vq_workspace_bytes(N, K, D), VQ_ERR_WORKSPACE, "vq_debug_tc_scores: workspace too small");
  FwdArgs a{};
  a.z = z; a.B = B; a.D = D; a.H = H; a.W = W; a.embed = embed; a.K = K;
  a.ws = carve_workspace(workspace, N, K, D);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_prep(a, true, s);
  if (rc) return rc;
  return launch_assign_tc_impl(a, out, s);
}

int vq_debug_fallback_rows(const void* workspace, int64_t N, int K, int D, vq_stream_t stream) {
  Workspace w = carve_workspace((void*)workspace, N, K, D);
  int v = -1;
  if (cudaMemcpyAsync(&v, w.misc, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return -1;
  return v;
}

int64_t vq_launch_count(void) { return (int64_t)launch_count(); }

int vq_profile_enable(int on) {
  profile_enable(on != 0);
  return VQ_OK;
}

int vq_profile_read(double* total_ms, int* launches) { return profile_read(total_ms, launches); }

}  // extern "C"

// vq_assign_tc.cu -- tcgen05/TMEM/TMA distance + argmin kernel (placeholder until the kernel lands).
#include "vq_common.cuh"
namespace vqb200 {
bool tc_path_supported(int, int, int, int, int) { return false; }
int launch_assign_tc(const FwdArgs&, cudaStream_t) {
  set_error("tensor-core path not built");
  return VQ_ERR_UNSUPPORTED;
}
}  // namespace vqb200

#include "gemm_tc.cuh"

namespace v2s {

int gemm_tc_init() { return 0; }

int launch_gemm_tc(const GemmDesc& d, int ta, int tb, int to, cudaStream_t stream, int* handled) {
  *handled = 0;
  return 0;
}

int gemm_tc_test(int which, const void* a, const void* b, void* c, int m, int n, int k, int variant,
                 cudaStream_t stream) {
  set_error("gemm_tc_test: not built yet");
  return 1;
}

}  // namespace v2s

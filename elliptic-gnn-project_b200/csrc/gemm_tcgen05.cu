// K6 (tensor-core path) -- placeholder until the tcgen05/TMA kernel lands; reports
// "unsupported" so egnn_gemm(impl=auto) uses the SIMT path and impl=2 fails loudly.
#include "common.cuh"
namespace egnn {
bool gemm_tcgen05_supported(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, const void*,
                            const void*, const void*) {
  return false;
}
int gemm_tcgen05_dispatch(const void*, int64_t, const void*, int64_t, void*, int, int64_t, int64_t,
                          int64_t, int64_t, const float*, int, cudaStream_t) {
  return fail("egnn_gemm", "tcgen05 path not built");
}
}  // namespace egnn

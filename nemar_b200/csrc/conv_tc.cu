// conv_tc.cu — tcgen05 / TMA implicit-GEMM engine (placeholder until the kernel lands; every query
// answers "unsupported" so the dispatcher never routes here).
#include "common.cuh"
#include "conv_internal.cuh"

bool tc_engine_built() { return false; }
bool tc_gather_supported(const nemar_tensor*, const nemar_tensor*, int, const GatherGeom&) { return false; }
int tc_gather_gemm(const nemar_tensor*, const nemar_tensor*, const void*, int, const float*, int, float*,
                   const GatherGeom&, cudaStream_t) {
  nemar_set_error("tcgen05 engine not built");
  return -1;
}
bool tc_wgrad_supported(const nemar_tensor*, const nemar_tensor*, int, int, int, int) { return false; }
int64_t tc_wgrad_workspace(const nemar_tensor*, const nemar_tensor*, int, int, int, int) { return 0; }
int tc_wgrad(const nemar_tensor*, const nemar_tensor*, float*, int, int, int, int, void*, int64_t, cudaStream_t) {
  nemar_set_error("tcgen05 engine not built");
  return -1;
}

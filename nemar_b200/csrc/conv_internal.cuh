// conv_internal.cuh — shared declarations between the conv dispatch layer and its two engines.
#pragma once
#include "common.cuh"

// Packed-weight layout.  Logical index (o, tap, i) with o < O_p output rows, i < I_p reduction channels:
//   I_p % 16 != 0 : row-major   [o][tap][i]
//   I_p % 16 == 0 : chunk-major [tap][i / BK][o][i % BK], BK = 64 / 32 / 16 (largest dividing I_p) — every TMA weight
//                   box (BK x BN rows) is then ONE contiguous BN*BK*2-byte region instead of BN strided 128-byte rows
__host__ __device__ __forceinline__ int packed_bk(int ip) {
  return (ip % 16 != 0) ? 0 : ((ip % 64 == 0) ? 64 : ((ip % 32 == 0) ? 32 : 16));
}
__host__ __device__ __forceinline__ int64_t packed_index(int o, int tap, int i, int op, int ip, int bk) {
  const int kchunks = ip / bk;
  return (((int64_t)tap * kchunks + i / bk) * op + o) * bk + (i % bk);
}
__host__ __device__ __forceinline__ int64_t packed_index_rm(int o, int tap, int i, int taps, int ip) {
  return ((int64_t)o * taps + tap) * ip + i;
}

struct GatherGeom {
  int kh, kw;
  int sm;          // multiplier applied to the destination coordinate
  int sd;          // divisor applied to the gathered coordinate (stride of a data-gradient pass)
  int pe;          // effective padding (rows): coordinate = dst*sm - pe + tap
  int pe_x;        // effective padding (columns); differs from pe only for rectangular kernels
  int dst_padded;  // 1: iterate over the destination's padded index space (gradient into a halo'd buffer)
};

int generic_pack_multi(const nemar_pack_job* jobs_dev, const int* blocks_dev, int nblocks, cudaStream_t s);

// generic (CUDA-core) engine — conv_generic.cu
int generic_gather_gemm(const nemar_tensor* src, const nemar_tensor* dst, const void* wp, int w_dtype, int wp_cs,
                        const float* bias, int act, const GatherGeom& gg, cudaStream_t s);
int generic_wgrad(const nemar_tensor* x, const nemar_tensor* dy, float* dw, int kh, int kw, int stride, int pe,
                  int accumulate, cudaStream_t s);
int generic_pack(const float* w, void* out, int dtype, int O, int op, int I, int ip, int kh, int kw, int w_is_oi,
                 int flip, cudaStream_t s);

// tcgen05 / TMA engine — conv_tc.cu
bool tc_engine_built();
int tc_set_option(const char* key, int value);   // -> previous value, -1: unknown key
bool tc_gather_supported(const nemar_tensor* src, const nemar_tensor* dst, int wp_cs, const GatherGeom& gg);
int tc_gather_gemm(const nemar_tensor* src, const nemar_tensor* dst, const void* wp, int wp_cs,
                   const float* bias, int act, float* stats, const GatherGeom& gg, cudaStream_t s,
                   float* stats_ws = nullptr, int64_t stats_ws_bytes = 0);
int64_t tc_gather_stats_workspace(const nemar_tensor* src, const nemar_tensor* dst, int wp_cs, const GatherGeom& gg);
bool tc_wgrad_supported(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, int stride, int pe);
int64_t tc_wgrad_workspace(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, int stride, int pe);
int tc_wgrad(const nemar_tensor* x, const nemar_tensor* dy, float* dw, int co_real, int ci_real, int kh, int kw, int stride,
             int pe, void* workspace, int64_t workspace_bytes, int accumulate, cudaStream_t s);

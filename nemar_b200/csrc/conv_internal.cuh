// conv_internal.cuh — shared declarations between the conv dispatch layer and its two engines.
#pragma once
#include "common.cuh"

struct GatherGeom {
  int kh, kw;
  int sm;          // multiplier applied to the destination coordinate
  int sd;          // divisor applied to the gathered coordinate (stride of a data-gradient pass)
  int pe;          // effective padding: coordinate = dst*sm - pe + tap
  int dst_padded;  // 1: iterate over the destination's padded index space (gradient into a halo'd buffer)
};

// generic (CUDA-core) engine — conv_generic.cu
int generic_gather_gemm(const nemar_tensor* src, const nemar_tensor* dst, const void* wp, int w_dtype, int wp_cs,
                        const float* bias, int act, const GatherGeom& gg, cudaStream_t s);
int generic_wgrad(const nemar_tensor* x, const nemar_tensor* dy, float* dw, int kh, int kw, int stride, int pe,
                  cudaStream_t s);
int generic_pack(const float* w, void* out, int dtype, int O, int op, int I, int ip, int kh, int kw, int w_is_oi,
                 int flip, cudaStream_t s);

// tcgen05 / TMA engine — conv_tc.cu
bool tc_engine_built();
bool tc_gather_supported(const nemar_tensor* src, const nemar_tensor* dst, int wp_cs, const GatherGeom& gg);
int tc_gather_gemm(const nemar_tensor* src, const nemar_tensor* dst, const void* wp, int wp_cs,
                   const float* bias, int act, float* stats, const GatherGeom& gg, cudaStream_t s);
bool tc_wgrad_supported(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, int stride, int pe);
int64_t tc_wgrad_workspace(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, int stride, int pe);
int tc_wgrad(const nemar_tensor* x, const nemar_tensor* dy, float* dw, int co_real, int ci_real, int kh, int kw, int stride,
             int pe, void* workspace, int64_t workspace_bytes, cudaStream_t s);

// umma_shift_probe.cu — hardware probe (not product code): can a K-major SWIZZLE_128B UMMA operand START at a row
// that is not a multiple of 8 rows (1024 B), and can its 8-row groups be strided by a non-multiple of 1024 B?
//
// Why: a 3x3 convolution re-fetches its A tile once per tap (9x the L2 -> SM traffic, which is what bounds the
// gather kernel).  If the MMA can read tap-shifted windows out of ONE resident (TH+2) x (TW+2) pixel patch, the
// patch is fetched once.  With 8-pixel-wide tiles the window of tap (dy, dx) starts at row dy*PITCH + dx of the
// patch and its 8-row groups are PITCH rows apart, so the descriptor needs start = base + (dy*PITCH+dx)*128 B and
// SBO = PITCH*128 B.  The probe loads 256 rows x 64 bf16 with TMA (SWIZZLE_128B), then runs M=128 N=64 K=64 MMAs
// with every (shift, SBO, base_offset) combination and reports which ones reproduce the CPU result.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I nemar_b200/csrc -o scripts/probe/umma_shift_probe.bin scripts/probe/umma_shift_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace tc;

struct Cfg { int shift_rows, sbo_bytes, base_offset; };

constexpr int ROWS = 256, KC = 64, NB = 64;

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    const Cfg* cfgs, int ncfg, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // 256 rows x 128 B
  uint8_t* sB = smem + ROWS * 128;          // 64 rows x 128 B
  uint64_t* bars = (uint64_t*)(sB + NB * 128);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], ROWS * 128 + NB * 128);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  constexpr uint32_t idesc = make_idesc_bf16(128, NB, 0, 0);
  for (int c = 0; c < ncfg; ++c) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const Cfg cf = cfgs[c];
      uint64_t adesc = make_smem_desc(smem_u32(sA) + (uint32_t)cf.shift_rows * 128u, 16, (uint32_t)cf.sbo_bytes, LAYOUT_SW128);
      adesc |= (uint64_t)(cf.base_offset & 7) << 49;
      const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 1024, LAYOUT_SW128);
      for (int k = 0; k < KC / 16; ++k) umma_bf16(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k > 0 ? 1u : 0u);
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], (uint32_t)(c & 1));
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int cc = 0; cc < NB; cc += 32) {
      float v[32];
      tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc, v);
      for (int j = 0; j < 32; ++j) out[((size_t)c * 128 + row) * NB + cc + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)p;
  std::vector<__nv_bfloat16> hA(ROWS * KC), hB(NB * KC);
  std::vector<float> fA(ROWS * KC), fB(NB * KC);
  unsigned s = 12345u;
  auto rnd = [&](int lo, int hi) { s = s * 1664525u + 1013904223u; return lo + (int)((s >> 16) % (unsigned)(hi - lo + 1)); };
  for (int i = 0; i < ROWS * KC; ++i) { fA[i] = (float)rnd(-3, 3); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < NB * KC; ++i) { fB[i] = (float)rnd(-2, 2); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {KC, ROWS}; cuuint64_t str[1] = {KC * 2}; cuuint32_t box[2] = {KC, ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
    cuuint64_t dimsb[2] = {KC, NB}; cuuint32_t boxb[2] = {KC, NB};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  std::vector<Cfg> cfgs;
  const int sbos[3] = {1024, 1280, 2048};
  for (int si = 0; si < 3; ++si)
    for (int sh = 0; sh <= 8; ++sh) {
      cfgs.push_back({sh, sbos[si], 0});
      if (sh & 7) cfgs.push_back({sh, sbos[si], sh & 7});
    }
  Cfg* dC; float* dO;
  CK(cudaMalloc(&dC, cfgs.size() * sizeof(Cfg)));
  CK(cudaMemcpy(dC, cfgs.data(), cfgs.size() * sizeof(Cfg), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dO, cfgs.size() * 128 * NB * sizeof(float)));
  const int smem = ROWS * 128 + NB * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(tmA, tmB, dC, (int)cfgs.size(), dO);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hO(cfgs.size() * 128 * NB);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  printf("UMMA K-major SWIZZLE_128B, M=128 N=64 K=64: A window = rows shift + (m/8)*(SBO/128) + m%%8 of a 256-row TMA tile\n");
  for (size_t c = 0; c < cfgs.size(); ++c) {
    const Cfg cf = cfgs[c];
    int good_rows = 0, first_bad = -1;
    for (int m = 0; m < 128; ++m) {
      const int r = cf.shift_rows + (m / 8) * (cf.sbo_bytes / 128) + (m % 8);
      bool ok = true;
      for (int n = 0; n < NB && ok; ++n) {
        float acc = 0.f;
        for (int k = 0; k < KC; ++k) acc += fA[r * KC + k] * fB[n * KC + k];
        ok = (acc == hO[(c * 128 + m) * NB + n]);
      }
      if (ok) ++good_rows; else if (first_bad < 0) first_bad = m;
    }
    printf("PROBE shift=%d sbo=%d base_offset=%d : %s (%d/128 rows exact, first bad row %d)\n", cf.shift_rows, cf.sbo_bytes,
           cf.base_offset, good_rows == 128 ? "MATCH" : "mismatch", good_rows, first_bad);
  }
  return 0;
}

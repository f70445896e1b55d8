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

constexpr int ROWS = 256, NB = 64;

template <int KC>
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    const Cfg* cfgs, int ncfg, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t RB = KC * 2;           // bytes per row: 128 / 64 / 32 -> SWIZZLE_128B / 64B / 32B
  constexpr uint64_t LAYOUT = KC == 64 ? LAYOUT_SW128 : (KC == 32 ? LAYOUT_SW64 : LAYOUT_SW32);
  uint8_t* sA = smem;                       // 256 rows
  uint8_t* sB = smem + ROWS * 128;          // 64 rows
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
    mbar_expect_tx(&bars[0], ROWS * RB + NB * RB);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  constexpr uint32_t idesc = make_idesc_bf16(128, NB, 0, 0);
  for (int c = 0; c < ncfg; ++c) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const Cfg cf = cfgs[c];
      uint64_t adesc = make_smem_desc(smem_u32(sA) + (uint32_t)cf.shift_rows * RB, 16, (uint32_t)cf.sbo_bytes, LAYOUT);
      adesc |= (uint64_t)(cf.base_offset & 7) << 49;
      const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 8 * RB, LAYOUT);
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

// ---- second probe: MN-major A operand (rows = K, the swizzled row holds CW consecutive M elements) -------------------
// D[m][n] = sum_k A[k][m] * B[n][k], M = 128 = (128 / CW) chunks of CW elements, chunk stride LBO, K = 64 = 8 groups of 8
// rows, group stride SBO.  A weight gradient over a resident patch needs: a shifted start row, SBO = PW rows, and LBO
// = ONE row (the chunks are the dx-adjacent taps of the same patch, i.e. overlapping windows).
struct CfgMN { int shift_rows, group_rows, lbo_rows; };

template <int CW>
__global__ void __launch_bounds__(128) probe_mn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                       const CfgMN* cfgs, int ncfg, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t RB = CW * 2;
  constexpr uint64_t LAYOUT = CW == 64 ? LAYOUT_SW128 : (CW == 32 ? LAYOUT_SW64 : LAYOUT_SW32);
  uint8_t* sA = smem;                       // 256 rows (K pool) x CW elements
  uint8_t* sB = smem + ROWS * 128;          // 64 rows (N) x 64 k, K-major SWIZZLE_128B
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
    mbar_expect_tx(&bars[0], ROWS * RB + NB * 128);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  constexpr uint32_t idesc = make_idesc_bf16(128, NB, 1, 0);      // A MN-major, B K-major
  for (int c = 0; c < ncfg; ++c) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const CfgMN cf = cfgs[c];
      const uint32_t sbo = (uint32_t)cf.group_rows * RB, lbo = (uint32_t)cf.lbo_rows * RB;
      const uint64_t adesc = make_smem_desc(smem_u32(sA) + (uint32_t)cf.shift_rows * RB, lbo, sbo, LAYOUT);
      const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 1024, LAYOUT_SW128);
      for (int k = 0; k < 4; ++k)        // 16 K rows per MMA = two 8-row groups
        umma_bf16(tmem, adesc + (uint64_t)((k * 2 * sbo) >> 4), bdesc + (uint64_t)(k * 2), idesc, k > 0 ? 1u : 0u);
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

template <int KC>
static int run(EncodeTiledFn enc) {
  constexpr int RB = KC * 2;
  std::vector<__nv_bfloat16> hA(ROWS * KC), hB(NB * KC);
  std::vector<float> fA(ROWS * KC), fB(NB * KC);
  unsigned s = 12345u + KC;
  auto rnd = [&](int lo, int hi) { s = s * 1664525u + 1013904223u; return lo + (int)((s >> 16) % (unsigned)(hi - lo + 1)); };
  for (int i = 0; i < ROWS * KC; ++i) { fA[i] = (float)rnd(-3, 3); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < NB * KC; ++i) { fB[i] = (float)rnd(-2, 2); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {KC, ROWS}; cuuint64_t str[1] = {KC * 2}; cuuint32_t box[2] = {KC, ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
    cuuint64_t dimsb[2] = {KC, NB}; cuuint32_t boxb[2] = {KC, NB};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  std::vector<Cfg> cfgs;
  const int group_rows[3] = {8, 10, 16};      // rows between consecutive 8-row groups (SBO / row bytes)
  for (int si = 0; si < 3; ++si)
    for (int sh = 0; sh <= 11; ++sh) {
      if (sh + 15 * group_rows[si] + 7 >= ROWS) continue;
      cfgs.push_back({sh, group_rows[si] * RB, 0});
      if (KC == 64 && (sh & 7) && sh < 8) cfgs.push_back({sh, group_rows[si] * RB, sh & 7});
    }
  Cfg* dC; float* dO;
  CK(cudaMalloc(&dC, cfgs.size() * sizeof(Cfg)));
  CK(cudaMemcpy(dC, cfgs.data(), cfgs.size() * sizeof(Cfg), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dO, cfgs.size() * 128 * NB * sizeof(float)));
  const int smem = ROWS * 128 + NB * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<KC><<<1, 128, smem>>>(tmA, tmB, dC, (int)cfgs.size(), dO);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hO(cfgs.size() * 128 * NB);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  printf("UMMA K-major SWIZZLE_%dB, M=128 N=64 K=%d: A window = rows shift + (m/8)*(SBO/%d) + m%%8 of a 256-row TMA tile\n", RB, KC, RB);
  int nmatch = 0;
  for (size_t c = 0; c < cfgs.size(); ++c) {
    const Cfg cf = cfgs[c];
    int good_rows = 0, first_bad = -1;
    for (int m = 0; m < 128; ++m) {
      const int r = cf.shift_rows + (m / 8) * (cf.sbo_bytes / RB) + (m % 8);
      bool ok = true;
      for (int n = 0; n < NB && ok; ++n) {
        float acc = 0.f;
        for (int k = 0; k < KC; ++k) acc += fA[r * KC + k] * fB[n * KC + k];
        ok = (acc == hO[(c * 128 + m) * NB + n]);
      }
      if (ok) ++good_rows; else if (first_bad < 0) first_bad = m;
    }
    nmatch += good_rows == 128;
    printf("PROBE sw=%dB shift=%d sbo=%d base_offset=%d : %s (%d/128 rows exact, first bad row %d)\n", RB, cf.shift_rows, cf.sbo_bytes,
           cf.base_offset, good_rows == 128 ? "MATCH" : "mismatch", good_rows, first_bad);
  }
  printf("SUMMARY sw=%dB: %d of %d configurations exact\n", RB, nmatch, (int)cfgs.size());
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dO);
  return 0;
}

template <int CW>
static int run_mn(EncodeTiledFn enc) {
  constexpr int RB = CW * 2, KT = 64;
  std::vector<__nv_bfloat16> hA(ROWS * CW), hB(NB * KT);
  std::vector<float> fA(ROWS * CW), fB(NB * KT);
  unsigned s = 777u + CW;
  auto rnd = [&](int lo, int hi) { s = s * 1664525u + 1013904223u; return lo + (int)((s >> 16) % (unsigned)(hi - lo + 1)); };
  for (int i = 0; i < ROWS * CW; ++i) { fA[i] = (float)rnd(-3, 3); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < NB * KT; ++i) { fB[i] = (float)rnd(-2, 2); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const CUtensorMapSwizzle sw = CW == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (CW == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {CW, ROWS}; cuuint64_t str[1] = {CW * 2}; cuuint32_t box[2] = {CW, ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
    cuuint64_t dimsb[2] = {KT, NB}; cuuint64_t strb[1] = {KT * 2}; cuuint32_t boxb[2] = {KT, NB};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, strb, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  constexpr int NCH = 128 / CW;
  std::vector<CfgMN> cfgs;
  const int shifts[4] = {0, 1, 3, 5}, groups[2] = {8, 10}, lbos[3] = {1, 2, 48};
  for (int gi = 0; gi < 2; ++gi)
    for (int li = 0; li < 3; ++li)
      for (int si = 0; si < 4; ++si) {
        if (shifts[si] + (NCH - 1) * lbos[li] + 7 * groups[gi] + 7 >= ROWS) continue;
        cfgs.push_back({shifts[si], groups[gi], lbos[li]});
      }
  CfgMN* dC; float* dO;
  CK(cudaMalloc(&dC, cfgs.size() * sizeof(CfgMN)));
  CK(cudaMemcpy(dC, cfgs.data(), cfgs.size() * sizeof(CfgMN), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dO, cfgs.size() * 128 * NB * sizeof(float)));
  const int smem = ROWS * 128 + NB * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_mn_kernel<CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_mn_kernel<CW><<<1, 128, smem>>>(tmA, tmB, dC, (int)cfgs.size(), dO);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hO(cfgs.size() * 128 * NB);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  printf("UMMA MN-major A, SWIZZLE_%dB (%d M-chunks of %d), M=128 N=64 K=64: row(k, chunk) = shift + chunk*LBO_rows + (k/8)*group_rows + k%%8\n", RB, NCH, CW);
  int nmatch = 0;
  for (size_t c = 0; c < cfgs.size(); ++c) {
    const CfgMN cf = cfgs[c];
    int good_rows = 0, first_bad = -1;
    for (int m = 0; m < 128; ++m) {
      const int ch = m / CW, j = m % CW;
      bool ok = true;
      for (int n = 0; n < NB && ok; ++n) {
        float acc = 0.f;
        for (int k = 0; k < KT; ++k) {
          const int r = cf.shift_rows + ch * cf.lbo_rows + (k / 8) * cf.group_rows + (k % 8);
          acc += fA[r * CW + j] * fB[n * KT + k];
        }
        ok = (acc == hO[(c * 128 + m) * NB + n]);
      }
      if (ok) ++good_rows; else if (first_bad < 0) first_bad = m;
    }
    nmatch += good_rows == 128;
    printf("PROBE-MN sw=%dB shift=%d group_rows=%d lbo_rows=%d : %s (%d/128 rows exact, first bad row %d)\n", RB, cf.shift_rows,
           cf.group_rows, cf.lbo_rows, good_rows == 128 ? "MATCH" : "mismatch", good_rows, first_bad);
  }
  printf("SUMMARY-MN sw=%dB: %d of %d configurations exact\n", RB, nmatch, (int)cfgs.size());
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dO);
  return 0;
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)p;
  if (run<64>(enc)) return 1;
  if (run<32>(enc)) return 1;
  if (run<16>(enc)) return 1;
  if (run_mn<64>(enc)) return 1;
  if (run_mn<32>(enc)) return 1;
  return 0;
}

// conv_tc.cu — tcgen05 / TMA implicit-GEMM convolution engine for sm_100a (bf16 operands, fp32 accumulation
// in TMEM).  Two kernels:
//
//  gather kernel (Conv2d fprop, Conv2d dgrad, ConvTranspose2d fprop/dgrad):
//      D[pixel][c_out] = sum_{tap, c_in} SRC[pixel @ tap][c_in] * Wp[c_out][tap][c_in]
//    M = 128 destination pixels (a TW x TH x TN box of one or more images), N = BN channels, K = taps * C_in.
//    A tiles are fetched by ONE 4-D TMA box per (tap, BK-channel chunk): the box start is shifted by the tap
//    offset, out-of-bounds rows are zero-filled by TMA (zero padding and tile overhang for free), reflect halos
//    are materialised by the producer pass so they are ordinary data, stride-2 convolutions use the TMA
//    traversal stride, and stride-2 data gradients / transposed convolutions are decomposed into the four
//    output-parity classes, each a stride-1 gather over a subset of taps.  Both operands land in shared memory
//    in the canonical K-major swizzled layout consumed directly by tcgen05.mma (cta_group::1, M=128).
//    BK = 64 / 32 / 16 channels per k-step (SWIZZLE_128B / 64B / 32B) covers every channel count of the path:
//    256/128/64, the 32- and 96-channel full-resolution STN layers, and 3/6-channel images padded to 16.
//
//  wgrad kernel:  dW[c_out][tap][c_in] = sum_pixels dY[pixel][c_out] * X[pixel @ tap][c_in]
//    M = 128 output channels, N = BN input channels, K = pixels; both operands are the same NHWC boxes, used
//    MN-major; split-K over pixel ranges with fp32 partials reduced by a finalize kernel.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-5 = epilogue
// (TMEM -> registers -> global).  smem ring of STAGES slots guarded by full/empty mbarriers; tcgen05.commit
// releases slots and publishes the accumulator.
//
// Variants in this file (same roles, same barrier protocol):
//   tc_gather_pair_kernel / tc_wgrad_pair_kernel — CTA pairs (tcgen05 cta_group::2): one M=256 x N=256 MMA per k-step
//       over the two SMs of a TPC, each CTA staging its own 128 rows of A and its half of B (bit-identical results;
//       the wgrad pairs are the default for 256-channel layers, NEMAR_TC_PAIR selects);
//   tc_rp3_kernel — resident-patch kernel for stride-1 k x k layers with <= 64 output channels: persistent CTAs,
//       weights resident in shared memory, one TMA patch per tile, tap-shifted UMMA windows (opt-in NEMAR_TC_RP3=1).
#include "common.cuh"
#include "conv_internal.cuh"
#include "tc_common.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

using namespace tc;

namespace {

constexpr int BM = 128;         // UMMA M (TMEM lanes)
constexpr int MAX_TAPS = 49;
constexpr int NTHREADS = 192;

__host__ __device__ constexpr uint32_t round1k(uint32_t v) { return (v + 1023u) & ~1023u; }
__host__ __device__ constexpr uint64_t layout_for(int chunk_elems) {
  return chunk_elems == 64 ? LAYOUT_SW128 : (chunk_elems == 32 ? LAYOUT_SW64 : LAYOUT_SW32);
}
static int sm_count() {
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
  }();
  return n;
}

static CUtensorMapSwizzle swizzle_for(int chunk_elems) {
  return chunk_elems == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (chunk_elems == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
static int chunk_for(int c) { return (c % 64 == 0) ? 64 : ((c % 32 == 0) ? 32 : 16); }

// -------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time dependency on libcuda)
// -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 4-D NHWC activation map: dims (C, X, Y, N) of a view read in padded coordinates; box (bc, bx*es, by*es, bn)
static int make_act_map(CUtensorMap* m, const nemar_tensor* t, int box_c, int bx, int by, int bn, int es) {
  EncodeTiledFn enc = get_encode();
  NEMAR_REQUIRE(enc, "cuTensorMapEncodeTiled unavailable");
  const int hp = t->h + 2 * t->pad, wp = t->w + 2 * t->pad;
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)wp, (cuuint64_t)hp, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->cs * 2, (cuuint64_t)wp * t->cs * 2, (cuuint64_t)hp * wp * t->cs * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(bx * es), (cuuint32_t)(by * es), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  void* base = (char*)t->ptr + (size_t)t->coff * 2;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for(box_c), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NEMAR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: %d (c=%d cs=%d wp=%d hp=%d n=%d box=%d,%d,%d,%d es=%d)",
                (int)r, t->c, t->cs, wp, hp, t->n, box_c, bx, by, bn, es);
  return 0;
}

// 3-D weight map over the chunk-major pack [tap*kchunks][rows][bk]: a box (bk, box_rows, 1) is one contiguous region
static int make_w_map(CUtensorMap* m, const void* w, int rows, int k_total, int bk, int box_rows) {
  EncodeTiledFn enc = get_encode();
  NEMAR_REQUIRE(enc, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)bk, (cuuint64_t)rows, (cuuint64_t)(k_total / bk)};
  cuuint64_t strides[2] = {(cuuint64_t)bk * 2, (cuuint64_t)rows * bk * 2};
  cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NEMAR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return 0;
}

// -------------------------------------------------------------------------------------------------
// gather kernel
// -------------------------------------------------------------------------------------------------
struct GatherParams {
  int ntaps, kchunks;            // K loop = ntaps * kchunks steps of BK channels
  short tdy[MAX_TAPS], tdx[MAX_TAPS], twi[MAX_TAPS];   // tap offsets (source coords) and weight tap index
  int cs;                        // source channels (weight row = taps_total * cs, tap-major)
  int tw, th, tn;                // tile box (tw*th*tn == 128)
  int tiles_x, tiles_y, tiles_n;
  int sm;                        // source coordinate multiplier of the box start
  int dw, dh, dn;                // extent of the destination index space covered by this launch
  int ostep, oy0, ox0;           // destination coordinate = t * ostep + o0  (parity classes)
  long long ds_n, ds_y, ds_x;    // destination strides (elements)
  void* dst;                     // destination base (channel offset applied); bf16 or fp32
  int cd;                        // destination channels
  const float* bias;
  int act;
  float* stats;                  // optional [tiles of the whole problem][cd][2]: per-TILE column sums / sums of squares of
                                 // the fp32 pre-activation (plain stores, no atomics; stats_finalize_kernel adds them up)
  int stat_tile0;                // index of this launch's (parity class's) first tile in that buffer
  int stages;                    // ring depth (<= GatherCfg::STAGES)
  int tpc;                       // destination tiles per CTA (each with its own TMEM accumulator)
  uint32_t tmem_cols;            // power of two >= tpc * accumulator stride
  int tile0, tile_end;           // tile range of this launch (tc_gather_kernel; set by launch_gather_t)
};

constexpr int MAX_TPC = 8;

// Sum each of a lane's 32 values across the 32 lanes of the warp with 31 shuffles: on return a[0] of lane L holds
// the warp-wide total of original index L (recursive halving: at offset o a lane keeps the half selected by bit o).
__device__ __forceinline__ void warp_transpose_sum(float (&a)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = upper ? a[i] : a[i + o];
      const float keep = upper ? a[i + o] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

// ---- epilogue --------------------------------------------------------------------------------------------------
// One 32-column chunk of a thread's accumulator row: bias from SHARED memory (every lane reads the same addresses:
// broadcast), the activation resolved ONCE per chunk (a per-element switch on a run-time value compiles to four
// branches per element behind a dependent global bias load: ~15 SASS instructions and ~80 stall cycles per output
// element, measured with ncu — the round-1 epilogue cost as much as the whole K loop of a 256-channel tile), then
// 16-byte stores of the first `ncols` columns (a multiple of 8).
template <bool F32OUT, bool STATS>
__device__ __forceinline__ void epi_chunk(const uint32_t (&r)[32], const float* __restrict__ sb, int act, void* dst,
                                          int ncols, bool valid, float* sst, int lane) {
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(sb + j);
    f[j] = __uint_as_float(r[j]) + b.x;
    f[j + 1] = __uint_as_float(r[j + 1]) + b.y;
    f[j + 2] = __uint_as_float(r[j + 2]) + b.z;
    f[j + 3] = __uint_as_float(r[j + 3]) + b.w;
  }
  if (act == NEMAR_ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.2f * f[j]);      // == x > 0 ? x : 0.2 x
  } else if (act == NEMAR_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  } else if (act == NEMAR_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = tanhf(f[j]);
  }
  if constexpr (STATS) {
    // InstanceNorm statistics of the fp32 pre-activation (act == NONE on these layers): column sums over this warp's
    // 32 pixels by a transposing butterfly (31 shuffles per statistic), left in shared memory for the per-tile combine
    float sv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) sv[j] = valid ? f[j] : 0.f;
    warp_transpose_sum(sv, lane);
    sst[lane * 2] = sv[0];
#pragma unroll
    for (int j = 0; j < 32; ++j) sv[j] = valid ? f[j] * f[j] : 0.f;
    warp_transpose_sum(sv, lane);
    sst[lane * 2 + 1] = sv[0];
  }
  if (!valid) return;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g * 8 < ncols) {
      if constexpr (F32OUT) {
        float* o = (float*)dst + g * 8;
        *reinterpret_cast<float4*>(o) = make_float4(f[g * 8], f[g * 8 + 1], f[g * 8 + 2], f[g * 8 + 3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(f[g * 8 + 4], f[g * 8 + 5], f[g * 8 + 6], f[g * 8 + 7]);
      } else {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[g * 8 + 2 * j], f[g * 8 + 2 * j + 1]);
        *reinterpret_cast<uint4*>((__nv_bfloat16*)dst + g * 8) = pk;
      }
    }
  }
}

// The whole accumulator row of one tile (NCH chunks of 32 columns starting at TMEM address `taddr`): the load of chunk
// c + 1 is in flight while chunk c is converted and stored.  `off` = element offset of the row's first column in dst;
// `cols` = number of real columns (multiple of 8; columns beyond it are not stored).
template <int NCH, bool F32OUT, bool STATS = false>
__device__ __forceinline__ void epi_row(uint32_t taddr, const float* __restrict__ sbias, int act, void* dst, long long off,
                                        int cols, bool valid, float* sstat = nullptr, int lane = 0) {
  uint32_t ra[32], rb[32];
  tmem_ld_32x32_issue(taddr, ra);
#pragma unroll
  for (int ch = 0; ch < NCH; ch += 2) {
    tmem_ld_wait(ra);
    if (ch + 1 < NCH) tmem_ld_32x32_issue(taddr + (uint32_t)(ch + 1) * 32u, rb);
    {
      void* d = F32OUT ? (void*)((float*)dst + off + ch * 32) : (void*)((__nv_bfloat16*)dst + off + ch * 32);
      epi_chunk<F32OUT, STATS>(ra, sbias + ch * 32, act, d, cols - ch * 32, valid, sstat + ch * 64, lane);
    }
    if (ch + 1 < NCH) {
      tmem_ld_wait(rb);
      if (ch + 2 < NCH) tmem_ld_32x32_issue(taddr + (uint32_t)(ch + 2) * 32u, ra);
      void* d = F32OUT ? (void*)((float*)dst + off + (ch + 1) * 32) : (void*)((__nv_bfloat16*)dst + off + (ch + 1) * 32);
      epi_chunk<F32OUT, STATS>(rb, sbias + (ch + 1) * 32, act, d, cols - (ch + 1) * 32, valid, sstat + (ch + 1) * 64, lane);
    }
  }
}

// bias of this CTA's channel tile into shared memory (zeros where there is no bias / beyond the destination channels)
__device__ __forceinline__ void load_bias_tile(float* sbias, const float* __restrict__ bias, int c0, int cd, int bn) {
  for (int k = threadIdx.x; k < bn; k += blockDim.x) sbias[k] = (bias && c0 + k < cd) ? __ldg(bias + c0 + k) : 0.f;
}

template <int BN, int BK>
struct GatherCfg {
  static constexpr uint32_t A_BYTES = round1k(BM * BK * 2), B_BYTES = round1k(BN * BK * 2);
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // BN <= 128: ~96 KB of stages so that two CTAs share an SM (one's epilogue overlaps the other's main loop);
  // BN == 256: one CTA per SM with a deep ring (halves the A re-reads per output channel)
  static constexpr int STAGES_RAW = (int)((BN == 256 ? 196608u : 98304u) / STAGE_BYTES);
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr uint32_t TX_BYTES = BM * BK * 2 + BN * BK * 2;
  static constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;     // TMEM columns of one accumulator
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + (size_t)ACC_COLS * sizeof(float) +
                                 (size_t)4 * ACC_COLS * 2 * sizeof(float);     // bias tile + per-warp statistics staging
};

template <int BN, int BK, bool F32OUT>
__global__ void __launch_bounds__(NTHREADS)
tc_gather_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ GatherParams P) {
  using Cfg = GatherCfg<BN, BK>;
  const int STAGES = P.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;     // one per accumulator
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + MAX_TPC);
  float* sbias = (float*)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);          // [ACC_COLS] bias of this CTA's channel tile
  float* sstat = sbias + Cfg::ACC_COLS;            // [4 epilogue warps][ACC_COLS][2] statistics of the current tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA owns the consecutive destination tiles [t_first, t_first + t_count): one K pipeline runs through all
  // of them while the epilogue warps drain accumulator i during the main loop of tile i + 1
  const int tiles_total = P.tile_end;
  const int t_first = P.tile0 + blockIdx.x * P.tpc;
  const int t_count = (tiles_total - t_first < P.tpc) ? (tiles_total - t_first) : P.tpc;
  const int c0 = blockIdx.y * BN;
  const int ksteps = P.ntaps * P.kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < MAX_TPC; ++i) mbar_init(&tmem_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, P.tmem_cols);
  load_bias_tile(sbias, P.bias, c0, P.cd, (int)Cfg::ACC_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps stay CONVERGED and elect one lane per issue (elect.sync): inside a plain `lane == 0` branch
  // ptxas cannot treat the TMA / tcgen05 operands as warp-uniform and wraps every UTMALDG / UTCHMMA / UTCBAR in a
  // "waterfall" loop (ELECT, R2UR.BROADCAST, BRA.U.ANY) — measured 219 clk per MMA whatever its N, i.e. the tensor pipe
  // capped at 29 % per CTA for N = 128 (scripts/probe/umma_rate_probe.cu, profiles/r02_umma_rate_probe.txt).
  if (warp == 0) {
    // ===== TMA producer =====
    int stage = 0; uint32_t phase = 0;
    // (rotating each CTA's K-loop start so that concurrent CTAs fetch different weight tiles was measured: no effect)
    for (int ti = 0; ti < t_count; ++ti) {
      int t = t_first + ti;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int x0 = tx * P.tw, y0 = ty * P.th, n0 = (t / P.tiles_y) * P.tn;
      for (int ks = 0; ks < ksteps; ++ks) {
        const int tap = ks / P.kchunks, kc = ks - tap * P.kchunks;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::TX_BYTES);
          tma_load_4d(sa, &tmA, &full_bar[stage], kc * BK, x0 * P.sm + P.tdx[tap], y0 * P.sm + P.tdy[tap], n0);
          tma_load_3d(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], 0, c0, P.twi[tap] * P.kchunks + kc);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t SBO = 8 * BK * 2;     // 8 rows of one swizzle atom
    int stage = 0; uint32_t phase = 0;
    for (int ti = 0; ti < t_count; ++ti) {
      const uint32_t acc = tmem_base + (uint32_t)ti * Cfg::ACC_COLS;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa, 16, SBO, layout_for(BK));
          const uint64_t bdesc = make_smem_desc(sa + Cfg::A_BYTES, 16, SBO, layout_for(BK));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)   // +32 bytes per UMMA_K inside the swizzled row
            umma_bf16(acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (ks > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (ks == ksteps - 1) umma_commit(&tmem_full[ti]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: warps 2..5; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rx = row % P.tw, ry = (row / P.tw) % P.th, rn = row / (P.tw * P.th);
    int cols = P.cd - c0;                                  // real columns of this channel tile (multiple of 8)
    if (cols > BN) cols = BN;
#pragma unroll 1
    for (int ti = 0; ti < t_count; ++ti) {
      int t = t_first + ti;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int x0 = tx * P.tw, y0 = ty * P.th, n0 = (t / P.tiles_y) * P.tn;
      const uint32_t acc = tmem_base + (uint32_t)ti * Cfg::ACC_COLS;
      const int px = x0 + rx, py = y0 + ry, pn = n0 + rn;
      const bool valid = px < P.dw && py < P.dh && pn < P.dn;
      const long long off = (long long)pn * P.ds_n + (long long)(py * P.ostep + P.oy0) * P.ds_y +
                            (long long)(px * P.ostep + P.ox0) * P.ds_x + c0;
      mbar_wait(&tmem_full[ti], 0);
      tc_fence_after();
      if (P.stats) {
        epi_row<(int)Cfg::ACC_COLS / 32, F32OUT, true>(acc + ((uint32_t)(q * 32) << 16), sbias, P.act, P.dst, off, cols, valid,
                                                       sstat + q * (int)Cfg::ACC_COLS * 2, lane);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // combine the four warps' column sums: one plain float2 store per channel of this tile
        float2* gp = reinterpret_cast<float2*>(P.stats) + ((long long)(P.stat_tile0 + t_first + ti) * P.cd + c0);
        for (int col = threadIdx.x - 64; col < cols; col += 128) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            a += sstat[(w4 * (int)Cfg::ACC_COLS + col) * 2];
            b += sstat[(w4 * (int)Cfg::ACC_COLS + col) * 2 + 1];
          }
          gp[col] = make_float2(a, b);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");       // the next tile overwrites the staging area
      } else {
        epi_row<(int)Cfg::ACC_COLS / 32, F32OUT>(acc + ((uint32_t)(q * 32) << 16), sbias, P.act, P.dst, off, cols, valid);
      }
    }   // tiles of this CTA
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// -------------------------------------------------------------------------------------------------
// gather kernel, CTA-pair variant (cta_group::2): destination tiles of 256 pixels x 256 channels per pair
// -------------------------------------------------------------------------------------------------
// The single-CTA kernel fetches 32 KB (A 128x64 + B 128x64) per 128x128x64 MMA step and is bound by the chip-wide
// L2 -> SM TMA throughput (~6.3 KB/clk), not by the tensor pipe.  Here two CTAs of a cluster (the two SMs of a TPC)
// run ONE M=256 x N=256 MMA per k-step: each CTA stages its own 128 pixels of A and its own 128 of the 256 weight
// rows, i.e. 32 KB per 128x256x64 of work per CTA — twice the flops per fetched byte.  Protocol:
//   * both producers issue TMA into their own ring; every byte is counted on the LEADER's full barrier (armed by
//     the leader with the bytes of both CTAs);
//   * the leader's elected thread issues tcgen05.mma.cta_group::2 (reads both CTAs' shared memory at the same
//     offsets, writes 128 accumulator lanes into each CTA's TMEM) and releases the slot in BOTH CTAs with a
//     multicast tcgen05.commit; the last commit of a tile publishes the accumulator to both epilogues;
//   * each CTA's epilogue warps drain their own 128 pixels x 256 channels.
constexpr int PAIR_BN = 256, PAIR_BK = 64, PAIR_MAX_STAGES = 6, PAIR_MAX_TPC = 2;
constexpr uint32_t PAIR_A_BYTES = BM * PAIR_BK * 2, PAIR_B_BYTES = (PAIR_BN / 2) * PAIR_BK * 2;
constexpr uint32_t PAIR_STAGE_BYTES = PAIR_A_BYTES + PAIR_B_BYTES;
static size_t pair_smem_bytes(int stages) { return (size_t)stages * PAIR_STAGE_BYTES + 1024 + 256 + PAIR_BN * sizeof(float); }

__global__ void __launch_bounds__(NTHREADS)
tc_gather_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ GatherParams P) {
  const int STAGES = P.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * PAIR_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + PAIR_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + PAIR_MAX_STAGES;     // one per accumulator
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + PAIR_MAX_TPC);
  float* sbias = (float*)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);                // [PAIR_BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = leader (cluster = blocks 2p, 2p+1 along x)
  // the pair owns the consecutive tile pairs [p_first, p_first + p_count); tile pair j = tiles 2j (leader), 2j+1
  const int tiles_total = P.tiles_x * P.tiles_y * P.tiles_n;
  const int pairs_total = (tiles_total + 1) >> 1;
  const int p_first = (int)(blockIdx.x >> 1) * P.tpc;
  const int p_count = (pairs_total - p_first < P.tpc) ? (pairs_total - p_first) : P.tpc;
  const int c0 = blockIdx.y * PAIR_BN;
  const int ksteps = P.ntaps * P.kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < PAIR_MAX_TPC; ++i) mbar_init(&tmem_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, P.tmem_cols);
  load_bias_tile(sbias, P.bias, c0, P.cd, PAIR_BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before anything of ours can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs); converged warp, one elected lane issues (see tc_gather_kernel) =====
    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[0]), 0);
    int stage = 0; uint32_t phase = 0;
    for (int ti = 0; ti < p_count; ++ti) {
      int t = (p_first + ti) * 2 + (int)rank;      // t == tiles_total (odd count): every row out of bounds -> zeros
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int x0 = tx * P.tw, y0 = ty * P.th, n0 = (t / P.tiles_y) * P.tn;
      for (int ks = 0; ks < ksteps; ++ks) {
        const int tap = ks / P.kchunks, kc = ks - tap * P.kchunks;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * PAIR_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * PAIR_STAGE_BYTES);
          const uint32_t fb = full_leader + (uint32_t)stage * 8u;
          tma_load_4d_pair(sa, &tmA, fb, kc * PAIR_BK, x0 * P.sm + P.tdx[tap], y0 * P.sm + P.tdy[tap], n0);
          tma_load_3d_pair(sa + PAIR_A_BYTES, &tmB, fb, 0, c0 + (int)rank * (PAIR_BN / 2), P.twi[tap] * P.kchunks + kc);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===== MMA issuer (leader CTA only) =====
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, PAIR_BN, 0, 0);
      constexpr uint32_t SBO = 8 * PAIR_BK * 2;
      int stage = 0; uint32_t phase = 0;
      for (int ti = 0; ti < p_count; ++ti) {
        const uint32_t acc = tmem_base + (uint32_t)ti * PAIR_BN;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * PAIR_STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(sa, 16, SBO, LAYOUT_SW128);
            const uint64_t bdesc = make_smem_desc(sa + PAIR_A_BYTES, 16, SBO, LAYOUT_SW128);
#pragma unroll
            for (int k = 0; k < PAIR_BK / 16; ++k)
              umma_bf16_pair(acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (ks > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], 3);
            if (ks == ksteps - 1) umma_commit_pair(&tmem_full[ti], 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs): warps 2..5, TMEM lanes 32*(warp%4) .. +31 of this CTA's 128 pixels =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rx = row % P.tw, ry = (row / P.tw) % P.th, rn = row / (P.tw * P.th);
#pragma unroll 1
    for (int ti = 0; ti < p_count; ++ti) {
      int t = (p_first + ti) * 2 + (int)rank;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int x0 = tx * P.tw, y0 = ty * P.th, n0 = (t / P.tiles_y) * P.tn;
      const uint32_t acc = tmem_base + (uint32_t)ti * PAIR_BN;
      const int px = x0 + rx, py = y0 + ry, pn = n0 + rn;
      const bool valid = px < P.dw && py < P.dh && pn < P.dn;
      const long long off = (long long)pn * P.ds_n + (long long)(py * P.ostep + P.oy0) * P.ds_y +
                            (long long)(px * P.ostep + P.ox0) * P.ds_x + c0;
      mbar_wait(&tmem_full[ti], 0);
      tc_fence_after();
      epi_row<PAIR_BN / 32, false>(acc + ((uint32_t)(q * 32) << 16), sbias, P.act, P.dst, off, PAIR_BN, valid);
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();          // neither CTA may leave (or free TMEM) while the pair's MMAs / commits can still touch it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, P.tmem_cols);
  }
}

// -------------------------------------------------------------------------------------------------
// gather kernel, resident-patch variant for stride-1 k x k convolutions with few channels (opt-in: NEMAR_TC_RP3=1)
// -------------------------------------------------------------------------------------------------
// The 32/64/96-channel full-resolution layers of the STN spend their time in the per-k-step machinery of the
// kernel above: nine k-steps of two tiny MMAs each, every one behind its own TMA box and barrier round trip, the
// same pixels fetched nine times.  Here a PERSISTENT CTA keeps the layer's whole weight pack in shared memory,
// fetches ONE (TH + kh - 1) x (TW + kw - 1) pixel patch per tile and channel chunk, and feeds every tap's MMA from a
// shifted WINDOW of that patch: with TW = 8 the window of tap (dy, dx) starts (dy*PW + dx) rows into the patch and
// its 8-row groups are PW rows apart (SBO = PW rows).  The swizzle XOR is a function of the absolute shared-memory
// address, so such windows are exact for SWIZZLE_128B/64B/32B (scripts/probe/umma_shift_probe.cu).  Accumulators
// are multi-buffered in TMEM: the epilogue of tile i overlaps the MMAs of tile i+1 and the TMA of tile i+2.
constexpr int RP_TW = 8, RP_TH = 16, RP_MAX_STAGES = 8, RP_MAX_ACCS = 4;

struct Rp3Params {
  GatherParams g;
  int dy_min, dx_min;        // source offset of the patch origin relative to the tile origin
  int pw, ph;                // patch extent in pixels
  uint32_t patch_bytes;      // one channel chunk of one patch, rounded to 1 KB
  uint32_t wtile_bytes;      // one (tap, chunk) weight tile, rounded to 1 KB
  uint32_t w_tx_bytes;       // bytes the weight loads deliver
  int stages, naccs;
  int tiles_total;
  int taps_total;            // kh * kw (weight tiles are addressed by the original tap index)
};

template <int BN, int BK, bool F32OUT>
__global__ void __launch_bounds__(NTHREADS)
tc_rp3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ Rp3Params R) {
  const GatherParams& P = R.g;
  constexpr uint32_t ROWB = BK * 2;                       // bytes of one pixel's channel chunk
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int wtiles = R.taps_total * P.kchunks;            // weight tiles are indexed by the ORIGINAL tap index
  uint8_t* wsm = smem;
  const uint32_t stage_bytes = (uint32_t)P.kchunks * R.patch_bytes;
  uint8_t* psm = smem + (size_t)wtiles * R.wtile_bytes;
  uint64_t* full_bar = (uint64_t*)(psm + (size_t)R.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + RP_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + RP_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + RP_MAX_ACCS;
  uint64_t* w_bar = tmem_empty + RP_MAX_ACCS;
  uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);
  float* sbias = (float*)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);                 // [ACC_COLS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * BN;
  const int S = R.stages, NA = R.naccs;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < NA; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)NA * ACC_COLS) tmem_cols <<= 1;
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  load_bias_tile(sbias, P.bias, c0, P.cd, (int)ACC_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the weight pack once, then one patch per tile (converged warp, elected lane issues) =====
    if (elect_one()) {
      mbar_expect_tx(w_bar, R.w_tx_bytes);
      for (int t = 0; t < P.ntaps; ++t)
        for (int kc = 0; kc < P.kchunks; ++kc) {
          const int wt = P.twi[t] * P.kchunks + kc;
          tma_load_3d(wsm + (size_t)wt * R.wtile_bytes, &tmB, w_bar, 0, c0, wt);
        }
    }
    __syncwarp();
    int i = 0;
    for (int tile = blockIdx.x; tile < R.tiles_total; tile += gridDim.x, ++i) {
      const int stage = i % S;
      const uint32_t phase = (uint32_t)(i / S) & 1u;
      int t = tile;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int x0 = tx * RP_TW, y0 = ty * RP_TH, n0 = t / P.tiles_y;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], (uint32_t)P.kchunks * (uint32_t)(R.pw * R.ph) * ROWB);
        for (int kc = 0; kc < P.kchunks; ++kc)
          tma_load_4d(psm + (size_t)stage * stage_bytes + (size_t)kc * R.patch_bytes, &tmA, &full_bar[stage], kc * BK,
                      x0 + R.dx_min, y0 + R.dy_min, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    const uint32_t sbo_a = (uint32_t)R.pw * ROWB;       // 8-row groups = consecutive tile rows, PW pixels apart
    mbar_wait(w_bar, 0);
    int i = 0;
    for (int tile = blockIdx.x; tile < R.tiles_total; tile += gridDim.x, ++i) {
      const int stage = i % S, a = i % NA;
      const uint32_t sphase = (uint32_t)(i / S) & 1u, aphase = (uint32_t)(i / NA) & 1u;
      const uint32_t acc = tmem_base + (uint32_t)a * ACC_COLS;
      mbar_wait(&tmem_empty[a], aphase ^ 1);            // the epilogue has drained this accumulator
      mbar_wait(&full_bar[stage], sphase);
      tc_fence_after();
      if (elect_one()) {
        uint32_t first = 0;
        for (int kc = 0; kc < P.kchunks; ++kc) {
          const uint32_t pa = smem_u32(psm + (size_t)stage * stage_bytes + (size_t)kc * R.patch_bytes);
          for (int t = 0; t < P.ntaps; ++t) {
            const uint32_t win = (uint32_t)((P.tdy[t] - R.dy_min) * R.pw + (P.tdx[t] - R.dx_min)) * ROWB;
            const uint64_t adesc = make_smem_desc(pa + win, 16, sbo_a, layout_for(BK));
            const uint64_t bdesc = make_smem_desc(smem_u32(wsm + (size_t)(P.twi[t] * P.kchunks + kc) * R.wtile_bytes), 16,
                                                  8 * ROWB, layout_for(BK));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_bf16(acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, first);
              first = 1u;
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[a]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lanes 32*(warp%4) .. +31; lane m of the tile = pixel (m % 8, m / 8) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rx = row % RP_TW, ry = row / RP_TW;
    int cols = P.cd - c0;
    if (cols > BN) cols = BN;
    int i = 0;
    for (int tile = blockIdx.x; tile < R.tiles_total; tile += gridDim.x, ++i) {
      const int a = i % NA;
      const uint32_t aphase = (uint32_t)(i / NA) & 1u;
      int t = tile;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int px = tx * RP_TW + rx, py = ty * RP_TH + ry, pn = t / P.tiles_y;
      const bool valid = px < P.dw && py < P.dh && pn < P.dn;
      const long long off = (long long)pn * P.ds_n + (long long)py * P.ds_y + (long long)px * P.ds_x + c0;
      const uint32_t acc = tmem_base + (uint32_t)a * ACC_COLS;
      mbar_wait(&tmem_full[a], aphase);
      tc_fence_after();
      // (the accumulator is handed back only after the row has been stored: with <= 64 columns the row is one or
      //  two chunk loads, and there are RP_MAX_ACCS accumulators in flight)
      epi_row<(int)ACC_COLS / 32, F32OUT>(acc + ((uint32_t)(q * 32) << 16), sbias, P.act, P.dst, off, cols, valid);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[a]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// Pixel box (tw x th x tn, all powers of two, product `total`) that wastes the fewest accumulator rows on the
// (dw x dh x dn) index space: e.g. the 66x66 reflect-padded maps of the ResnetBlock dgrad fill only 52 % of 128x1
// boxes but 94 % of 4x4x8 ones.  Narrow boxes cost a little TMA efficiency, hence the small penalty below 8 pixels.
static void pick_tile(int dw, int dh, int dn, int& tw, int& th, int& tn, int total, bool one_sample = false) {
  double best = 1e300;
  tw = total; th = 1; tn = 1;
  for (int a = total; a >= 1; a >>= 1)
    for (int b = total / a; b >= 1; b >>= 1) {
      const int c = total / (a * b);
      if (one_sample && c != 1) continue;
      const double padded = (double)((dw + a - 1) / a * a) * ((dh + b - 1) / b * b) * ((dn + c - 1) / c * c);
      double pen = 1.0;
      for (int q = a; q < 8; q <<= 1) pen += 0.02;
      const double cost = padded * pen;
      if (cost < best * (1.0 - 1e-9)) { best = cost; tw = a; th = b; tn = c; }   // ties keep the widest box
    }
}

// tile0 / tile_end: the range of destination tiles this launch covers (tile_end < 0: all of them); force_tpc > 0
// overrides the tiles-per-CTA heuristic
template <int BN, int BK, bool F32OUT>
static int launch_gather_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const GatherParams& P, int ctiles, cudaStream_t s,
                           int tile0 = 0, int tile_end = -1, int force_tpc = 0) {
  using Cfg = GatherCfg<BN, BK>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gather_kernel<BN, BK, F32OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    NEMAR_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  // several destination tiles per CTA when the K loop is short (the fixed cost per CTA — launch, barrier and TMEM
  // set-up, pipeline fill, epilogue — then dominates), as long as the grid still spans >= 4 waves of resident CTAs
  GatherParams Q = P;
  const int tiles_all = P.tiles_x * P.tiles_y * P.tiles_n, ksteps = P.ntaps * P.kchunks;
  Q.tile0 = tile0;
  Q.tile_end = tile_end < 0 ? tiles_all : tile_end;
  const int tiles = Q.tile_end - Q.tile0;
  if (tiles <= 0) return 0;
  const int stages = Cfg::STAGES;       // ~96 KB ring (two CTAs per SM) below 256-wide tiles, ~192 KB at 256
  Q.stages = stages;
  const size_t smem_bytes = Cfg::SMEM - (size_t)(Cfg::STAGES - stages) * Cfg::STAGE_BYTES;
  int occ = (int)(233472 / (smem_bytes + 1024));
  if (occ < 1) occ = 1;
  if (occ > 8) occ = 8;
  int tpc_limit = (int)((512u / (uint32_t)occ) / Cfg::ACC_COLS);     // TMEM columns shared by the resident CTAs
  if (tpc_limit > MAX_TPC) tpc_limit = MAX_TPC;
  if (tpc_limit < 1) tpc_limit = 1;
  static const int tpc_env = [] { const char* e = getenv("NEMAR_TC_TPC"); return e ? atoi(e) : 0; }();
  int tpc = 1;
  static const int ktarget = [] { const char* e = getenv("NEMAR_TC_TPC_K"); return e ? atoi(e) : 64; }();
  while (tpc * ksteps < ktarget && tpc * 2 <= tpc_limit) tpc *= 2;
  static const int waves = [] { const char* e = getenv("NEMAR_TC_TPC_WAVES"); return e ? atoi(e) : 1; }();
  while (tpc > 1 && (long long)((tiles + tpc - 1) / tpc) * ctiles < (long long)waves * occ * sm_count()) tpc >>= 1;
  if (tpc_env > 0) tpc = tpc_env < tpc_limit ? tpc_env : tpc_limit;
  if (force_tpc > 0) tpc = force_tpc < tpc_limit ? force_tpc : tpc_limit;
  if (tpc < 1) tpc = 1;
  Q.tpc = tpc;
  uint32_t cols = 32;
  while (cols < (uint32_t)tpc * Cfg::ACC_COLS) cols <<= 1;
  Q.tmem_cols = cols;
  dim3 grid((unsigned)((tiles + tpc - 1) / tpc), (unsigned)ctiles);
  tc_gather_kernel<BN, BK, F32OUT><<<grid, NTHREADS, smem_bytes, s>>>(tmA, tmB, Q);
  nemar_note_conv_kernel("tc_gather_kernel<%d,%d,%s>", BN, BK, F32OUT ? "f32" : "bf16");
  NEMAR_LAUNCH_CHECK();
  return 0;
}

template <int BN, int BK>
static int launch_gather_f(const CUtensorMap& a, const CUtensorMap& b, const GatherParams& P, int ct, bool f32, cudaStream_t s,
                           int tile0 = 0, int tile_end = -1, int force_tpc = 0) {
  return f32 ? launch_gather_t<BN, BK, true>(a, b, P, ct, s, tile0, tile_end, force_tpc)
             : launch_gather_t<BN, BK, false>(a, b, P, ct, s, tile0, tile_end, force_tpc);
}

template <int BN>
static int launch_gather_k(const CUtensorMap& a, const CUtensorMap& b, const GatherParams& P, int ct, int bk, bool f32, cudaStream_t s,
                           int tile0 = 0, int tile_end = -1, int force_tpc = 0) {
  if (bk == 64) return launch_gather_f<BN, 64>(a, b, P, ct, f32, s, tile0, tile_end, force_tpc);
  if (bk == 32) return launch_gather_f<BN, 32>(a, b, P, ct, f32, s, tile0, tile_end, force_tpc);
  return launch_gather_f<BN, 16>(a, b, P, ct, f32, s, tile0, tile_end, force_tpc);
}

// CTA-pair launch: clusters of 2 along x; grid.x = 2 * ceil(tile pairs / tpc)
static int launch_gather_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const GatherParams& P, int ctiles, cudaStream_t s) {
  static const int stages_env = [] { const char* e = getenv("NEMAR_TC_PAIR_STAGES"); return e ? atoi(e) : 3; }();
  static const int tpc_env = [] { const char* e = getenv("NEMAR_TC_PAIR_TPC"); return e ? atoi(e) : 1; }();
  int stages = stages_env < 2 ? 2 : (stages_env > PAIR_MAX_STAGES ? PAIR_MAX_STAGES : stages_env);
  const size_t smem_bytes = pair_smem_bytes(stages);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gather_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)pair_smem_bytes(PAIR_MAX_STAGES));
    NEMAR_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(pair) failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  GatherParams Q = P;
  Q.stages = stages;
  const int occ = (int)(233472 / (smem_bytes + 1024)) < 1 ? 1 : (int)(233472 / (smem_bytes + 1024));
  int tpc = tpc_env < 1 ? 1 : (tpc_env > PAIR_MAX_TPC ? PAIR_MAX_TPC : tpc_env);
  if (occ * tpc * PAIR_BN > 512) tpc = 1;               // TMEM columns shared by the CTAs resident on one SM
  Q.tpc = tpc;
  Q.tmem_cols = (uint32_t)tpc * PAIR_BN;                // 256 or 512: powers of two
  Q.stats = nullptr;
  const int tiles = P.tiles_x * P.tiles_y * P.tiles_n, pairs = (tiles + 1) / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * ((pairs + tpc - 1) / tpc)), (unsigned)ctiles, 1);
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gather_pair_kernel, tmA, tmB, Q);
  NEMAR_REQUIRE(e == cudaSuccess, "tc_gather_pair_kernel launch failed: %s", cudaGetErrorString(e));
  nemar_note_conv_kernel("tc_gather_pair_kernel");
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// 0: single-CTA tiles everywhere; 1: CTA pairs (cta_group::2, 256x256 tiles) wherever a side has k*256 channels;
// 2: pairs in the gather kernel only; 3: pairs in the weight-gradient kernel only
static int g_pair_mode = -1;
static int pair_mode() {
  // default 3: measured on B200 (C2, batch 16) the weight-gradient pairs are 1.15x faster on the 256-channel layers
  // (3.25 -> 2.78 ms/step) while the gather pairs are within noise of the single-CTA tiles, which are not TMA-bound
  if (g_pair_mode < 0) { const char* e = getenv("NEMAR_TC_PAIR"); g_pair_mode = e ? atoi(e) : 3; }
  return g_pair_mode;
}
static bool pair_gather() { const int m = pair_mode(); return m == 1 || m == 2; }
static bool pair_wgrad() { const int m = pair_mode(); return m == 1 || m == 3; }

// resident-patch launch: persistent CTAs, one per SM (x output-channel tiles)
template <int BN, int BK, bool F32OUT>
static int launch_rp3_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const Rp3Params& R, int ctiles, size_t smem_bytes, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_rp3_kernel<BN, BK, F32OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    NEMAR_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(rp3) failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int gx = sm_count() / ctiles;
  if (gx < 1) gx = 1;
  if (gx > R.tiles_total) gx = R.tiles_total;
  dim3 grid((unsigned)gx, (unsigned)ctiles);
  tc_rp3_kernel<BN, BK, F32OUT><<<grid, NTHREADS, smem_bytes, s>>>(tmA, tmB, R);
  nemar_note_conv_kernel("tc_rp3_kernel<%d,%d,%s>", BN, BK, F32OUT ? "f32" : "bf16");
  NEMAR_LAUNCH_CHECK();
  return 0;
}
template <int BN>
static int launch_rp3_k(const CUtensorMap& a, const CUtensorMap& b, const Rp3Params& R, int ct, int bk, bool f32, size_t smem, cudaStream_t s) {
  if (bk == 64) return f32 ? launch_rp3_t<BN, 64, true>(a, b, R, ct, smem, s) : launch_rp3_t<BN, 64, false>(a, b, R, ct, smem, s);
  if (bk == 32) return f32 ? launch_rp3_t<BN, 32, true>(a, b, R, ct, smem, s) : launch_rp3_t<BN, 32, false>(a, b, R, ct, smem, s);
  return f32 ? launch_rp3_t<BN, 16, true>(a, b, R, ct, smem, s) : launch_rp3_t<BN, 16, false>(a, b, R, ct, smem, s);
}

// Plans the resident-patch kernel for one gather problem; false when it does not apply (then the tiled kernel runs).
static bool plan_rp3(const GatherParams& P, int taps_total, int BN, int BK, Rp3Params& R, size_t& smem_bytes) {
  if (BN > 64 || P.ntaps < 2) return false;
  int dy0 = P.tdy[0], dy1 = P.tdy[0], dx0 = P.tdx[0], dx1 = P.tdx[0];
  for (int t = 1; t < P.ntaps; ++t) {
    dy0 = P.tdy[t] < dy0 ? P.tdy[t] : dy0; dy1 = P.tdy[t] > dy1 ? P.tdy[t] : dy1;
    dx0 = P.tdx[t] < dx0 ? P.tdx[t] : dx0; dx1 = P.tdx[t] > dx1 ? P.tdx[t] : dx1;
  }
  if (dy1 - dy0 > 6 || dx1 - dx0 > 6) return false;
  R.g = P;
  R.dy_min = dy0; R.dx_min = dx0;
  R.pw = RP_TW + (dx1 - dx0); R.ph = RP_TH + (dy1 - dy0);
  const uint32_t rowb = (uint32_t)BK * 2;
  R.patch_bytes = round1k((uint32_t)(R.pw * R.ph) * rowb);
  R.wtile_bytes = round1k((uint32_t)BN * rowb);
  R.w_tx_bytes = (uint32_t)(P.ntaps * P.kchunks) * (uint32_t)BN * rowb;
  R.taps_total = taps_total;
  const size_t w_total = (size_t)taps_total * P.kchunks * R.wtile_bytes;
  const size_t stage = (size_t)P.kchunks * R.patch_bytes;
  const size_t fixed = 1024 /*alignment*/ + 1024 /*barriers + bias tile*/ + w_total;
  if (fixed + 2 * stage > 232448) return false;
  static const int st_env = [] { const char* e = getenv("NEMAR_TC_RP3_STAGES"); return e ? atoi(e) : 6; }();
  int stages = (int)((232448 - fixed) / stage);
  if (stages > st_env) stages = st_env;
  if (stages > RP_MAX_STAGES) stages = RP_MAX_STAGES;
  if (stages < 2) return false;
  R.stages = stages;
  R.naccs = RP_MAX_ACCS;
  R.g.tw = RP_TW; R.g.th = RP_TH; R.g.tn = 1;
  R.g.tiles_x = (P.dw + RP_TW - 1) / RP_TW;
  R.g.tiles_y = (P.dh + RP_TH - 1) / RP_TH;
  R.g.tiles_n = P.dn;
  R.tiles_total = R.g.tiles_x * R.g.tiles_y * R.g.tiles_n;
  smem_bytes = fixed + (size_t)stages * stage;
  return true;
}

static bool tc_view_ok(const nemar_tensor* t, bool allow_f32) {
  const bool dt_ok = t->dtype == NEMAR_BF16 || (allow_f32 && t->dtype == NEMAR_F32);
  return dt_ok && t->c % 16 == 0 && t->cs % 8 == 0 && t->coff % 8 == 0 && ((((uintptr_t)t->ptr) & 15) == 0);
}

static int gather_bn(int cd, int bk, bool f32) {
  // 256-channel destination tiles (one CTA per SM, 48 KB per 128x256x64 step instead of 2 x 32 KB): measured with the
  // round-2 epilogue 84 / 76 us against 94 / 85 us (fprop / dgrad of the 256 -> 256 ResnetBlock conv); NEMAR_TC_WIDE=0
  // selects the 128-wide tiles
  static const int wide = [] { const char* e = getenv("NEMAR_TC_WIDE"); return e ? atoi(e) : 1; }();
  if (wide && cd % 256 == 0 && bk == 64 && !f32) return 256;
  return (cd % 128 == 0) ? 128 : ((cd % 64 == 0) ? 64 : ((cd % 32 == 0) ? 32 : 16));
}

}  // namespace

bool tc_engine_built() { return true; }

int tc_set_option(const char* key, int value) {
  if (key && !strcmp(key, "pair")) { const int old = pair_mode(); if (value >= 0) g_pair_mode = value > 3 ? 1 : value; return old; }
  return -1;
}

// ---- InstanceNorm statistics from per-tile partials ----------------------------------------------------------------
struct StatClasses { int nclass; int off[4]; int txy[4]; };      // per parity class: first tile, tiles per sample

// stats[n][c] = sum over the tiles of sample n (all classes) of part[tile][c]; fixed order, no atomics.  A block owns 32
// channels of one sample: its 8 warps take every 8th tile (one coalesced 256-byte row of float2 per warp and tile: the
// pass is a chain of dependent L2 round trips, so the tiles are spread over the warps) and are added up in warp order.
__global__ void __launch_bounds__(256)
stats_finalize_kernel(const float2* __restrict__ part, float2* __restrict__ stats, int c, StatClasses K) {
  __shared__ float2 sp[8][32];
  const int nn = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;
  float a = 0.f, b = 0.f;
  if (ch < c) {
    for (int k = 0; k < K.nclass; ++k) {
      const float2* p = part + ((long long)K.off[k] + (long long)nn * K.txy[k]) * c + ch;
      for (int j = w; j < K.txy[k]; j += 8) {
        const float2 v = __ldg(p + (long long)j * c);
        a += v.x; b += v.y;
      }
    }
  }
  sp[w][lane] = make_float2(a, b);
  __syncthreads();
  if (w == 0 && ch < c) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a += sp[k][lane].x; b += sp[k][lane].y; }
    stats[(long long)nn * c + ch] = make_float2(a, b);
  }
}

// geometry of the destination index space of parity class `cls` (shared by the planner below and tc_gather_gemm)
static void class_extent(const nemar_tensor& dst, const GatherGeom& gg, int cls, int& dh, int& dw) {
  const int pyc = cls >> 1, pxc = cls & 1;
  dh = (gg.sd == 2) ? (dst.h - pyc + 1) / 2 : dst.h;
  dw = (gg.sd == 2) ? (dst.w - pxc + 1) / 2 : dst.w;
}

static int fused_stats_env() {
  static const int v = [] { const char* e = getenv("NEMAR_FUSED_STATS"); return e ? atoi(e) : 1; }();
  return v;
}

// Bytes of per-tile partial statistics the tiled gather kernel needs to produce InstanceNorm statistics in its
// epilogue; 0: this geometry takes the separate reduction pass (tiles spanning several samples, fp32 output, the
// CTA-pair or resident-patch kernels).
int64_t tc_gather_stats_workspace(const nemar_tensor* src_in, const nemar_tensor* dst_in, int wp_cs, const GatherGeom& gg) {
  if (!fused_stats_env() || !tc_gather_supported(src_in, dst_in, wp_cs, gg)) return 0;
  nemar_tensor dst = *dst_in;
  dst.h += 2 * dst.pad; dst.w += 2 * dst.pad; dst.pad = 0;
  const int BK = chunk_for(src_in->c);
  if (dst.dtype == NEMAR_F32) return 0;
  if (pair_gather() && dst.c % PAIR_BN == 0 && BK == PAIR_BK) return 0;
  const int BN = gather_bn(dst.c, BK, false);
  if (BN < 32) return 0;
  const int nclass = (gg.sd == 2) ? 4 : 1;
  int64_t tiles = 0;
  for (int cls = 0; cls < nclass; ++cls) {
    int dh, dw, tw, th, tn;
    class_extent(dst, gg, cls, dh, dw);
    if (dh <= 0 || dw <= 0) continue;
    pick_tile(dw, dh, dst.n, tw, th, tn, BM);
    if (tn != 1) return 0;
    tiles += (int64_t)((dw + tw - 1) / tw) * ((dh + th - 1) / th) * dst.n;
  }
  // the resident-patch kernel takes stride-1 k x k layers with <= 64 output channels: no fused statistics there
  static const int rp3_env = [] { const char* e = getenv("NEMAR_TC_RP3"); return e ? atoi(e) : 1; }();
  if (rp3_env && gg.sm == 1 && gg.sd == 1 && BN <= 64 && gg.kh * gg.kw >= 2) return 0;
  return tiles * dst.c * 2 * (int64_t)sizeof(float);
}

bool tc_gather_supported(const nemar_tensor* src, const nemar_tensor* dst, int wp_cs, const GatherGeom& gg) {
  if (!tc_view_ok(src, false) || !tc_view_ok(dst, true)) return false;
  if (wp_cs != src->c) return false;
  if (gg.kh * gg.kw > MAX_TAPS) return false;
  if (!((gg.sm == 1 || gg.sm == 2) && (gg.sd == 1 || gg.sd == 2)) || (gg.sm == 2 && gg.sd == 2)) return false;
  if (dst->pad > 0 && !gg.dst_padded) return false;   // interior-only writes into a halo'd buffer: not needed on the path
  return get_encode() != nullptr;
}

int tc_gather_gemm(const nemar_tensor* src_in, const nemar_tensor* dst_in, const void* wp, int wp_cs, const float* bias,
                   int act, float* stats, const GatherGeom& gg, cudaStream_t s, float* stats_ws, int64_t stats_ws_bytes) {
  NEMAR_REQUIRE(tc_gather_supported(src_in, dst_in, wp_cs, gg), "tc_gather_gemm: unsupported geometry");
  // read / write padded buffers as plain images of extent (h+2p, w+2p): the effective padding `pe` already
  // accounts for the halo
  nemar_tensor src = *src_in, dst = *dst_in;
  src.h += 2 * src.pad; src.w += 2 * src.pad; src.pad = 0;
  dst.h += 2 * dst.pad; dst.w += 2 * dst.pad; dst.pad = 0;
  const int BK = chunk_for(src.c);
  const bool f32 = dst.dtype == NEMAR_F32;
  const bool pair = pair_gather() && dst.c % PAIR_BN == 0 && BK == PAIR_BK && !f32;
  const int BN = pair ? PAIR_BN / 2 : gather_bn(dst.c, BK, f32);     // pair: TMA box = one CTA's half of the weight rows
  const int esz = f32 ? 4 : 2;
  const int taps_total = gg.kh * gg.kw;
  CUtensorMap tmB;
  int rc = make_w_map(&tmB, wp, dst.c, taps_total * src.c, BK, BN);
  if (rc) return rc;

  const int nclass = (gg.sd == 2) ? 4 : 1;
  if (stats) NEMAR_REQUIRE(act == NEMAR_ACT_NONE, "tc_gather_gemm: statistics need the pre-activation output");
  // statistics in the epilogue (per-tile partials in the caller's workspace) when the planner says this geometry can
  const int64_t ws_need = stats ? tc_gather_stats_workspace(src_in, dst_in, wp_cs, gg) : 0;
  const bool fused = stats && stats_ws && ws_need > 0 && stats_ws_bytes >= ws_need;
  StatClasses SC;
  SC.nclass = 0;
  int stat_tiles = 0;
  for (int cls = 0; cls < nclass; ++cls) {
    GatherParams P;
    const int pyc = cls >> 1, pxc = cls & 1;   // destination parity of this class (sd == 2)
    P.ntaps = 0;
    for (int a = 0; a < gg.kh; ++a)
      for (int b = 0; b < gg.kw; ++b) {
        int oy = -gg.pe + a, ox = -gg.pe_x + b;   // source offset relative to dst*sm
        if (gg.sd == 2) {
          int uy = pyc - gg.pe + a, ux = pxc - gg.pe_x + b;
          if ((uy & 1) || (ux & 1)) continue;
          oy = uy >> 1; ox = ux >> 1;           // arithmetic shift == floor division (numerators are even)
        }
        P.tdy[P.ntaps] = (short)oy; P.tdx[P.ntaps] = (short)ox; P.twi[P.ntaps] = (short)(a * gg.kw + b);
        ++P.ntaps;
      }
    P.ostep = (gg.sd == 2) ? 2 : 1;
    P.oy0 = (gg.sd == 2) ? pyc : 0;
    P.ox0 = (gg.sd == 2) ? pxc : 0;
    P.dh = (gg.sd == 2) ? (dst.h - pyc + 1) / 2 : dst.h;
    P.dw = (gg.sd == 2) ? (dst.w - pxc + 1) / 2 : dst.w;
    P.dn = dst.n;
    if (P.dh <= 0 || P.dw <= 0) continue;
    NEMAR_REQUIRE(P.ntaps > 0, "tc_gather_gemm: parity class without taps");
    P.kchunks = src.c / BK;
    P.cs = src.c;
    pick_tile(P.dw, P.dh, P.dn, P.tw, P.th, P.tn, BM);
    P.tiles_x = (P.dw + P.tw - 1) / P.tw;
    P.tiles_y = (P.dh + P.th - 1) / P.th;
    P.tiles_n = (P.dn + P.tn - 1) / P.tn;
    P.sm = gg.sm;
    P.ds_x = dst.cs; P.ds_y = (long long)dst.w * dst.cs; P.ds_n = (long long)dst.h * dst.w * dst.cs;
    P.dst = (char*)dst.ptr + (size_t)dst.coff * esz;
    P.cd = dst.c;
    P.bias = bias;
    P.act = act;
    P.stats = fused ? stats_ws : nullptr;
    P.stat_tile0 = stat_tiles;
    if (fused) {
      SC.off[SC.nclass] = stat_tiles;
      SC.txy[SC.nclass] = P.tiles_x * P.tiles_y;
      ++SC.nclass;
      stat_tiles += P.tiles_x * P.tiles_y * P.tiles_n;
    }
    // resident-patch kernel for stride-1 k x k layers with <= 64 output channels: default since the converged-warp issue
    // fix (32 -> 32 @256^2: 75 us against 119 us tiled; 96 -> 32: 146 / 182 against 250 / 295); NEMAR_TC_RP3=0 disables
    static const int rp3_env = [] { const char* e = getenv("NEMAR_TC_RP3"); return e ? atoi(e) : 1; }();
    if (rp3_env && !pair && !fused && gg.sm == 1 && gg.sd == 1) {
      Rp3Params R;
      size_t rp_smem = 0;
      if (plan_rp3(P, taps_total, BN, BK, R, rp_smem)) {
        CUtensorMap tmP;
        rc = make_act_map(&tmP, &src, BK, R.pw, R.ph, 1, 1);
        if (rc) return rc;
        const int ct = (dst.c + BN - 1) / BN;
        switch (BN) {
          case 64: rc = launch_rp3_k<64>(tmP, tmB, R, ct, BK, f32, rp_smem, s); break;
          case 32: rc = launch_rp3_k<32>(tmP, tmB, R, ct, BK, f32, rp_smem, s); break;
          default: rc = launch_rp3_k<16>(tmP, tmB, R, ct, BK, f32, rp_smem, s); break;
        }
        if (rc) return rc;
        continue;
      }
    }
    CUtensorMap tmA;
    rc = make_act_map(&tmA, &src, BK, P.tw, P.th, P.tn, gg.sm);
    if (rc) return rc;
    const int ctiles = pair ? dst.c / PAIR_BN : (dst.c + BN - 1) / BN;
    if (pair) {
      rc = launch_gather_pair(tmA, tmB, P, ctiles, s);
      if (rc) return rc;
      continue;
    }
    switch (BN) {
      case 256: rc = launch_gather_t<256, 64, false>(tmA, tmB, P, ctiles, s); break;
      case 128: rc = launch_gather_k<128>(tmA, tmB, P, ctiles, BK, f32, s); break;
      case 64: rc = launch_gather_k<64>(tmA, tmB, P, ctiles, BK, f32, s); break;
      case 32: rc = launch_gather_k<32>(tmA, tmB, P, ctiles, BK, f32, s); break;
      default: rc = launch_gather_k<16>(tmA, tmB, P, ctiles, BK, f32, s); break;
    }
    if (rc) return rc;
  }
  if (fused) {
    stats_finalize_kernel<<<dim3((unsigned)((dst.c + 31) / 32), (unsigned)dst.n), 256, 0, s>>>(
        reinterpret_cast<const float2*>(stats_ws), reinterpret_cast<float2*>(stats), dst.c, SC);
    NEMAR_LAUNCH_CHECK();
    return 0;
  }
  if (stats) {
    // InstanceNorm statistics of the (L2-resident) output: separate reduction pass
    return nemar_instnorm_stats(dst_in, stats, (void*)s);        // accumulates: the caller hands in zeroed stats
  }
  return 0;
}

// =================================================================================================
// weight gradient
// =================================================================================================
namespace {

constexpr int WG_KP = 64;     // pixels per k-step (box of tw x th x tn pixels)

struct WgradParams {
  int tw, th, tn;                 // pixel box (tw*th*tn == 64)
  int tiles_x, tiles_y, tiles_n;  // pixel tiles over dy
  int tiles_per_split;            // pixel tiles handled by one CTA
  int stride, kw, pe;
  int co_tiles, ci_tiles, taps;
  int ci;                         // channels of the N operand (row length of the partial buffer)
  int m_channels;                 // channels of the M operand (chunks beyond it are neither loaded nor used)
  int shift_on_a;                 // 0: A = dY, B = X (tap-shifted);  1 (swapped roles): A = X (tap-shifted), B = dY
  float* partial;                 // [split][tap][co_tiles*128][ci]
  int stages;                     // ring depth (<= WG_MAX_STAGES)
  uint32_t a_bytes, stage_bytes;  // bytes of the M-operand chunks actually loaded per stage / of one whole stage
  uint32_t bar_offset;            // barriers live behind the ring plus the slack the MMA's unused M rows read into
};

constexpr int WG_MAX_STAGES = 8;

// CA / CB: channels per TMA chunk of dY / X (64, 32 or 16 -> swizzle 128/64/32 B); BN: input channels per CTA
template <int CA, int CB, int BN>
struct WgradCfg {
  static constexpr int NA = BM / CA, NB = BN / CB;
  static constexpr uint32_t CHUNK_A = WG_KP * CA * 2, CHUNK_B = WG_KP * CB * 2;     // >= 2 KB, multiples of 1 KB
  static constexpr uint32_t B_BYTES = NB * CHUNK_B;
  static constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
};

template <int CA, int CB, int BN>
__global__ void __launch_bounds__(NTHREADS)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ WgradParams P) {
  using Cfg = WgradCfg<CA, CB, BN>;
  const int STAGES = P.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + P.bar_offset);
  uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + WG_MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x;
  const int cit = b % P.ci_tiles; b /= P.ci_tiles;
  const int cot = b % P.co_tiles;
  const int tap = b / P.co_tiles;
  const int split = blockIdx.y;
  const int ta = tap / P.kw, tb = tap % P.kw;
  const int total_tiles = P.tiles_x * P.tiles_y * P.tiles_n;
  const int t_lo = split * P.tiles_per_split;
  int t_hi = t_lo + P.tiles_per_split;
  if (t_hi > total_tiles) t_hi = total_tiles;
  const int ksteps = t_hi - t_lo;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (converged warp, elected lane issues: see tc_gather_kernel) =====
    int stage = 0; uint32_t phase = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      int t = t_lo + ks;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int tn = t / P.tiles_y;
      const int x0 = tx * P.tw, y0 = ty * P.th, n0 = tn * P.tn;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + stage * P.stage_bytes;
        // channel chunks of the M operand beyond its extent feed accumulator rows nobody reads: their loads are
        // skipped, and when the whole M operand is narrower than 128 channels the stage does not even reserve room
        // for them (the MMA then reads those rows from whatever follows — the N operand, the next stage, the slack)
        int na = (P.m_channels - cot * BM + CA - 1) / CA;
        if (na > Cfg::NA) na = Cfg::NA;
        mbar_expect_tx(&full_bar[stage], (uint32_t)na * Cfg::CHUNK_A + Cfg::B_BYTES);
        const int xs = x0 * P.stride - P.pe + tb, ys = y0 * P.stride - P.pe + ta;   // tap-shifted box of X
        const int xa = P.shift_on_a ? xs : x0, ya = P.shift_on_a ? ys : y0;
        const int xb = P.shift_on_a ? x0 : xs, yb = P.shift_on_a ? y0 : ys;
#pragma unroll
        for (int c = 0; c < Cfg::NA; ++c)
          if (c < na) tma_load_4d(sa + c * Cfg::CHUNK_A, &tmDY, &full_bar[stage], cot * BM + c * CA, xa, ya, n0);
#pragma unroll
        for (int c = 0; c < Cfg::NB; ++c)
          tma_load_4d(sa + P.a_bytes + c * Cfg::CHUNK_B, &tmX, &full_bar[stage], cit * BN + c * CB, xb, yb, n0);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);
    // MN-major: CA (CB) channels per swizzled row; LBO = distance between channel chunks (one TMA box),
    // SBO = distance between 8-pixel K groups; one UMMA consumes 16 pixels
    constexpr uint32_t SBO_A = 8 * CA * 2, SBO_B = 8 * CB * 2, KADV_A = (16 * CA * 2) >> 4, KADV_B = (16 * CB * 2) >> 4;
    int stage = 0; uint32_t phase = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * P.stage_bytes);
        const uint64_t adesc = make_smem_desc(sa, Cfg::CHUNK_A, SBO_A, layout_for(CA));
        const uint64_t bdesc = make_smem_desc(sa + P.a_bytes, Cfg::CHUNK_B, SBO_B, layout_for(CB));
#pragma unroll
        for (int k = 0; k < WG_KP / 16; ++k)
          umma_bf16(tmem_base, adesc + (uint64_t)(k * KADV_A), bdesc + (uint64_t)(k * KADV_B), idesc,
                    (ks > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (ks == ksteps - 1) umma_commit(tmem_full);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;            // output channel within the 128-tile
    float* out = P.partial + ((((long long)split * P.taps + tap) * (P.co_tiles * BM) + cot * BM + row) * (long long)P.ci) + cit * BN;
    if (ksteps > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int cc = 0; cc < (int)Cfg::TMEM_COLS; cc += 32) {
      float v[32];
      if (ksteps > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (cot * BM + row < P.m_channels) {     // rows beyond the M operand are never read by the finalize pass
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (cc + j < BN) *reinterpret_cast<float4*>(out + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// -------------------------------------------------------------------------------------------------
// weight gradient, CTA-pair variant (cta_group::2): M = 256 output channels x N = 256 input channels per pair
// -------------------------------------------------------------------------------------------------
// The single-CTA kernel stages dY[64 px][128 co] + X[64 px][256 ci] = 48 KB per 128x256x64 step.  A pair stages, per
// CTA, its own 128 output channels of dY and its own 128 of the 256 input channels of X (32 KB) for the same
// amount of MMA work per CTA: 1.5x the flops per fetched byte.  Same barrier protocol as tc_gather_pair_kernel.
constexpr uint32_t WGP_CHUNK = WG_KP * 64 * 2;                 // one 64-channel x 64-pixel TMA box (8 KB)
constexpr uint32_t WGP_A_BYTES = 2 * WGP_CHUNK, WGP_B_BYTES = 2 * WGP_CHUNK, WGP_STAGE_BYTES = WGP_A_BYTES + WGP_B_BYTES;
constexpr int WGP_MAX_STAGES = 6;
static size_t wgrad_pair_smem_bytes(int stages) { return (size_t)stages * WGP_STAGE_BYTES + 1024 + 256; }

__global__ void __launch_bounds__(NTHREADS)
tc_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ WgradParams P) {
  const int STAGES = P.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * WGP_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WGP_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + WGP_MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // blockIdx.x = ((tap * co_pairs + co_pair) * ci_tiles + cit) * 2 + rank ; P.co_tiles counts 128-channel tiles
  int b = (int)(blockIdx.x >> 1);
  const int cit = b % P.ci_tiles; b /= P.ci_tiles;
  const int co_pairs = P.co_tiles >> 1;
  const int cop = b % co_pairs;
  const int tap = b / co_pairs;
  const int cot = cop * 2 + (int)rank;          // this CTA's 128-channel tile of dY
  const int split = blockIdx.y;
  const int ta = tap / P.kw, tb = tap % P.kw;
  const int total_tiles = P.tiles_x * P.tiles_y * P.tiles_n;
  const int t_lo = split * P.tiles_per_split;
  int t_hi = t_lo + P.tiles_per_split;
  if (t_hi > total_tiles) t_hi = total_tiles;
  const int ksteps = t_hi - t_lo;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs; converged warp, elected lane issues) =====
    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[0]), 0);
    int stage = 0; uint32_t phase = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      int t = t_lo + ks;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y;
      const int tn = t / P.tiles_y;
      const int x0 = tx * P.tw, y0 = ty * P.th, n0 = tn * P.tn;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + stage * WGP_STAGE_BYTES;
        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * WGP_STAGE_BYTES);
        const uint32_t fb = full_leader + (uint32_t)stage * 8u;
        const int xs = x0 * P.stride - P.pe + tb, ys = y0 * P.stride - P.pe + ta;   // tap-shifted box of X
#pragma unroll
        for (int c = 0; c < 2; ++c) tma_load_4d_pair(sa + c * WGP_CHUNK, &tmDY, fb, cot * BM + c * 64, x0, y0, n0);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_4d_pair(sa + WGP_A_BYTES + c * WGP_CHUNK, &tmX, fb, cit * 256 + (int)rank * 128 + c * 64, xs, ys, n0);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, 256, 1, 1);
      constexpr uint32_t SBO = 8 * 64 * 2, KADV = (16 * 64 * 2) >> 4;
      int stage = 0; uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * WGP_STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa, WGP_CHUNK, SBO, LAYOUT_SW128);
          const uint64_t bdesc = make_smem_desc(sa + WGP_A_BYTES, WGP_CHUNK, SBO, LAYOUT_SW128);
#pragma unroll
          for (int k = 0; k < WG_KP / 16; ++k)
            umma_bf16_pair(tmem_base, adesc + (uint64_t)(k * KADV), bdesc + (uint64_t)(k * KADV), idesc, (ks > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage], 3);
          if (ks == ksteps - 1) umma_commit_pair(tmem_full, 3);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;            // output channel within this CTA's 128-tile
    float* out = P.partial + ((((long long)split * P.taps + tap) * (P.co_tiles * BM) + cot * BM + row) * (long long)P.ci) + cit * 256;
    if (ksteps > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int cc = 0; cc < 256; cc += 32) {
      float v[32];
      if (ksteps > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 256);
  }
}

// dw[co][ci][tap] = sum_split partial[split][tap][co][ci]   (co < co_real, ci < ci_real)
// partial is [split][tap][m_pad][n_pad]; (m,n) = (co,ci), or (ci,co) when the operand roles were swapped.
// One block per (output channel, run of input channels): the split sums are read along the partial buffer's fastest axis,
// staged in shared memory in the gradient's own [ci][tap] order, and written (or accumulated) as one contiguous run
// (round 1 wrote tap-strided 4-byte elements).  ~256 elements per block: the pass is latency-bound (one dependent chain of
// `splits` loads per element), so it wants many small blocks — a first version with one block per output channel took
// 2.1 ms per step against 1.1 ms.
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int taps, int co_real,
                      int ci_real, int m_pad, int n_pad, int swapped, int accumulate, int ci_chunk) {
  extern __shared__ float tile[];            // [ci_chunk][taps]
  const int c_o = blockIdx.x;
  const int ci0 = blockIdx.y * ci_chunk;
  const int nci = min(ci_chunk, ci_real - ci0);
  const int items = nci * taps;
  const size_t split_stride = (size_t)taps * m_pad * n_pad;
  for (int idx = threadIdx.x; idx < items; idx += blockDim.x) {
    const int tap = idx / nci, cl = idx - tap * nci, c_i = ci0 + cl;
    const int mm = swapped ? c_i : c_o, nn = swapped ? c_o : c_i;
    const float* p = partial + ((size_t)tap * m_pad + mm) * n_pad + nn;
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += __ldg(p + sp * split_stride);
    tile[cl * taps + tap] = acc;
  }
  __syncthreads();
  float* o = dw + ((size_t)c_o * ci_real + ci0) * taps;
  for (int idx = threadIdx.x; idx < items; idx += blockDim.x) o[idx] = accumulate ? o[idx] + tile[idx] : tile[idx];
}

struct WgradPlan {
  int CA, CB, BN, tw, th, tn, tiles_x, tiles_y, tiles_n, splits, tiles_per_split, co_tiles, ci_tiles, taps;
  int stages, occupancy;
  uint32_t a_bytes, stage_bytes, bar_offset, smem_bytes;
  int swapped;      // 1: M operand = X (input channels), N operand = dY — for heads with few output channels
  int pair;         // 1: CTA-pair kernel (cta_group::2, M = 256 output channels per pair)
  int64_t ws_bytes;
};

static bool wgrad_bn(int c, int& cb, int& bn) {
  cb = chunk_for(c);
  if (cb == 64) { bn = (c % 256 == 0) ? 256 : ((c % 128 == 0) ? 128 : 64); return true; }
  if (cb == 32) { bn = c; return c == 32 || c == 96 || c == 160; }
  bn = c;
  return c == 16;
}

static bool plan_wgrad(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, WgradPlan& p) {
  // the MMA cost scales with N and the accumulator has 128 M rows regardless: put the SMALLER channel count on N
  p.swapped = (dy->c <= 32 && dy->c < x->c) ? 1 : 0;
  const nemar_tensor* mop = p.swapped ? x : dy;      // M operand
  const nemar_tensor* nop = p.swapped ? dy : x;      // N operand
  if (!wgrad_bn(nop->c, p.CB, p.BN)) return false;
  p.CA = chunk_for(mop->c);
  pick_tile(dy->w, dy->h, dy->n, p.tw, p.th, p.tn, WG_KP);
  p.tiles_x = (dy->w + p.tw - 1) / p.tw;
  p.tiles_y = (dy->h + p.th - 1) / p.th;
  p.tiles_n = (dy->n + p.tn - 1) / p.tn;
  p.taps = kh * kw;
  p.co_tiles = (mop->c + BM - 1) / BM;
  p.ci_tiles = nop->c / p.BN;
  const int total = p.tiles_x * p.tiles_y * p.tiles_n;
  const int base = p.taps * p.co_tiles * p.ci_tiles;
  // shared-memory ring: only the M chunks that exist are given room when a single tile covers the M operand
  const int na_full = BM / p.CA;
  int a_chunks = (mop->c + p.CA - 1) / p.CA;
  if (p.co_tiles > 1 || a_chunks > na_full) a_chunks = na_full;
  const uint32_t chunk_a = WG_KP * p.CA * 2, b_bytes = (uint32_t)(p.BN / p.CB) * WG_KP * p.CB * 2;
  p.a_bytes = a_chunks * chunk_a;
  p.stage_bytes = p.a_bytes + b_bytes;
  p.pair = (pair_wgrad() && !p.swapped && p.CA == 64 && p.CB == 64 && p.BN == 256 && mop->c % 256 == 0) ? 1 : 0;
  if (p.pair) {
    static const int st_env = [] { const char* e = getenv("NEMAR_WG_PAIR_STAGES"); return e ? atoi(e) : 3; }();   // 3 stages: two CTAs per SM
    p.stages = st_env < 2 ? 2 : (st_env > WGP_MAX_STAGES ? WGP_MAX_STAGES : st_env);
    p.a_bytes = WGP_A_BYTES;
    p.stage_bytes = WGP_STAGE_BYTES;
    p.bar_offset = (uint32_t)p.stages * WGP_STAGE_BYTES;
    p.smem_bytes = (uint32_t)wgrad_pair_smem_bytes(p.stages);
    int occ = (int)(233472 / (p.smem_bytes + 1024));
    // ONE CTA per SM (half the K splits, half the partial-sum traffic): the kernel alone is within 5 % either way, but
    // it runs on the weight-gradient stream beside the data-gradient chain and the step is 0.7-0.9 ms shorter this way
    // (24.9 -> 24.0-24.2 ms; a 96 KB ring also leaves the SM room for the main stream's CTAs).  NEMAR_WG_PAIR_OCC=2: two.
    static const int occ_pair = [] { const char* e = getenv("NEMAR_WG_PAIR_OCC"); return e ? atoi(e) : 1; }();
    p.occupancy = occ < 1 ? 1 : (occ > 2 ? 2 : occ);        // 256 TMEM columns per CTA
    if (occ_pair >= 1 && p.occupancy > occ_pair) p.occupancy = occ_pair;
    int splits = (sm_count() * p.occupancy) / base;
    if (splits > total / 8) splits = total / 8;
    if (splits > total) splits = total;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (total + splits - 1) / splits;
    p.splits = (total + p.tiles_per_split - 1) / p.tiles_per_split;
    p.ws_bytes = (int64_t)p.splits * p.taps * p.co_tiles * BM * nop->c * 4;
    return true;
  }
  const uint32_t slack = (na_full - a_chunks) * chunk_a;
  // residency: as many CTAs per SM (<= 4, TMEM permitting) as still leaves each a ring of >= 4 stages
  const int tmem_cols = p.BN <= 32 ? 32 : (p.BN <= 64 ? 64 : (p.BN <= 128 ? 128 : 256));
  static const int occ_env = [] { const char* e = getenv("NEMAR_WG_OCC"); return e ? atoi(e) : 0; }();
  static const int occ_max = [] { const char* e = getenv("NEMAR_WG_OCC_MAX"); return e ? atoi(e) : 4; }();
  static const int min_stages = [] { const char* e = getenv("NEMAR_WG_MIN_STAGES"); return e ? atoi(e) : 4; }();
  int occ = occ_max < 1 ? 1 : (occ_max > 8 ? 8 : occ_max);
  for (;; --occ) {
    const int64_t per = 233472 / occ - 1024 /*reserved per CTA*/ - 1024 /*alignment*/ - 256 - (int64_t)slack;
    int st = (int)(per / (int64_t)p.stage_bytes);
    if (st > WG_MAX_STAGES) st = WG_MAX_STAGES;
    p.stages = st;
    const bool fits_tmem = occ * tmem_cols <= 512;
    if (occ == 1 || (fits_tmem && st >= min_stages && (occ_env <= 0 || occ <= occ_env))) break;
  }
  if (p.stages < 2) return false;
  p.occupancy = occ;
  p.bar_offset = (uint32_t)p.stages * p.stage_bytes + slack;
  p.smem_bytes = p.bar_offset + 1024 + 256;
  // split K so that the grid is (at most) ONE full wave of resident CTAs: a partial extra wave costs a whole one
  int splits = (sm_count() * occ) / base;
  if (splits > total / 8) splits = total / 8;     // >= 8 k-steps per CTA
  if (splits > total) splits = total;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (total + splits - 1) / splits;
  p.splits = (total + p.tiles_per_split - 1) / p.tiles_per_split;
  p.ws_bytes = (int64_t)p.splits * p.taps * p.co_tiles * BM * nop->c * 4;
  return true;
}

template <int CA, int CB, int BN>
static int launch_wgrad_t(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradParams& P, const WgradPlan& pl, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<CA, CB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    NEMAR_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)(pl.taps * pl.co_tiles * pl.ci_tiles), (unsigned)pl.splits);
  tc_wgrad_kernel<CA, CB, BN><<<grid, NTHREADS, pl.smem_bytes, s>>>(tmDY, tmX, P);
  nemar_note_conv_kernel("tc_wgrad_kernel<%d,%d,%d>", CA, CB, BN);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

static int launch_wgrad_pair(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradParams& P, const WgradPlan& pl, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)wgrad_pair_smem_bytes(WGP_MAX_STAGES));
    NEMAR_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(wgrad pair) failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pl.taps * pl.co_tiles * pl.ci_tiles), (unsigned)pl.splits, 1);   // co_tiles is even: 2 CTAs per pair
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_wgrad_pair_kernel, tmDY, tmX, P);
  NEMAR_REQUIRE(e == cudaSuccess, "tc_wgrad_pair_kernel launch failed: %s", cudaGetErrorString(e));
  nemar_note_conv_kernel("tc_wgrad_pair_kernel");
  NEMAR_LAUNCH_CHECK();
  return 0;
}

template <int CA>
static int launch_wgrad_a(const CUtensorMap& d, const CUtensorMap& x, const WgradParams& P, const WgradPlan& pl, cudaStream_t s) {
  if (pl.CB == 64) {
    if (pl.BN == 256) return launch_wgrad_t<CA, 64, 256>(d, x, P, pl, s);
    if (pl.BN == 128) return launch_wgrad_t<CA, 64, 128>(d, x, P, pl, s);
    return launch_wgrad_t<CA, 64, 64>(d, x, P, pl, s);
  }
  if (pl.CB == 32) {
    if (pl.BN == 160) return launch_wgrad_t<CA, 32, 160>(d, x, P, pl, s);
    if (pl.BN == 96) return launch_wgrad_t<CA, 32, 96>(d, x, P, pl, s);
    return launch_wgrad_t<CA, 32, 32>(d, x, P, pl, s);
  }
  return launch_wgrad_t<CA, 16, 16>(d, x, P, pl, s);
}

}  // namespace

bool tc_wgrad_supported(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, int stride, int pe) {
  if (!tc_view_ok(x, false) || !tc_view_ok(dy, false) || dy->pad != 0) return false;
  if (kh * kw > MAX_TAPS || !(stride == 1 || stride == 2) || pe < 0) return false;
  WgradPlan p;
  nemar_tensor xx = *x;
  xx.h += 2 * x->pad; xx.w += 2 * x->pad; xx.pad = 0;
  if (!plan_wgrad(&xx, dy, kh, kw, p)) return false;
  return get_encode() != nullptr;
}

int64_t tc_wgrad_workspace(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, int stride, int pe) {
  if (!tc_wgrad_supported(x, dy, kh, kw, stride, pe)) return 0;
  WgradPlan p;
  plan_wgrad(x, dy, kh, kw, p);
  return p.ws_bytes;
}

int tc_wgrad(const nemar_tensor* x_in, const nemar_tensor* dy, float* dw, int co_real, int ci_real, int kh, int kw,
             int stride, int pe, void* workspace, int64_t workspace_bytes, int accumulate, cudaStream_t s) {
  NEMAR_REQUIRE(tc_wgrad_supported(x_in, dy, kh, kw, stride, pe), "tc_wgrad: unsupported geometry");
  nemar_tensor x = *x_in;
  x.h += 2 * x.pad; x.w += 2 * x.pad; x.pad = 0;      // halo = real data; `pe` is relative to the padded buffer
  WgradPlan pl;
  plan_wgrad(&x, dy, kh, kw, pl);
  NEMAR_REQUIRE(workspace && workspace_bytes >= pl.ws_bytes, "tc_wgrad: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)pl.ws_bytes);
  CUtensorMap tmDY, tmX;   // named after the default roles: first = M operand (A), second = N operand (B)
  int rc = pl.swapped ? make_act_map(&tmDY, &x, pl.CA, pl.tw, pl.th, pl.tn, stride)
                      : make_act_map(&tmDY, dy, pl.CA, pl.tw, pl.th, pl.tn, 1);
  if (rc) return rc;
  rc = pl.swapped ? make_act_map(&tmX, dy, pl.CB, pl.tw, pl.th, pl.tn, 1)
                  : make_act_map(&tmX, &x, pl.CB, pl.tw, pl.th, pl.tn, stride);
  if (rc) return rc;
  WgradParams P;
  P.tw = pl.tw; P.th = pl.th; P.tn = pl.tn;
  P.tiles_x = pl.tiles_x; P.tiles_y = pl.tiles_y; P.tiles_n = pl.tiles_n;
  P.tiles_per_split = pl.tiles_per_split;
  P.stride = stride; P.kw = kw; P.pe = pe;
  P.co_tiles = pl.co_tiles; P.ci_tiles = pl.ci_tiles; P.taps = pl.taps;
  P.ci = pl.swapped ? dy->c : x.c;
  P.m_channels = pl.swapped ? x.c : dy->c;
  P.shift_on_a = pl.swapped;
  P.partial = (float*)workspace;
  P.stages = pl.stages; P.a_bytes = pl.a_bytes; P.stage_bytes = pl.stage_bytes; P.bar_offset = pl.bar_offset;
  if (pl.pair) rc = launch_wgrad_pair(tmDY, tmX, P, pl, s);
  else if (pl.CA == 64) rc = launch_wgrad_a<64>(tmDY, tmX, P, pl, s);
  else if (pl.CA == 32) rc = launch_wgrad_a<32>(tmDY, tmX, P, pl, s);
  else rc = launch_wgrad_a<16>(tmDY, tmX, P, pl, s);
  if (rc) return rc;
  int ci_chunk = 256 / pl.taps;
  if (ci_chunk < 1) ci_chunk = 1;
  if (ci_chunk > ci_real) ci_chunk = ci_real;
  const size_t fin_smem = sizeof(float) * (size_t)ci_chunk * pl.taps;
  wgrad_finalize_kernel<<<dim3((unsigned)co_real, (unsigned)((ci_real + ci_chunk - 1) / ci_chunk)), 256, fin_smem, s>>>(
      (const float*)workspace, dw, pl.splits, pl.taps, co_real, ci_real, pl.co_tiles * BM, P.ci, pl.swapped, accumulate, ci_chunk);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

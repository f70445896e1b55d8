// conv_tc.cu — tcgen05 / TMA implicit-GEMM convolution engine for sm_100a (bf16 operands, fp32 accumulation
// in TMEM).  Two kernels:
//
//  gather kernel (Conv2d fprop, Conv2d dgrad, ConvTranspose2d fprop/dgrad):
//      D[pixel][c_out] = sum_{tap, c_in} SRC[pixel @ tap][c_in] * Wp[c_out][tap][c_in]
//    M = 128 destination pixels (a TW x TH x TN box of one or more images), N = BN channels, K = taps * C_in.
//    A tiles are fetched by ONE 4-D TMA box per (tap, 64-channel chunk): the box start is shifted by the tap
//    offset, out-of-bounds rows are zero-filled by TMA (zero padding and tile overhang for free), reflect halos
//    are materialised by the producer pass so they are ordinary data, stride-2 convolutions use the TMA
//    traversal stride, and stride-2 data gradients / transposed convolutions are decomposed into the four
//    output-parity classes, each a stride-1 gather over a subset of taps.  Both operands land in shared memory
//    in the canonical K-major SWIZZLE_128B layout consumed directly by tcgen05.mma (cta_group::1, M=128).
//
//  wgrad kernel:  dW[c_out][tap][c_in] = sum_pixels dY[pixel][c_out] * X[pixel @ tap][c_in]
//    M = 128 output channels, N = BN input channels, K = pixels; both operands are the same NHWC boxes, used
//    MN-major; split-K over pixel ranges with fp32 partials reduced by a finalize kernel.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-5 = epilogue
// (TMEM -> registers -> global).  smem ring of STAGES slots guarded by full/empty mbarriers; tcgen05.commit
// releases slots and publishes the accumulator.
#include "common.cuh"
#include "conv_internal.cuh"
#include "tc_common.cuh"

#include <mutex>

using namespace tc;

namespace {

constexpr int BM = 128;         // UMMA M (TMEM lanes)
constexpr int BK = 64;          // bf16 elements per k-step = one 128-byte swizzle row
constexpr int MAX_TAPS = 49;
constexpr int NTHREADS = 192;

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
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NEMAR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: %d (c=%d cs=%d wp=%d hp=%d n=%d box=%d,%d,%d,%d es=%d)",
                (int)r, t->c, t->cs, wp, hp, t->n, box_c, bx, by, bn, es);
  return 0;
}

// 2-D weight map: rows = output channels, K contiguous
static int make_w_map(CUtensorMap* m, const void* w, int rows, int k_total, int box_rows) {
  EncodeTiledFn enc = get_encode();
  NEMAR_REQUIRE(enc, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  __nv_bfloat16* dst;            // destination base (channel offset applied)
  int cd;                        // destination channels
  const float* bias;
  int act;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS)
tc_gather_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ GatherParams P) {
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int t = blockIdx.x;
  const int tx = t % P.tiles_x; t /= P.tiles_x;
  const int ty = t % P.tiles_y;
  const int tn = t / P.tiles_y;
  const int x0 = tx * P.tw, y0 = ty * P.th, n0 = tn * P.tn;
  const int c0 = blockIdx.y * BN;
  const int ksteps = P.ntaps * P.kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0; uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        const int tap = ks / P.kchunks, kc = ks - tap * P.kchunks;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        tma_load_4d(sa, &tmA, &full_bar[stage], kc * BK, x0 * P.sm + P.tdx[tap], y0 * P.sm + P.tdy[tap], n0);
        tma_load_2d(sa + A_BYTES, &tmB, &full_bar[stage], P.twi[tap] * P.cs + kc * BK, c0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa, 16, 1024, LAYOUT_SW128);
        const uint64_t bdesc = make_smem_desc(sa + A_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)   // +32 bytes per UMMA_K inside the 128-byte swizzle row
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (ks > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else {
    // ===== epilogue: warps 2..5; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rx = row % P.tw, ry = (row / P.tw) % P.th, rn = row / (P.tw * P.th);
    const int px = x0 + rx, py = y0 + ry, pn = n0 + rn;
    const bool valid = px < P.dw && py < P.dh && pn < P.dn;
    __nv_bfloat16* out = P.dst + (long long)pn * P.ds_n + (long long)(py * P.ostep + P.oy0) * P.ds_y +
                         (long long)(px * P.ostep + P.ox0) * P.ds_x + c0;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < BN; cc += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c0 + cc + g * 8 < P.cd) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float b = P.bias ? __ldg(P.bias + c0 + cc + g * 8 + j) : 0.f;
              f[j] = act_fwd(v[g * 8 + j] + b, P.act);
            }
            uint4 pk;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            *reinterpret_cast<uint4*>(out + cc + g * 8) = pk;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

static size_t gather_smem_bytes(int BN, int stages) {
  return (size_t)stages * (BM * BK * 2 + (size_t)BN * BK * 2) + 1024 + 256;
}

static void pick_tile(int dw, int dh, int dn, int& tw, int& th, int& tn) {
  tw = 1;
  while (tw < dw && tw < 128) tw <<= 1;
  th = 1;
  while (th < dh && tw * th < 128) th <<= 1;
  tn = 128 / (tw * th);
  (void)dn;
}

template <int BN>
static int launch_gather(const CUtensorMap& tmA, const CUtensorMap& tmB, const GatherParams& P, int ctiles, cudaStream_t s) {
  constexpr int STAGES = (BN == 128) ? 3 : 4;
  size_t smem = gather_smem_bytes(BN, STAGES);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gather_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NEMAR_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)(P.tiles_x * P.tiles_y * P.tiles_n), (unsigned)ctiles);
  tc_gather_kernel<BN, STAGES><<<grid, NTHREADS, smem, s>>>(tmA, tmB, P);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

static bool tc_view_ok(const nemar_tensor* t) {
  return t->dtype == NEMAR_BF16 && t->c % 64 == 0 && t->cs % 8 == 0 && t->coff % 8 == 0 && ((((uintptr_t)t->ptr) & 15) == 0);
}

}  // namespace

bool tc_engine_built() { return true; }

bool tc_gather_supported(const nemar_tensor* src, const nemar_tensor* dst, int wp_cs, const GatherGeom& gg) {
  if (!tc_view_ok(src) || !tc_view_ok(dst)) return false;
  if (wp_cs != src->c) return false;
  if (gg.kh * gg.kw > MAX_TAPS) return false;
  if (!((gg.sm == 1 || gg.sm == 2) && (gg.sd == 1 || gg.sd == 2)) || (gg.sm == 2 && gg.sd == 2)) return false;
  if (dst->pad > 0 && !gg.dst_padded) return false;   // interior-only writes into a halo'd buffer: not needed on the path
  return get_encode() != nullptr;
}

int tc_gather_gemm(const nemar_tensor* src_in, const nemar_tensor* dst_in, const void* wp, int wp_cs, const float* bias,
                   int act, float* stats, const GatherGeom& gg, cudaStream_t s) {
  NEMAR_REQUIRE(tc_gather_supported(src_in, dst_in, wp_cs, gg), "tc_gather_gemm: unsupported geometry");
  // read / write padded buffers as plain images of extent (h+2p, w+2p): the effective padding `pe` already
  // accounts for the halo
  nemar_tensor src = *src_in, dst = *dst_in;
  src.h += 2 * src.pad; src.w += 2 * src.pad; src.pad = 0;
  dst.h += 2 * dst.pad; dst.w += 2 * dst.pad; dst.pad = 0;
  const int BN = (dst.c % 128 == 0) ? 128 : 64;
  const int taps_total = gg.kh * gg.kw;
  CUtensorMap tmB;
  int rc = make_w_map(&tmB, wp, dst.c, taps_total * src.c, BN);
  if (rc) return rc;

  const int nclass = (gg.sd == 2) ? 4 : 1;
  for (int cls = 0; cls < nclass; ++cls) {
    GatherParams P;
    const int pyc = cls >> 1, pxc = cls & 1;   // destination parity of this class (sd == 2)
    P.ntaps = 0;
    for (int a = 0; a < gg.kh; ++a)
      for (int b = 0; b < gg.kw; ++b) {
        int oy = -gg.pe + a, ox = -gg.pe + b;   // source offset relative to dst*sm
        if (gg.sd == 2) {
          int uy = pyc - gg.pe + a, ux = pxc - gg.pe + b;
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
    if (P.ntaps == 0) {
      // a parity class no tap reaches (possible only for k < stride): its outputs are bias-only; not on the path
      NEMAR_REQUIRE(false, "tc_gather_gemm: parity class without taps");
    }
    P.kchunks = src.c / BK;
    P.cs = src.c;
    pick_tile(P.dw, P.dh, P.dn, P.tw, P.th, P.tn);
    P.tiles_x = (P.dw + P.tw - 1) / P.tw;
    P.tiles_y = (P.dh + P.th - 1) / P.th;
    P.tiles_n = (P.dn + P.tn - 1) / P.tn;
    P.sm = gg.sm;
    P.ds_x = dst.cs; P.ds_y = (long long)dst.w * dst.cs; P.ds_n = (long long)dst.h * dst.w * dst.cs;
    P.dst = (__nv_bfloat16*)dst.ptr + dst.coff;
    P.cd = dst.c;
    P.bias = bias;
    P.act = act;
    CUtensorMap tmA;
    rc = make_act_map(&tmA, &src, BK, P.tw, P.th, P.tn, gg.sm);
    if (rc) return rc;
    const int ctiles = (dst.c + BN - 1) / BN;
    rc = (BN == 128) ? launch_gather<128>(tmA, tmB, P, ctiles, s) : launch_gather<64>(tmA, tmB, P, ctiles, s);
    if (rc) return rc;
  }
  if (stats) {
    NEMAR_REQUIRE(act == NEMAR_ACT_NONE, "tc_gather_gemm: statistics need the pre-activation output");
    return nemar_instnorm_stats(dst_in, stats, (void*)s);
  }
  return 0;
}

// =================================================================================================
// weight gradient
// =================================================================================================
namespace {

constexpr int WG_KP = 64;     // pixels per k-step (box of kw x kh x kn pixels)

struct WgradParams {
  int tw, th, tn;                 // pixel box (tw*th*tn == 64)
  int tiles_x, tiles_y, tiles_n;  // pixel tiles over dy
  int tiles_per_split;            // pixel tiles handled by one CTA
  int stride;                     // conv stride (x box start = dy coord * stride + tap offset)
  int tap_dy, tap_dx;             // filled per launch? no: derived from blockIdx (see kernel)
  int kw, pe;
  int co_tiles, ci_tiles;
  int co, ci, taps;
  float* partial;                 // [split][tap][co_pad][ci]  (co_pad = co_tiles*128)
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ WgradParams P) {
  // A = dY box: two 64-channel chunks (M = 128 output channels), MN-major; B = X box: BN/64 chunks, MN-major
  constexpr uint32_t CHUNK_BYTES = WG_KP * 128;                 // 64 pixels x 64 channels x 2 B
  constexpr uint32_t A_BYTES = 2 * CHUNK_BYTES, B_BYTES = (BN / 64) * CHUNK_BYTES, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
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
  if (warp == 2) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        int t = t_lo + ks;
        const int tx = t % P.tiles_x; t /= P.tiles_x;
        const int ty = t % P.tiles_y;
        const int tn = t / P.tiles_y;
        const int x0 = tx * P.tw, y0 = ty * P.th, n0 = tn * P.tn;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_4d(sa + c * CHUNK_BYTES, &tmDY, &full_bar[stage], cot * 128 + c * 64, x0, y0, n0);
#pragma unroll
        for (int c = 0; c < BN / 64; ++c)
          tma_load_4d(sa + A_BYTES + c * CHUNK_BYTES, &tmX, &full_bar[stage], cit * BN + c * 64,
                      x0 * P.stride - P.pe + tb, y0 * P.stride - P.pe + ta, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        // MN-major SW128: 64 MN elements per 128-byte row; LBO = distance between 64-element MN chunks
        // (= one TMA box), SBO = distance between 8-row K groups (1024 B); UMMA_K = 16 rows = 2048 B
        const uint64_t adesc = make_smem_desc(sa, CHUNK_BYTES, 1024, LAYOUT_SW128);
        const uint64_t bdesc = make_smem_desc(sa + A_BYTES, CHUNK_BYTES, 1024, LAYOUT_SW128);
#pragma unroll
        for (int k = 0; k < WG_KP / 16; ++k)
          umma_bf16(tmem_base, adesc + (uint64_t)(k * (2048 >> 4)), bdesc + (uint64_t)(k * (2048 >> 4)), idesc,
                    (ks > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;            // output channel within the 128-tile
    const int co = cot * 128 + row;
    float* out = P.partial + ((((long long)split * P.taps + tap) * (P.co_tiles * 128) + co) * (long long)P.ci) + cit * BN;
    if (ksteps > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int cc = 0; cc < BN; cc += 32) {
      float v[32];
      if (ksteps > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(out + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

// dw[co][ci][tap] = sum_split partial[split][tap][co][ci]
__global__ void wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int taps,
                                      int co, int co_pad, int ci) {
  const int64_t total = (int64_t)co * ci * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // iterate with ci fastest for coalesced partial reads
    int c_i = (int)(i % ci);
    int64_t r = i / ci;
    int tap = (int)(r % taps);
    int c_o = (int)(r / taps);
    float acc = 0.f;
    for (int s = 0; s < splits; ++s)
      acc += __ldg(partial + (((int64_t)s * taps + tap) * co_pad + c_o) * ci + c_i);
    dw[((int64_t)c_o * ci + c_i) * taps + tap] = acc;
  }
}

struct WgradPlan {
  int BN, tw, th, tn, tiles_x, tiles_y, tiles_n, splits, tiles_per_split, co_tiles, ci_tiles, taps;
  int64_t ws_bytes;
};

static void pick_pixel_tile(int dw, int dh, int& tw, int& th, int& tn) {
  tw = 1;
  while (tw < dw && tw < WG_KP) tw <<= 1;
  th = 1;
  while (th < dh && tw * th < WG_KP) th <<= 1;
  tn = WG_KP / (tw * th);
}

static WgradPlan plan_wgrad(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw) {
  WgradPlan p;
  p.BN = (x->c % 256 == 0) ? 256 : ((x->c % 128 == 0) ? 128 : 64);
  pick_pixel_tile(dy->w, dy->h, p.tw, p.th, p.tn);
  p.tiles_x = (dy->w + p.tw - 1) / p.tw;
  p.tiles_y = (dy->h + p.th - 1) / p.th;
  p.tiles_n = (dy->n + p.tn - 1) / p.tn;
  p.taps = kh * kw;
  p.co_tiles = (dy->c + 127) / 128;
  p.ci_tiles = x->c / p.BN;
  const int total = p.tiles_x * p.tiles_y * p.tiles_n;
  const int base = p.taps * p.co_tiles * p.ci_tiles;
  int splits = (148 * 2 + base - 1) / base;        // ~2 CTAs per SM in flight
  if (splits > total) splits = total;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (total + splits - 1) / splits;
  p.splits = (total + p.tiles_per_split - 1) / p.tiles_per_split;
  p.ws_bytes = (int64_t)p.splits * p.taps * p.co_tiles * 128 * x->c * 4;
  return p;
}

template <int BN>
static int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradParams& P, const WgradPlan& pl, cudaStream_t s) {
  constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 4 : 6);
  constexpr uint32_t CHUNK = WG_KP * 128;
  size_t smem = (size_t)STAGES * (2 * CHUNK + (BN / 64) * CHUNK) + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NEMAR_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)(pl.taps * pl.co_tiles * pl.ci_tiles), (unsigned)pl.splits);
  tc_wgrad_kernel<BN, STAGES><<<grid, NTHREADS, smem, s>>>(tmDY, tmX, P);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool tc_wgrad_supported(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, int stride, int pe) {
  if (!tc_view_ok(x) || !tc_view_ok(dy) || dy->pad != 0) return false;
  if (kh * kw > MAX_TAPS || !(stride == 1 || stride == 2) || pe < 0) return false;
  return get_encode() != nullptr;
}

int64_t tc_wgrad_workspace(const nemar_tensor* x, const nemar_tensor* dy, int kh, int kw, int stride, int pe) {
  if (!tc_wgrad_supported(x, dy, kh, kw, stride, pe)) return 0;
  return plan_wgrad(x, dy, kh, kw).ws_bytes;
}

int tc_wgrad(const nemar_tensor* x_in, const nemar_tensor* dy, float* dw, int kh, int kw, int stride, int pe,
             void* workspace, int64_t workspace_bytes, cudaStream_t s) {
  NEMAR_REQUIRE(tc_wgrad_supported(x_in, dy, kh, kw, stride, pe), "tc_wgrad: unsupported geometry");
  nemar_tensor x = *x_in;
  x.h += 2 * x.pad; x.w += 2 * x.pad; x.pad = 0;      // halo = real data; `pe` is relative to the padded buffer
  WgradPlan pl = plan_wgrad(&x, dy, kh, kw);
  NEMAR_REQUIRE(workspace && workspace_bytes >= pl.ws_bytes, "tc_wgrad: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)pl.ws_bytes);
  CUtensorMap tmDY, tmX;
  int rc = make_act_map(&tmDY, dy, 64, pl.tw, pl.th, pl.tn, 1);
  if (rc) return rc;
  rc = make_act_map(&tmX, &x, 64, pl.tw, pl.th, pl.tn, stride);
  if (rc) return rc;
  WgradParams P;
  P.tw = pl.tw; P.th = pl.th; P.tn = pl.tn;
  P.tiles_x = pl.tiles_x; P.tiles_y = pl.tiles_y; P.tiles_n = pl.tiles_n;
  P.tiles_per_split = pl.tiles_per_split;
  P.stride = stride; P.kw = kw; P.pe = pe; P.tap_dy = 0; P.tap_dx = 0;
  P.co_tiles = pl.co_tiles; P.ci_tiles = pl.ci_tiles;
  P.co = dy->c; P.ci = x.c; P.taps = pl.taps;
  P.partial = (float*)workspace;
  if (pl.BN == 256) rc = launch_wgrad<256>(tmDY, tmX, P, pl, s);
  else if (pl.BN == 128) rc = launch_wgrad<128>(tmDY, tmX, P, pl, s);
  else rc = launch_wgrad<64>(tmDY, tmX, P, pl, s);
  if (rc) return rc;
  const int64_t total = (int64_t)dy->c * x.c * pl.taps;
  wgrad_finalize_kernel<<<grid_for(total, 256), 256, 0, s>>>((const float*)workspace, dw, pl.splits, pl.taps, dy->c,
                                                              pl.co_tiles * 128, x.c);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// conv_generic.cu — CUDA-core implicit-GEMM convolution for every geometry on the path (any channel
// count, stride, padding, Conv2d or ConvTranspose2d).  It is the fp32 "parity mode" engine, the path for
// layers the tcgen05 kernel does not take (3/6-channel heads, 1/2/3-channel tails, tiny bottleneck maps),
// and the cross-check for conv_tc.cu.  fp32 accumulation; storage fp32 or bf16.
//
// A forward / data-gradient pass is one "gather GEMM":
//   dst[n, oy, ox, oc] = sum_{a,b,ic} src[n, (oy*sm - pe + a)/sd, (ox*sm - pe + b)/sd, ic] * Wp[oc][a][b][ic]
// (terms whose coordinate is not divisible by sd, or falls outside src, are zero), with
//   Conv2d fprop / ConvT dgrad : sm = stride, sd = 1      Conv2d dgrad / ConvT fprop : sm = 1, sd = stride.
#include "common.cuh"
#include "conv_internal.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
gather_gemm_kernel(TView src, TView dst, const void* __restrict__ wp, int w_dtype, int wp_cs,
                   const float* __restrict__ bias, int act, GatherGeom gg) {
  // dst index space: [n][DH][DW] (padded coords when gg.dst_padded, else interior)
  const int DH = gg.dst_padded ? dst.hp : dst.h, DW = gg.dst_padded ? dst.wp : dst.w;
  const int SH = src.hp, SW = src.wp;  // source read in padded coordinates (halo is real data)
  const int64_t M = (int64_t)dst.n * DH * DW;
  const int N = dst.c, CS = src.c;
  const int taps = gg.kh * gg.kw;
  const int kchunks = (CS + TK - 1) / TK;
  const int wbk = packed_bk(wp_cs);

  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ int pn[TM], py[TM], px[TM];

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;

  if (tid < TM) {
    int64_t m = m0 + tid;
    if (m < M) {
      int x = (int)(m % DW);
      int64_t r = m / DW;
      int y = (int)(r % DH);
      pn[tid] = (int)(r / DH);
      py[tid] = y * gg.sm - gg.pe;
      px[tid] = x * gg.sm - gg.pe_x;
    } else {
      pn[tid] = -1; py[tid] = 0; px[tid] = 0;
    }
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int tr = tid / 16, tc = tid % 16;  // micro-tile origin: rows tr*4.., cols tc*4..
  // loader mapping: A: 64 rows x 16 k -> 1024 elements, 4 per thread (row = tid/4, k = (tid%4)*4..+3)
  const int arow = tid >> 2, ak = (tid & 3) * 4;
  // B: 64 cols x 16 k (col = tid/4, k = (tid%4)*4..+3), contiguous along k in Wp
  const int bcol = tid >> 2, bk = (tid & 3) * 4;

  for (int tap = 0; tap < taps; ++tap) {
    const int a = tap / gg.kw, b = tap % gg.kw;
    // source pixel for my A row under this tap
    int sn = pn[arow];
    int uy = py[arow] + a, ux = px[arow] + b;
    bool ok = sn >= 0 && uy >= 0 && ux >= 0;
    if (gg.sd > 1) {
      ok = ok && (uy % gg.sd == 0) && (ux % gg.sd == 0);
      uy /= gg.sd; ux /= gg.sd;
    }
    ok = ok && uy < SH && ux < SW;
    const int64_t sp = ok ? src.pix_p(sn, uy, ux) : 0;
    for (int kc = 0; kc < kchunks; ++kc) {
      const int c0 = kc * TK;
      // ---- load A tile
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int ch = c0 + ak + q;
        float v = 0.f;
        if (ok && ch < CS) v = ld_rt(src.ptr, src.dtype, sp + ch);
        As[ak + q][arow] = v;
      }
      // ---- load B tile
      {
        int oc = n0 + bcol;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int ch = c0 + bk + q;
          float v = 0.f;
          if (oc < N && ch < CS)
            v = ld_rt(wp, w_dtype, wbk ? packed_index(oc, tap, ch, N, wp_cs, wbk) : packed_index_rm(oc, tap, ch, taps, wp_cs));
          Bs[bk + q][bcol] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[k][tr * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tc * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = tr * 4 + i;
    int64_t m = m0 + row;
    if (m >= M) continue;
    int x = (int)(m % DW);
    int64_t r = m / DW;
    int y = (int)(r % DH);
    int nn = (int)(r / DH);
    int64_t base = gg.dst_padded ? dst.pix_p(nn, y, x) : dst.pix(nn, y, x);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int oc = n0 + tc * 4 + j;
      if (oc >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + oc) : 0.f);
      st_rt(dst.ptr, dst.dtype, base + oc, act_fwd(v, act));
    }
  }
}

// weight gradient, conv-view:  dw[co][ci][a][b] += sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*s - pe + a, ox*s - pe + b, ci]
__global__ void __launch_bounds__(256)
wgrad_kernel(TView x, TView dy, float* __restrict__ dw, int kh, int kw, int stride, int pe, int64_t pix_per_split) {
  constexpr int TP = 32;  // pixels per smem tile
  __shared__ float Ds[TP][32 + 1];  // [pix][co]
  __shared__ float Xs[TP][32 + 1];  // [pix][ci]
  const int taps = kh * kw;
  const int ci_tiles = (x.c + 31) / 32;
  const int tap = blockIdx.x / ci_tiles, ci0 = (blockIdx.x % ci_tiles) * 32;
  const int co0 = blockIdx.y * 32;
  const int a = tap / kw, b = tap % kw;
  const int64_t P = (int64_t)dy.n * dy.h * dy.w;
  const int64_t p_lo = blockIdx.z * pix_per_split;
  int64_t p_hi = p_lo + pix_per_split;
  if (p_hi > P) p_hi = P;
  const int tid = threadIdx.x;
  const int tco = tid / 16, tci = tid % 16;  // each thread: 2 co x 2 ci
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const int lrow = tid >> 3, lcol = (tid & 7) * 4;  // loader: 32 rows x 32 cols, 4 per thread
  for (int64_t p0 = p_lo; p0 < p_hi; p0 += TP) {
    int64_t p = p0 + lrow;
    bool pv = p < p_hi;
    int ox = 0, oy = 0, nn = 0;
    if (pv) {
      ox = (int)(p % dy.w);
      int64_t r = p / dy.w;
      oy = (int)(r % dy.h);
      nn = (int)(r / dy.h);
    }
    {
      const int64_t dp = pv ? dy.pix(nn, oy, ox) : 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int co = co0 + lcol + q;
        Ds[lrow][lcol + q] = (pv && co < dy.c) ? ld_rt(dy.ptr, dy.dtype, dp + co) : 0.f;
      }
      int iy = oy * stride - pe + a, ix = ox * stride - pe + b;  // padded coords of x
      bool xv = pv && iy >= 0 && iy < x.hp && ix >= 0 && ix < x.wp;
      const int64_t xp = xv ? x.pix_p(nn, iy, ix) : 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int ci = ci0 + lcol + q;
        Xs[lrow][lcol + q] = (xv && ci < x.c) ? ld_rt(x.ptr, x.dtype, xp + ci) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < TP; ++k) {
      float d0 = Ds[k][tco * 2], d1 = Ds[k][tco * 2 + 1];
      float x0 = Xs[k][tci * 2], x1 = Xs[k][tci * 2 + 1];
      acc[0][0] = fmaf(d0, x0, acc[0][0]); acc[0][1] = fmaf(d0, x1, acc[0][1]);
      acc[1][0] = fmaf(d1, x0, acc[1][0]); acc[1][1] = fmaf(d1, x1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int co = co0 + tco * 2 + i, ci = ci0 + tci * 2 + j;
      if (co < dy.c && ci < x.c) atomicAdd(dw + (((int64_t)co * x.c + ci) * taps + tap), acc[i][j]);
    }
}

// ---- weight packing --------------------------------------------------------------------------
// out[o][a][b][i] (o padded to op, i padded to ip; padding is zero) = w[(o,i) or (i,o)][a or KH-1-a][b or KW-1-b]
template <typename T>
__global__ void pack_kernel(const float* __restrict__ w, T* __restrict__ out, int O, int op, int I, int ip, int kh,
                            int kw, int w_is_oi /* w indexed [o][i] else [i][o] */, int flip) {
  const int64_t total = (int64_t)op * kh * kw * ip;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % ip);
    int64_t r = idx / ip;
    int b = (int)(r % kw); r /= kw;
    int a = (int)(r % kh);
    int o = (int)(r / kh);
    float v = 0.f;
    if (i < I && o < O) {
      int aa = flip ? kh - 1 - a : a, bb = flip ? kw - 1 - b : b;
      int64_t widx = w_is_oi ? (((int64_t)o * I + i) * kh + aa) * kw + bb
                             : (((int64_t)i * O + o) * kh + aa) * kw + bb;
      v = __ldg(w + widx);
    }
    const int bk = packed_bk(ip);
    const int64_t oidx = bk ? packed_index(o, a * kw + b, i, op, ip, bk) : idx;
    out[oidx] = from_f<T>(v);
  }
}

// ---- multi-tensor packing: every layer of an optimizer in ONE launch --------------------------------------------
// jobs[j] describes one pack (or, with kind 2, one zero-padded fp32 bias copy); blocks[b] = (job, first element of the
// 2048-element slice this block handles).  Same element mapping as pack_kernel.
__global__ void __launch_bounds__(256) pack_multi_kernel(const nemar_pack_job* __restrict__ jobs, const int2* __restrict__ blocks) {
  const int2 bj = blocks[blockIdx.x];
  const nemar_pack_job J = jobs[bj.x];
  const int64_t total = (int64_t)J.op * J.kh * J.kw * J.ip;
  const int64_t lo = (int64_t)bj.y * 2048;
  const int bk = (J.kind == 2) ? 0 : packed_bk(J.ip);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t idx = lo + u * 256 + threadIdx.x;
    if (idx >= total) break;
    int i = (int)(idx % J.ip);
    int64_t r = idx / J.ip;
    int b = (int)(r % J.kw); r /= J.kw;
    int a = (int)(r % J.kh);
    int o = (int)(r / J.kh);
    float v = 0.f;
    if (i < J.I && o < J.O) {
      int aa = J.flip ? J.kh - 1 - a : a, bb = J.flip ? J.kw - 1 - b : b;
      int64_t widx = J.w_is_oi ? (((int64_t)o * J.I + i) * J.kh + aa) * J.kw + bb
                               : (((int64_t)i * J.O + o) * J.kh + aa) * J.kw + bb;
      v = __ldg(J.w + widx);
    }
    const int64_t oidx = bk ? packed_index(o, a * J.kw + b, i, J.op, J.ip, bk) : idx;
    if (J.dtype == NEMAR_BF16) ((__nv_bfloat16*)J.out)[oidx] = __float2bfloat16(v);
    else ((float*)J.out)[oidx] = v;
  }
}

}  // namespace

int generic_pack_multi(const nemar_pack_job* jobs_dev, const int* blocks_dev, int nblocks, cudaStream_t s) {
  pack_multi_kernel<<<nblocks, 256, 0, s>>>(jobs_dev, (const int2*)blocks_dev);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

int generic_gather_gemm(const nemar_tensor* src, const nemar_tensor* dst, const void* wp, int w_dtype, int wp_cs,
                        const float* bias, int act, const GatherGeom& gg, cudaStream_t s) {
  TView sv = make_view(src), dv = make_view(dst);
  nemar_note_conv_kernel("generic_gather_kernel");
  const int DH = gg.dst_padded ? dv.hp : dv.h, DW = gg.dst_padded ? dv.wp : dv.w;
  const int64_t M = (int64_t)dv.n * DH * DW;
  dim3 grid((unsigned)ceil_div64(M, TM), (unsigned)((dv.c + TN - 1) / TN));
  gather_gemm_kernel<<<grid, 256, 0, s>>>(sv, dv, wp, w_dtype, wp_cs, bias, act, gg);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

int generic_wgrad(const nemar_tensor* x, const nemar_tensor* dy, float* dw, int kh, int kw, int stride, int pe,
                  int accumulate, cudaStream_t s) {
  TView xv = make_view(x), dv = make_view(dy);
  nemar_note_conv_kernel("generic_wgrad_kernel");
  const int taps = kh * kw;
  if (!accumulate) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)dv.c * xv.c * taps, s);
  const int64_t P = (int64_t)dv.n * dv.h * dv.w;
  int ci_tiles = (xv.c + 31) / 32, co_tiles = (dv.c + 31) / 32;
  int64_t base_blocks = (int64_t)taps * ci_tiles * co_tiles;
  int splits = (int)((148 * 8 + base_blocks - 1) / base_blocks);
  int64_t max_splits = (P + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t pps = (P + splits - 1) / splits;
  pps = (pps + 31) / 32 * 32;
  splits = (int)((P + pps - 1) / pps);
  dim3 grid(taps * ci_tiles, co_tiles, splits);
  wgrad_kernel<<<grid, 256, 0, s>>>(xv, dv, dw, kh, kw, stride, pe, pps);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

int generic_pack(const float* w, void* out, int dtype, int O, int op, int I, int ip, int kh, int kw, int w_is_oi,
                 int flip, cudaStream_t s) {
  int64_t total = (int64_t)op * kh * kw * ip;
  DISPATCH_DTYPE(dtype, T, (pack_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(w, (T*)out, O, op, I, ip, kh, kw,
                                                                                 w_is_oi, flip)));
  NEMAR_LAUNCH_CHECK();
  return 0;
}

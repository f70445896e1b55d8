// misc.cu — scalar losses, the affine STN's Linear head, flat Adam, error plumbing.
// Reference call sites: models/nemar_model.py:68,128-137,179,195; models/networks.py:237-238,273-275;
// models/stn/affine_stn.py:69-72,136-138.
#include "common.cuh"
#include <mutex>
#include <string>

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void nemar_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
NEMAR_API const char* nemar_last_error(void) { return g_err; }
// name of the kernel the most recent convolution call on this thread launched (measurement aid: bench.py keys its
// per-kernel timings on it, so that "the dominant kernel" is a kernel instance and not an op class)
static thread_local char g_conv_kernel[96] = "";
void nemar_note_conv_kernel(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_conv_kernel, sizeof(g_conv_kernel), fmt, ap);
  va_end(ap);
}
NEMAR_API const char* nemar_last_conv_kernel(void) { return g_conv_kernel; }
NEMAR_API int nemar_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
// L1 / mean|x| / MSE-vs-constant.  out[0] += scale * mean(...)
// ---------------------------------------------------------------------------------------------
template <int MODE>  // 0: |a-b|   1: |a|
__global__ void __launch_bounds__(256)
abs_mean_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t numel, float k,
                    float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = __ldg(a + i);
    if (MODE == 0) v -= __ldg(b + i);
    acc += fabsf(v);
  }
  float s = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, s * k);
}

template <int MODE>
__global__ void abs_mean_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t numel,
                                    float k, const float* __restrict__ gscale, float* __restrict__ da,
                                    int accumulate) {
  const float g = __ldg(gscale) * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = __ldg(a + i);
    if (MODE == 0) v -= __ldg(b + i);
    float d = g * (float)((v > 0.f) - (v < 0.f));
    if (accumulate) da[i] += d;
    else da[i] = d;
  }
}

NEMAR_API int nemar_l1_fwd(const float* a, const float* b, int64_t numel, float scale, float* out,
                           void* stream) {
  NEMAR_REQUIRE(a && b && out && numel > 0, "l1_fwd: bad args");
  abs_mean_fwd_kernel<0><<<grid_for(numel, 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(
      a, b, numel, scale / (float)numel, out);
  NEMAR_LAUNCH_CHECK();
  return 0;
}
NEMAR_API int nemar_l1_bwd(const float* a, const float* b, int64_t numel, float scale, const float* gscale,
                           float* da, int accumulate, void* stream) {
  NEMAR_REQUIRE(a && b && gscale && da && numel > 0, "l1_bwd: bad args");
  abs_mean_bwd_kernel<0><<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(
      a, b, numel, scale / (float)numel, gscale, da, accumulate);
  NEMAR_LAUNCH_CHECK();
  return 0;
}
NEMAR_API int nemar_mean_abs_fwd(const float* a, int64_t numel, float scale, float* out, void* stream) {
  NEMAR_REQUIRE(a && out && numel > 0, "mean_abs_fwd: bad args");
  abs_mean_fwd_kernel<1><<<grid_for(numel, 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(
      a, nullptr, numel, scale / (float)numel, out);
  NEMAR_LAUNCH_CHECK();
  return 0;
}
NEMAR_API int nemar_mean_abs_bwd(const float* a, int64_t numel, float scale, const float* gscale, float* da,
                                 void* stream) {
  NEMAR_REQUIRE(a && gscale && da && numel > 0, "mean_abs_bwd: bad args");
  abs_mean_bwd_kernel<1><<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(
      a, nullptr, numel, scale / (float)numel, gscale, da, 0);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256)
mse_const_fwd_kernel(TView p, float target, float k, float* __restrict__ out) {
  __shared__ float red[32];
  const int64_t total = (int64_t)p.n * p.h * p.w * p.c;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % p.c);
    int64_t r = i / p.c;
    int x = (int)(r % p.w); r /= p.w;
    int y = (int)(r % p.h);
    int nn = (int)(r / p.h);
    float d = to_f<T>(((const T*)p.ptr)[p.pix(nn, y, x) + ch]) - target;
    acc += d * d;
  }
  float s = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, s * k);
}

template <typename T>
__global__ void mse_const_bwd_kernel(TView p, float target, float k, const float* __restrict__ gscale,
                                     TView dp) {
  const int64_t total = (int64_t)p.n * p.h * p.w * p.c;
  const float g = __ldg(gscale) * k * 2.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % p.c);
    int64_t r = i / p.c;
    int x = (int)(r % p.w); r /= p.w;
    int y = (int)(r % p.h);
    int nn = (int)(r / p.h);
    float d = to_f<T>(((const T*)p.ptr)[p.pix(nn, y, x) + ch]) - target;
    ((T*)dp.ptr)[dp.pix(nn, y, x) + ch] = from_f<T>(g * d);
  }
}

NEMAR_API int nemar_mse_const_fwd(const nemar_tensor* pred, float target, float scale, float* out,
                                  void* stream) {
  NEMAR_REQUIRE(view_ok(pred) && out, "mse_const_fwd: bad args");
  TView p = make_view(pred);
  int64_t total = (int64_t)p.n * p.h * p.w * p.c;
  DISPATCH_DTYPE(p.dtype, T, (mse_const_fwd_kernel<T><<<grid_for(total, 256, 148 * 2), 256, 0,
                                                        (cudaStream_t)stream>>>(p, target, scale / (float)total,
                                                                                out)));
  NEMAR_LAUNCH_CHECK();
  return 0;
}
NEMAR_API int nemar_mse_const_bwd(const nemar_tensor* pred, float target, float scale, const float* gscale,
                                  const nemar_tensor* dpred, void* stream) {
  NEMAR_REQUIRE(view_ok(pred) && view_ok(dpred) && gscale && same_shape(pred, dpred) &&
                    pred->dtype == dpred->dtype,
                "mse_const_bwd: bad args");
  TView p = make_view(pred), d = make_view(dpred);
  int64_t total = (int64_t)p.n * p.h * p.w * p.c;
  DISPATCH_DTYPE(p.dtype, T, (mse_const_bwd_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                                 p, target, scale / (float)total, gscale, d)));
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Linear: one warp per output element (n,o); the reduction dimension is read with coalesced loads.
// ---------------------------------------------------------------------------------------------
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ b, int n, int in, int o, int act,
                                  float* __restrict__ y) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n * o) return;
  int nn = warp / o, oo = warp % o;
  const float* xr = x + (int64_t)nn * in;
  const float* wr = w + (int64_t)oo * in;
  float acc = 0.f;
  for (int k = lane; k < in; k += 32) acc = fmaf(__ldg(xr + k), __ldg(wr + k), acc);
  acc = warp_sum(acc);
  if (lane == 0) y[warp] = act_fwd(acc + (b ? __ldg(b + oo) : 0.f), act);
}

// dz = dy*act'(y);  dx[n,i] = sum_o dz[n,o] W[o,i]
__global__ void linear_bwd_dx_kernel(const float* __restrict__ w, const float* __restrict__ y,
                                     const float* __restrict__ dy, int n, int in, int o, int act,
                                     float* __restrict__ dx) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * in) return;
  int nn = (int)(idx / in), ii = (int)(idx % in);
  float acc = 0.f;
  for (int oo = 0; oo < o; ++oo) {
    float dz = __ldg(dy + nn * o + oo) * act_grad_from_y(__ldg(y + nn * o + oo), act);
    acc = fmaf(dz, __ldg(w + (int64_t)oo * in + ii), acc);
  }
  dx[idx] = acc;
}
// dW[o,i] = sum_n dz[n,o] x[n,i];  db[o] = sum_n dz[n,o]
__global__ void linear_bwd_dw_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ dy, int n, int in, int o, int act,
                                     float* __restrict__ dw, float* __restrict__ db) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)o * in) return;
  int oo = (int)(idx / in), ii = (int)(idx % in);
  float acc = 0.f, accb = 0.f;
  for (int nn = 0; nn < n; ++nn) {
    float dz = __ldg(dy + nn * o + oo) * act_grad_from_y(__ldg(y + nn * o + oo), act);
    acc = fmaf(dz, __ldg(x + (int64_t)nn * in + ii), acc);
    accb += dz;
  }
  dw[idx] = acc;
  if (ii == 0 && db) db[oo] = accb;
}

NEMAR_API int nemar_linear_fwd(const float* x, const float* w, const float* b, int n, int i, int o, int act,
                               float* y, void* stream) {
  NEMAR_REQUIRE(x && w && y && n > 0 && i > 0 && o > 0, "linear_fwd: bad args");
  int64_t threads = (int64_t)n * o * 32;
  linear_fwd_kernel<<<(int)ceil_div64(threads, 256), 256, 0, (cudaStream_t)stream>>>(x, w, b, n, i, o, act, y);
  NEMAR_LAUNCH_CHECK();
  return 0;
}
NEMAR_API int nemar_linear_bwd(const float* x, const float* w, const float* y, const float* dy, int n, int i,
                               int o, int act, float* dx, float* dw, float* db, void* stream) {
  NEMAR_REQUIRE(x && w && y && dy && n > 0 && i > 0 && o > 0, "linear_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (dx) linear_bwd_dx_kernel<<<(int)ceil_div64((int64_t)n * i, 256), 256, 0, s>>>(w, y, dy, n, i, o, act, dx);
  if (dw) linear_bwd_dw_kernel<<<(int)ceil_div64((int64_t)o * i, 256), 256, 0, s>>>(x, y, dy, n, i, o, act, dw, db);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Adam over a flat buffer (torch.optim.Adam, amsgrad=False, weight_decay=0, maximize=False):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t numel, float b1, float b2, float eps,
                            float step_size, float inv_bc2_sqrt, float gscale) {
  const int64_t n4 = numel >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = ga[k] * gscale;
      ma[k] = b1 * ma[k] + (1.f - b1) * gr;
      va[k] = b2 * va[k] + (1.f - b2) * gr * gr;
      float denom = sqrtf(va[k]) * inv_bc2_sqrt + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gr = g[i] * gscale;
    float mm = b1 * m[i] + (1.f - b1) * gr;
    float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mm; v[i] = vv;
    p[i] -= step_size * (mm / (sqrtf(vv) * inv_bc2_sqrt + eps));
  }
}

// graph-replayable variant: bias corrections derived on the device from a step counter kept in device memory
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t numel, float lr, float b1, float b2, float eps,
                                const int32_t* __restrict__ step_dev, float gscale) {
  const int t = *step_dev + 1;
  const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
  const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
    float gr = g[i] * gscale;
    float mm = b1 * m[i] + (1.f - b1) * gr;
    float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mm; v[i] = vv;
    p[i] -= step_size * (mm / (sqrtf(vv) * inv_bc2_sqrt + eps));
  }
}
__global__ void counter_inc_kernel(int32_t* c) { *c += 1; }

NEMAR_API int nemar_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1,
                                  float beta2, float eps, int32_t* step_dev, float grad_scale, void* stream) {
  NEMAR_REQUIRE(p && g && m && v && step_dev && numel > 0, "adam_step_dev: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  adam_dev_kernel<<<grid_for(numel, 256), 256, 0, s>>>(p, g, m, v, numel, lr, beta1, beta2, eps, step_dev, grad_scale);
  counter_inc_kernel<<<1, 1, 0, s>>>(step_dev);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

NEMAR_API int nemar_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr,
                              float beta1, float beta2, float eps, int step_count, float grad_scale,
                              void* stream) {
  NEMAR_REQUIRE(p && g && m && v && numel > 0 && step_count >= 1, "adam_step: bad args");
  NEMAR_REQUIRE(((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0,
                "adam_step: buffers must be 16-byte aligned");
  double bc1 = 1.0 - pow((double)beta1, (double)step_count);
  double bc2 = 1.0 - pow((double)beta2, (double)step_count);
  float step_size = (float)((double)lr / bc1);
  float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  adam_kernel<<<grid_for(numel / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, numel, beta1, beta2, eps, step_size, inv_bc2_sqrt, grad_scale);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

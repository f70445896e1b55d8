// vec.cuh — 128-bit vectorised element access for NHWC channel runs.
#pragma once
#include "common.cuh"

template <typename T> struct VecTraits;
template <> struct VecTraits<float> { static constexpr int V = 4; };          // 16 B
template <> struct VecTraits<__nv_bfloat16> { static constexpr int V = 8; };  // 16 B

// load V consecutive elements (V == VecTraits<T>::V -> one 128-bit access, V == 1 -> scalar)
template <typename T, int V> __device__ __forceinline__ void ldv(const T* p, float (&f)[V]) {
  if constexpr (V == 1) {
    f[0] = to_f<T>(p[0]);
  } else if constexpr (sizeof(T) == 4) {
    float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  } else {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
}

template <typename T, int V> __device__ __forceinline__ void stv(T* p, const float (&f)[V]) {
  if constexpr (V == 1) {
    p[0] = from_f<T>(f[0]);
  } else if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  } else {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
}

// can the view be accessed with full vectors of T?
template <typename T> static inline bool view_vec_ok(const nemar_tensor* t) {
  constexpr int V = VecTraits<T>::V;
  return (t->c % V == 0) && (t->cs % V == 0) && (t->coff % V == 0) && ((((uintptr_t)t->ptr) & 15) == 0);
}

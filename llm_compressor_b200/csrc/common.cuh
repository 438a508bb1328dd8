// Shared device / host helpers of liblcb200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "../../include/lcb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "liblcb200 is written for sm_100a (B200) only"
#endif

namespace lcb {

// ------------------------------------------------------------------ host-side error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define LCB_CUDA(expr)                                                         \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::lcb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

// every kernel launch in the library is followed by this: error check + launch accounting
#define LCB_LAUNCH_CHECK()            \
  do {                                \
    ::lcb::count_launch();            \
    LCB_CUDA(cudaGetLastError());     \
  } while (0)

#define LCB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::lcb::set_error(__VA_ARGS__);  \
      return LCB_ERR_INVALID;         \
    }                                 \
  } while (0)

int sm_count();
void count_launch();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ dtype emulation
// Values live in fp32 registers.  R<DT>(v) is the rounding torch applies after every primitive
// op on a tensor of dtype DT: nothing for fp32, round-to-nearest-even for bf16.
template <int DT>
__device__ __forceinline__ float R(float v) {
  if constexpr (DT == LCB_BF16) {
    return __bfloat162float(__float2bfloat16_rn(v));
  } else {
    return v;
  }
}

// clamp(min=1e-5) in the tensor dtype (bf16: 1.00135803e-05)
template <int DT>
__device__ __forceinline__ float scale_floor() {
  if constexpr (DT == LCB_BF16) {
    return __uint_as_float(0x37280000u);
  } else {
    return 1e-5f;
  }
}

// torch.clamp: NaN propagates (fmaxf / fminf would drop it)
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) {
  v = (v < lo) ? lo : v;
  v = (v > hi) ? hi : v;
  return v;
}
__device__ __forceinline__ float clamp_min_nan(float v, float lo) { return (v < lo) ? lo : v; }

// torch.amax / amin: NaN propagates
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ float nan_min(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

__device__ __forceinline__ float exp2i(int n) { return __int_as_float((n + 127) << 23); }  // n in [-126, 127]

template <typename T>
struct DtOf;
template <>
struct DtOf<float> {
  static constexpr int value = LCB_F32;
};
template <>
struct DtOf<__nv_bfloat16> {
  static constexpr int value = LCB_BF16;
};

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// 16-byte vector of T held as floats
template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
};

template <typename T>
__device__ __forceinline__ void load16(const T* p, float (&v)[16 / sizeof(T)]);
template <>
__device__ __forceinline__ void load16<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <typename T>
__device__ __forceinline__ void store16(T* p, const float (&v)[16 / sizeof(T)]);
template <>
__device__ __forceinline__ void store16<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
// values are already bf16-representable (every op was rounded), so packing is a truncation
template <>
__device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    w[i] = (__float_as_uint(v[2 * i]) >> 16) | (__float_as_uint(v[2 * i + 1]) & 0xffff0000u);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// order-preserving float <-> uint key (for atomicMax based global reductions)
__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

}  // namespace lcb

// Numerical profile of a fake-quant op on the GPU (second half): the statistics the reference's
// `record_stats` (ref: quantizers/base.py:30-113) gets by copying both tensors to the CPU and sorting one of them --
// max / min of x and of QDQ(x), the SQNR of the min-max normalised pair, the clipping error -- as two flat reduction
// passes over HBM; the 99th percentile is an exact k-th order statistic by the radix select of masks.cu (the Python
// side drives lcb_select_*), which needs the values as fp32 (lcb_profile_to_f32).
#include "common.cuh"

namespace lcb {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) profile_minmax_kernel(const T* __restrict__ x, const T* __restrict__ q, int64_t n,
                                                            uint32_t* keys) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float red[4][8];
  float xmn = INFINITY, xmx = -INFINITY, qmn = INFINITY, qmx = -INFINITY;
  const int64_t nvec = n / VEC;
  const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = base + u * 256;
    if (i < nvec) {
      float a[VEC], b[VEC];
      load16<T>(x + i * VEC, a);
      load16<T>(q + i * VEC, b);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        xmn = nan_min(xmn, a[j]); xmx = nan_max(xmx, a[j]);
        qmn = nan_min(qmn, b[j]); qmx = nan_max(qmx, b[j]);
      }
    }
  }
  if (blockIdx.x == 0) {  // ragged tail
    for (int64_t i = nvec * VEC + threadIdx.x; i < n; i += 256) {
      const float a = to_f<T>(x[i]), b = to_f<T>(q[i]);
      xmn = nan_min(xmn, a); xmx = nan_max(xmx, a); qmn = nan_min(qmn, b); qmx = nan_max(qmx, b);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmn = nan_min(xmn, __shfl_xor_sync(0xffffffffu, xmn, o)); xmx = nan_max(xmx, __shfl_xor_sync(0xffffffffu, xmx, o));
    qmn = nan_min(qmn, __shfl_xor_sync(0xffffffffu, qmn, o)); qmx = nan_max(qmx, __shfl_xor_sync(0xffffffffu, qmx, o));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = xmn; red[1][wid] = xmx; red[2][wid] = qmn; red[3][wid] = qmx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      xmn = nan_min(xmn, red[0][k]); xmx = nan_max(xmx, red[1][k]); qmn = nan_min(qmn, red[2][k]); qmx = nan_max(qmx, red[3][k]);
    }
    atomicMax(&keys[0], f2key(-xmn));
    atomicMax(&keys[1], f2key(xmx));
    atomicMax(&keys[2], f2key(-qmn));
    atomicMax(&keys[3], f2key(qmx));
  }
}

// sum over elements of ((x - xmin) / (xmax - xmin) - (q - qmin) / (qmax - qmin))^2, fp32 per element, fp64 across
template <typename T>
__global__ void __launch_bounds__(256) profile_sqerr_kernel(const T* __restrict__ x, const T* __restrict__ q, int64_t n,
                                                           const uint32_t* keys, double* out) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ double red[8];
  const float xmn = -key2f(keys[0]), xmx = key2f(keys[1]), qmn = -key2f(keys[2]), qmx = key2f(keys[3]);
  const float xr = __fsub_rn(xmx, xmn), qr = __fsub_rn(qmx, qmn);
  float acc = 0.0f;
  const int64_t nvec = n / VEC;
  const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = base + u * 256;
    if (i < nvec) {
      float a[VEC], b[VEC];
      load16<T>(x + i * VEC, a);
      load16<T>(q + i * VEC, b);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float d = __fsub_rn(__fdiv_rn(__fsub_rn(a[j], xmn), xr), __fdiv_rn(__fsub_rn(b[j], qmn), qr));
        acc = fmaf(d, d, acc);
      }
    }
  }
  if (blockIdx.x == 0) {
    for (int64_t i = nvec * VEC + threadIdx.x; i < n; i += 256) {
      const float d = __fsub_rn(__fdiv_rn(__fsub_rn(to_f<T>(x[i]), xmn), xr), __fdiv_rn(__fsub_rn(to_f<T>(q[i]), qmn), qr));
      acc = fmaf(d, d, acc);
    }
  }
  double s = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) s += red[k];
    atomicAdd(out, s);
  }
}

__global__ void profile_finish_kernel(const uint32_t* keys, const double* acc, float* stats) {
  stats[0] = -key2f(keys[0]); stats[1] = key2f(keys[1]); stats[2] = -key2f(keys[2]); stats[3] = key2f(keys[3]);
  stats[4] = (float)acc[0]; stats[5] = stats[6] = stats[7] = 0.0f;
}

template <typename T>
__global__ void __launch_bounds__(256) to_f32_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = to_f<T>(x[i]);
}

}  // namespace
}  // namespace lcb

using namespace lcb;

// stats (device, 8 floats): [0] min x, [1] max x, [2] min q, [3] max q, [4] sum of squared normalised differences (as
// float), [5..7] reserved.  ws: 64 bytes of scratch.
extern "C" size_t lcb_profile_ws_bytes(void) { return 64; }

extern "C" int lcb_profile_stats(const void* x, const void* q, int dtype, int64_t n, float* stats, void* ws, size_t ws_bytes,
                                 void* stream) {
  LCB_REQUIRE(x && q && stats && n > 0, "lcb_profile_stats: bad arguments");
  LCB_REQUIRE(dtype == LCB_F32 || dtype == LCB_BF16, "lcb_profile_stats: dtype must be LCB_F32 or LCB_BF16");
  LCB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(q)) & 15) == 0,
              "lcb_profile_stats: x and q must be 16-byte aligned");
  if (ws == nullptr || ws_bytes < 64 || (reinterpret_cast<uintptr_t>(ws) & 7)) {
    set_error("lcb_profile_stats: 8-byte aligned workspace of 64 bytes needed");
    return LCB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint32_t* keys = static_cast<uint32_t*>(ws);
  double* acc = reinterpret_cast<double*>(static_cast<char*>(ws) + 32);
  LCB_CUDA(cudaMemsetAsync(ws, 0, 64, st));
  const int vec = dtype == LCB_BF16 ? 8 : 4;
  const unsigned grid = (unsigned)ceil_div(ceil_div(n, vec), 1024);
  if (dtype == LCB_BF16) {
    profile_minmax_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(q), n, keys);
    LCB_LAUNCH_CHECK();
    profile_sqerr_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(q), n, keys, acc);
  } else {
    profile_minmax_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(q), n, keys);
    LCB_LAUNCH_CHECK();
    profile_sqerr_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(q), n, keys, acc);
  }
  LCB_LAUNCH_CHECK();
  profile_finish_kernel<<<1, 1, 0, st>>>(keys, acc, stats);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_profile_to_f32(const void* x, int dtype, float* out, int64_t n, void* stream) {
  LCB_REQUIRE(x && out && n > 0, "lcb_profile_to_f32: bad arguments");
  LCB_REQUIRE(dtype == LCB_F32 || dtype == LCB_BF16, "lcb_profile_to_f32: dtype must be LCB_F32 or LCB_BF16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)ceil_div(n, 256);
  if (dtype == LCB_BF16) to_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), out, n);
  else to_f32_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), out, n);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

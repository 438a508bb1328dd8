// (d) Unstructured-mask scoring and exact top-k selection (no sorting).
//
// ref: pruning/wanda/core.py:116-126 (per-row stable sort -> first int(K*ratio) indices),
//      pruning/ria/core.py:118-126 and pruning/magnitude/core.py:38-43 (global k-th value, <=),
//      pruning/sparsegpt/core.py:201-203 (k-th value of a [N,128] block, <=).
// Selection is an MSB-first radix select on order-preserving uint32 keys: exact k-th order
// statistic in 4 passes of 8 bits; NaN keys sort last like torch.sort.
#include <cstddef>

#include "common.cuh"

namespace lcb {

namespace {

struct SelState {
  uint32_t prefix;
  uint32_t pad;
  unsigned long long krem;
  uint32_t hist[256];
};

static_assert(offsetof(SelState, hist) == LCB_SELECT_HIST_OFFSET, "lcb200.h documents the histogram offset");

__global__ void sel_init_kernel(SelState* s, unsigned long long kth) {
  if (threadIdx.x == 0) { s->prefix = 0; s->krem = kth; }
  s->hist[threadIdx.x] = 0;
}

// histogram of byte `pass` (0 = most significant) among keys whose higher bytes equal prefix
__global__ void __launch_bounds__(256) sel_hist_kernel(const float* __restrict__ v, int64_t n, SelState* s, int pass) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const uint32_t prefix = s->prefix;
  const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = f2key(v[i]);
    if ((key & himask) == (prefix & himask)) atomicAdd(&h[(key >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&s->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void sel_scan_kernel(SelState* s, int pass, float* out) {
  if (threadIdx.x != 0) return;
  const int shift = 24 - 8 * pass;
  unsigned long long k = s->krem, cum = 0;
  int b = 0;
  for (; b < 256; ++b) {
    const unsigned long long c = s->hist[b];
    if (cum + c > k) break;
    cum += c;
  }
  if (b > 255) b = 255;
  s->prefix |= (uint32_t)b << shift;
  s->krem = k - cum;
  for (int i = 0; i < 256; ++i) s->hist[i] = 0;
  if (pass == 3) out[0] = key2f(s->prefix);
}

template <typename T>
__global__ void abs_metric_kernel(const T* __restrict__ w, float* __restrict__ m, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m[i] = fabsf(to_f<T>(w[i]));
}

__global__ void le_mask_kernel(const float* __restrict__ m, const float* thresh, uint8_t* __restrict__ mask, int64_t n) {
  const float th = thresh[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = m[i] <= th;
}

// ---- RIA statistics: column / row sums of |W| accumulated in fp32, rounded once to W's dtype
template <typename T>
__global__ void __launch_bounds__(256) abs_colsum_kernel(const T* __restrict__ w, float* __restrict__ cs, int64_t n,
                                                         int64_t k, bool round_out) {
  constexpr int DT = DtOf<T>::value;
  __shared__ float red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < k)
    for (int64_t r = ty; r < n; r += 8) s += fabsf(to_f<T>(w[r * k + c]));
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < k) {
    for (int i = 1; i < 8; ++i) s += red[i][tx];
    cs[c] = round_out ? R<DT>(s) : s;  // row-sharded callers all-reduce the fp32 partial sums first
  }
}
template <typename T>
__global__ void __launch_bounds__(256) abs_rowsum_kernel(const T* __restrict__ w, float* __restrict__ rs, int64_t n,
                                                         int64_t k) {
  constexpr int DT = DtOf<T>::value;
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  float s = 0.f;
  for (int64_t c = lane; c < k; c += 32) s += fabsf(to_f<T>(w[r * k + c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) rs[r] = R<DT>(s);
}
// metric = (|W|/colsum + |W|/rowsum) [W dtype, one rounding per op] * sqrt(scaler_row)^alpha [fp32]
template <typename T>
__global__ void ria_metric_kernel(const T* __restrict__ w, const float* __restrict__ cs, const float* __restrict__ rs,
                                  const float* __restrict__ srow, float* __restrict__ m, int64_t n, int64_t k,
                                  float alpha) {
  constexpr int DT = DtOf<T>::value;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * k; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k, c = i - r * k;
    const float a = fabsf(to_f<T>(w[i]));
    const float t0 = R<DT>(__fdiv_rn(a, R<DT>(cs[c])));  // R is idempotent on an already rounded sum
    const float t1 = R<DT>(__fdiv_rn(a, rs[r]));
    const float base = R<DT>(__fadd_rn(t0, t1));
    const float sq = __fsqrt_rn(srow[c]);
    const float fac = (alpha == 0.5f) ? __fsqrt_rn(sq) : ((alpha == 1.0f) ? sq : powf(sq, alpha));
    m[i] = __fmul_rn(base, fac);
  }
}

// ---- Wanda: one CTA per row, the row's keys live in registers, radix select in shared memory,
// ties broken by column index (stable sort semantics).
template <typename T, int E>
__global__ void __launch_bounds__(256) wanda_row_kernel(const T* __restrict__ w, const float* __restrict__ srow,
                                                        uint8_t* __restrict__ mask, int64_t n, int64_t k, int kprune) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sh_prefix, sh_krem;
  __shared__ uint32_t warp_cnt[8];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int64_t row = blockIdx.x;
  const int64_t c0 = (int64_t)t * E;
  uint32_t key[E];
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const int64_t c = c0 + j;
    if (c < k) {
      const float m = __fmul_rn(fabsf(to_f<T>(w[row * k + c])), __fsqrt_rn(srow[c]));
      key[j] = f2key(m);
    } else {
      key[j] = 0xffffffffu;  // beyond the row: never selected (kprune <= k)
    }
  }
  if (t == 0) { sh_prefix = 0; sh_krem = (uint32_t)kprune; }
  uint8_t* mrow = mask + row * k;
  if (kprune <= 0) {
#pragma unroll
    for (int j = 0; j < E; ++j)
      if (c0 + j < k) mrow[c0 + j] = 0;
    return;
  }
  // find the key of the element with 0-based rank kprune-1
  if (t == 0) sh_krem = (uint32_t)(kprune - 1);
  for (int pass = 0; pass < 4; ++pass) {
    hist[t] = 0;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const uint32_t prefix = sh_prefix;
    const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
#pragma unroll
    for (int j = 0; j < E; ++j)
      if (c0 + j < k && (key[j] & himask) == (prefix & himask)) atomicAdd(&hist[(key[j] >> shift) & 0xffu], 1u);
    __syncthreads();
    if (t == 0) {
      uint32_t kr = sh_krem, cum = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (cum + hist[b] > kr) break;
        cum += hist[b];
      }
      if (b > 255) b = 255;
      sh_prefix = prefix | ((uint32_t)b << shift);
      sh_krem = kr - cum;
    }
    __syncthreads();
  }
  const uint32_t T_key = sh_prefix;
  const uint32_t tie_need = sh_krem + 1;  // how many elements equal to T_key are pruned (lowest indices first)
  uint32_t my_ties = 0;
#pragma unroll
  for (int j = 0; j < E; ++j) my_ties += (c0 + j < k && key[j] == T_key) ? 1u : 0u;
  // exclusive scan of tie counts over threads (index order == thread order, blocked layout)
  uint32_t incl = my_ties;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_cnt[wid] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (int i = 0; i < wid; ++i) base += warp_cnt[i];
  uint32_t rank = base + incl - my_ties;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    if (c0 + j < k) {
      uint8_t m = 0;
      if (key[j] < T_key) m = 1;
      else if (key[j] == T_key) { m = rank < tie_need ? 1 : 0; ++rank; }
      mrow[c0 + j] = m;
    }
  }
}

template <typename T>
__global__ void apply_mask_kernel(T* __restrict__ w, const uint8_t* __restrict__ mask, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (mask[i]) w[i] = from_f<T>(0.0f);
}

int grid1d(int64_t n) {
  int64_t g = ceil_div(n, 256 * 4);
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace

size_t select_ws_bytes() { return sizeof(SelState) + 64; }

int select_kth_f32(const float* vals, int64_t n, int64_t kth, float* out, void* ws, size_t ws_bytes, cudaStream_t st,
                   lcb_reduce_u32_fn reduce, void* reduce_user) {
  if (ws == nullptr || ws_bytes < select_ws_bytes()) {
    set_error("select_kth_f32: workspace too small");
    return LCB_ERR_WORKSPACE;
  }
  SelState* s = static_cast<SelState*>(ws);
  sel_init_kernel<<<1, 256, 0, st>>>(s, (unsigned long long)kth);
  LCB_LAUNCH_CHECK();
  for (int pass = 0; pass < 4; ++pass) {
    sel_hist_kernel<<<grid1d(n), 256, 0, st>>>(vals, n, s, pass);
    LCB_LAUNCH_CHECK();
    if (reduce != nullptr) {  // row-sharded scores: sum the histograms over the ranks (enqueued on `st` by the caller)
      if (reduce(s->hist, 256, reduce_user, st) != 0) {
        set_error("select_kth_f32: the reduce callback failed");
        return LCB_ERR_CUDA;
      }
    }
    sel_scan_kernel<<<1, 32, 0, st>>>(s, pass, out);
    LCB_LAUNCH_CHECK();
  }
  return LCB_OK;
}

}  // namespace lcb

using namespace lcb;

extern "C" size_t lcb_mask_ws_bytes(int64_t n, int64_t k) {
  return (size_t)(n * k) * sizeof(float) + (size_t)(n + k) * sizeof(float) + 256 + select_ws_bytes();
}

template <typename T>
static int wanda_typed(const T* W, const float* srow, uint8_t* mask, int64_t n, int64_t k, int kprune, cudaStream_t st) {
  const int64_t e = ceil_div(k, 256);
  if (e <= 8) wanda_row_kernel<T, 8><<<(unsigned)n, 256, 0, st>>>(W, srow, mask, n, k, kprune);
  else if (e <= 16) wanda_row_kernel<T, 16><<<(unsigned)n, 256, 0, st>>>(W, srow, mask, n, k, kprune);
  else if (e <= 32) wanda_row_kernel<T, 32><<<(unsigned)n, 256, 0, st>>>(W, srow, mask, n, k, kprune);
  else if (e <= 64) wanda_row_kernel<T, 64><<<(unsigned)n, 256, 0, st>>>(W, srow, mask, n, k, kprune);
  else {
    set_error("lcb_mask_wanda: rows longer than 16384 are not supported");
    return LCB_ERR_UNSUPPORTED;
  }
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_mask_wanda(const void* W, int dtype, const float* scaler_row, uint8_t* mask, int64_t n, int64_t k,
                              double ratio, void* ws, size_t ws_bytes, void* stream) {
  (void)ws; (void)ws_bytes;
  LCB_REQUIRE(W && scaler_row && mask && n > 0 && k > 0, "lcb_mask_wanda: bad arguments");
  const int kprune = (int)((double)k * ratio);  // int(W_metric.shape[1] * sparsity_ratio)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == LCB_BF16 ? wanda_typed(static_cast<const __nv_bfloat16*>(W), scaler_row, mask, n, k, kprune, st)
                           : wanda_typed(static_cast<const float*>(W), scaler_row, mask, n, k, kprune, st);
}

static int threshold_mask(float* metric, int64_t numel, double ratio, uint8_t* mask, float* thresh, void* sel_ws,
                          cudaStream_t st) {
  int64_t kth = (int64_t)((double)numel * ratio);  // int(W.numel() * sparsity_ratio)
  if (kth >= numel) kth = numel - 1;
  int rc = select_kth_f32(metric, numel, kth, thresh, sel_ws, select_ws_bytes(), st, nullptr, nullptr);
  if (rc != LCB_OK) return rc;
  le_mask_kernel<<<grid1d(numel), 256, 0, st>>>(metric, thresh, mask, numel);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_mask_magnitude(const void* W, int dtype, uint8_t* mask, int64_t n, int64_t k, double ratio, void* ws,
                                  size_t ws_bytes, void* stream) {
  LCB_REQUIRE(W && mask && n > 0 && k > 0, "lcb_mask_magnitude: bad arguments");
  if (ws == nullptr || ws_bytes < lcb_mask_ws_bytes(n, k)) {
    set_error("lcb_mask_magnitude: workspace of %zu bytes needed", lcb_mask_ws_bytes(n, k));
    return LCB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* metric = static_cast<float*>(ws);
  float* thresh = metric + n * k + n + k;
  void* sel_ws = thresh + 64;
  if (dtype == LCB_BF16) abs_metric_kernel<<<grid1d(n * k), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(W), metric, n * k);
  else abs_metric_kernel<<<grid1d(n * k), 256, 0, st>>>(static_cast<const float*>(W), metric, n * k);
  LCB_LAUNCH_CHECK();
  return threshold_mask(metric, n * k, ratio, mask, thresh, sel_ws, st);
}

template <typename T>
static int ria_typed(const T* W, const float* srow, float* metric, float* cs, float* rs, int64_t n, int64_t k,
                     float alpha, cudaStream_t st) {
  abs_colsum_kernel<T><<<(unsigned)ceil_div(k, 32), 256, 0, st>>>(W, cs, n, k, true);
  LCB_LAUNCH_CHECK();
  abs_rowsum_kernel<T><<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(W, rs, n, k);
  LCB_LAUNCH_CHECK();
  ria_metric_kernel<T><<<grid1d(n * k), 256, 0, st>>>(W, cs, rs, srow, metric, n, k, alpha);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_mask_ria(const void* W, int dtype, const float* scaler_row, uint8_t* mask, int64_t n, int64_t k,
                            double ratio, float alpha, void* ws, size_t ws_bytes, void* stream) {
  LCB_REQUIRE(W && scaler_row && mask && n > 0 && k > 0, "lcb_mask_ria: bad arguments");
  if (ws == nullptr || ws_bytes < lcb_mask_ws_bytes(n, k)) {
    set_error("lcb_mask_ria: workspace of %zu bytes needed", lcb_mask_ws_bytes(n, k));
    return LCB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* metric = static_cast<float*>(ws);
  float* cs = metric + n * k;
  float* rs = cs + k;
  float* thresh = rs + n;
  void* sel_ws = thresh + 64;
  int rc = dtype == LCB_BF16
               ? ria_typed(static_cast<const __nv_bfloat16*>(W), scaler_row, metric, cs, rs, n, k, alpha, st)
               : ria_typed(static_cast<const float*>(W), scaler_row, metric, cs, rs, n, k, alpha, st);
  if (rc != LCB_OK) return rc;
  return threshold_mask(metric, n * k, ratio, mask, thresh, sel_ws, st);
}

// ---- phase API for row-sharded global thresholds: the caller all-reduces hist[256] between
// lcb_select_hist and lcb_select_scan, and the fp32 column sums between lcb_ria_sums and lcb_ria_metric.
extern "C" size_t lcb_select_state_bytes(void) { return sizeof(SelState); }

extern "C" int lcb_select_init(void* state, int64_t kth, void* stream) {
  LCB_REQUIRE(state != nullptr && kth >= 0, "lcb_select_init: bad arguments");
  sel_init_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<SelState*>(state), (unsigned long long)kth);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_select_hist(const float* vals, int64_t n, void* state, int pass, void* stream) {
  LCB_REQUIRE(state != nullptr && n >= 0 && pass >= 0 && pass < 4 && (vals != nullptr || n == 0), "lcb_select_hist: bad arguments");
  if (n == 0) return LCB_OK;  // an empty shard contributes nothing
  sel_hist_kernel<<<grid1d(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(vals, n, static_cast<SelState*>(state), pass);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_select_scan(void* state, int pass, float* thresh, void* stream) {
  LCB_REQUIRE(state != nullptr && thresh != nullptr && pass >= 0 && pass < 4, "lcb_select_scan: bad arguments");
  sel_scan_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<SelState*>(state), pass, thresh);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_metric_magnitude(const void* W, int dtype, float* metric, int64_t numel, void* stream) {
  LCB_REQUIRE(numel >= 0 && (numel == 0 || (W && metric)), "lcb_metric_magnitude: bad arguments");
  if (numel == 0) return LCB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == LCB_BF16) abs_metric_kernel<<<grid1d(numel), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(W), metric, numel);
  else abs_metric_kernel<<<grid1d(numel), 256, 0, st>>>(static_cast<const float*>(W), metric, numel);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_ria_sums(const void* W, int dtype, float* colsum_partial, float* rowsum, int64_t n, int64_t k,
                            void* stream) {
  LCB_REQUIRE(W && colsum_partial && rowsum && n > 0 && k > 0, "lcb_ria_sums: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == LCB_BF16) {
    const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(W);
    abs_colsum_kernel<__nv_bfloat16><<<(unsigned)ceil_div(k, 32), 256, 0, st>>>(w, colsum_partial, n, k, false);
    abs_rowsum_kernel<__nv_bfloat16><<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(w, rowsum, n, k);
  } else {
    const float* w = static_cast<const float*>(W);
    abs_colsum_kernel<float><<<(unsigned)ceil_div(k, 32), 256, 0, st>>>(w, colsum_partial, n, k, false);
    abs_rowsum_kernel<float><<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(w, rowsum, n, k);
  }
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_ria_metric(const void* W, int dtype, const float* colsum, const float* rowsum, const float* scaler_row,
                              float* metric, int64_t n, int64_t k, float alpha, void* stream) {
  LCB_REQUIRE(W && colsum && rowsum && scaler_row && metric && n > 0 && k > 0, "lcb_ria_metric: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == LCB_BF16)
    ria_metric_kernel<__nv_bfloat16><<<grid1d(n * k), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(W), colsum, rowsum,
                                                                   scaler_row, metric, n, k, alpha);
  else
    ria_metric_kernel<float><<<grid1d(n * k), 256, 0, st>>>(static_cast<const float*>(W), colsum, rowsum, scaler_row, metric,
                                                            n, k, alpha);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_mask_le(const float* metric, const float* thresh, uint8_t* mask, int64_t numel, void* stream) {
  LCB_REQUIRE(thresh && numel >= 0 && (numel == 0 || (metric && mask)), "lcb_mask_le: bad arguments");
  if (numel == 0) return LCB_OK;
  le_mask_kernel<<<grid1d(numel), 256, 0, static_cast<cudaStream_t>(stream)>>>(metric, thresh, mask, numel);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_apply_mask(void* W, int dtype, const uint8_t* mask, int64_t numel, void* stream) {
  LCB_REQUIRE(W && mask && numel >= 0, "lcb_apply_mask: bad arguments");
  if (numel == 0) return LCB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == LCB_BF16) apply_mask_kernel<<<grid1d(numel), 256, 0, st>>>(static_cast<__nv_bfloat16*>(W), mask, numel);
  else apply_mask_kernel<<<grid1d(numel), 256, 0, st>>>(static_cast<float*>(W), mask, numel);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

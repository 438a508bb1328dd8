// Randomised Hadamard rotation of weight rows (config 5's R1 / R2 stage).
//
// The reference materialises R = diag(s) * matmul_hadU(I) as a dense fp64 [n, n] matrix
// (ref: spinquant/rotation_utils.py:40-45, hadamard_utils.py:88-111) and rotates every weight with an fp64 GEMM
// (rotation_utils.py:57-113, hadamard_utils.py:135-172): 2 n flop per weight element and an [n, n] operand.
// Row-wise this is  y = T(s .* x) / sqrt(n)  with  T = (H_K (x) I_L)(I_K (x) H_L),  n = K * L,  L = 2^m  and H_K one of
// the reference's fixed Hadamard matrices (hadamard_utils.py:17-85) -- a fast Walsh-Hadamard transform:
// log2(L) + K adds per element instead of n multiply-adds, no rotation matrix in memory, HBM bound
// (read + write of the weight).  This file is that transform:
//   * one CTA owns `rpc` rows in shared memory (padded against bank conflicts); the radix-2 stages run in
//     register-blocked passes of up to 4 bits (16 values per thread), the first pass straight from global
//     memory and the last one straight to it;
//   * the H_K stage keeps the K strided values of one column in registers (sign table in the kernel parameters);
//   * accumulation in fp64 (default: results round to the same bf16 / fp32 values as the reference's fp64 GEMM
//     up to ~1e-16 relative) or fp32.
#include <algorithm>

#include "common.cuh"

namespace lcb {

namespace {

constexpr int HAD_THREADS = 256;
constexpr int HAD_MAXK = 64;

struct HadArgs {
  const void* x;
  void* y;
  const float* signs;  // [n] or null
  int64_t rows;
  int n, m, K, rpc;
  int dt_in, dt_out;
  double divisor;  // the reference divides by float32(sqrt(n)) (hadamard_utils.py:111)
  unsigned long long hk[HAD_MAXK];  // bit a of hk[a'] set  <=>  H_K[a'][a] == -1
};

__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }

template <typename Acc>
__device__ __forceinline__ Acc load_in(const void* p, int dt, int64_t i) {
  if (dt == LCB_BF16) return (Acc)__bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  if (dt == LCB_F32) return (Acc) static_cast<const float*>(p)[i];
  return (Acc) static_cast<const double*>(p)[i];
}
template <typename Acc>
__device__ __forceinline__ void store_out(void* p, int dt, int64_t i, Acc v) {
  if (dt == LCB_BF16) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn((float)v);  // torch: double -> float -> bf16
  else if (dt == LCB_F32) static_cast<float*>(p)[i] = (float)v;
  else static_cast<double*>(p)[i] = (double)v;
}
template <>
__device__ __forceinline__ void store_out<float>(void* p, int dt, int64_t i, float v) {
  if (dt == LCB_BF16) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else if (dt == LCB_F32) static_cast<float*>(p)[i] = v;
  else static_cast<double*>(p)[i] = (double)v;
}

// in-register radix-2 stages over 2^R values (natural order, like the view(.., 2, ..) steps of matmul_hadU)
template <typename Acc, int R>
__device__ __forceinline__ void butterfly(Acc (&v)[16]) {
#pragma unroll
  for (int s = 0; s < R; ++s) {
    const int h = 1 << s;
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
      if ((j & h) == 0) {
        const Acc a = v[j], b = v[j | h];
        v[j] = a + b;
        v[j | h] = a - b;
      }
    }
  }
}

template <typename Acc, int R>
__device__ __forceinline__ void first_pass(const HadArgs& a, Acc* sm, int64_t row0, int nr, bool to_global) {
  const int groups = (nr * a.n) >> R;
  for (int g = threadIdx.x; g < groups; g += HAD_THREADS) {
    const int flat = g << R;
    const int row = flat / a.n, i = flat - row * a.n;
    const int64_t gbase = (row0 + row) * (int64_t)a.n + i;
    Acc v[16];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
      Acc t = load_in<Acc>(a.x, a.dt_in, gbase + j);
      if (a.signs != nullptr) t *= (Acc)a.signs[i + j];
      v[j] = t;
    }
    butterfly<Acc, R>(v);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
      if (to_global) store_out<Acc>(a.y, a.dt_out, gbase + j, v[j] / (Acc)a.divisor);
      else sm[pad(flat + j)] = v[j];
    }
  }
}

template <typename Acc, int R>
__device__ __forceinline__ void mid_pass(const HadArgs& a, Acc* sm, int64_t row0, int nr, int b0, bool to_global) {
  const int groups = (nr * a.n) >> R;
  const int lomask = (1 << b0) - 1;
  for (int g = threadIdx.x; g < groups; g += HAD_THREADS) {
    const int base = ((g >> b0) << (b0 + R)) | (g & lomask);
    Acc v[16];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) v[j] = sm[pad(base + (j << b0))];
    butterfly<Acc, R>(v);
    if (to_global) {
      const int row = base / a.n, i = base - row * a.n;  // the whole group lies in one row (b0 + R <= m)
      const int64_t gbase = (row0 + row) * (int64_t)a.n + i;
#pragma unroll
      for (int j = 0; j < (1 << R); ++j) store_out<Acc>(a.y, a.dt_out, gbase + ((int64_t)j << b0), v[j] / (Acc)a.divisor);
    } else {
#pragma unroll
      for (int j = 0; j < (1 << R); ++j) sm[pad(base + (j << b0))] = v[j];
    }
  }
}

// y[a', b] = sum_a H_K[a'][a] * v[a, b]; KT > 0: the K values of a column live in registers
template <typename Acc, int KT>
__device__ __forceinline__ void hadk_pass(const HadArgs& a, const Acc* sm, int64_t row0, int nr) {
  const int L = 1 << a.m;
  for (int q = threadIdx.x; q < nr * L; q += HAD_THREADS) {
    const int row = q >> a.m, b = q & (L - 1);
    const int sbase = row * a.n + b;
    const int64_t gbase = (row0 + row) * (int64_t)a.n + b;
    if constexpr (KT > 0) {
      Acc v[KT];
#pragma unroll
      for (int c = 0; c < KT; ++c) v[c] = sm[pad(sbase + (c << a.m))];
#pragma unroll 4
      for (int r = 0; r < KT; ++r) {
        const unsigned long long bits = a.hk[r];
        Acc acc = 0;
#pragma unroll
        for (int c = 0; c < KT; ++c) acc += ((bits >> c) & 1ull) ? -v[c] : v[c];
        store_out<Acc>(a.y, a.dt_out, gbase + ((int64_t)r << a.m), acc / (Acc)a.divisor);
      }
    } else {
      for (int r = 0; r < a.K; ++r) {
        const unsigned long long bits = a.hk[r];
        Acc acc = 0;
        for (int c = 0; c < a.K; ++c) {
          const Acc t = sm[pad(sbase + (c << a.m))];
          acc += ((bits >> c) & 1ull) ? -t : t;
        }
        store_out<Acc>(a.y, a.dt_out, gbase + ((int64_t)r << a.m), acc / (Acc)a.divisor);
      }
    }
  }
}

template <typename Acc, int KT>
__global__ void __launch_bounds__(HAD_THREADS) hadamard_rows_kernel(const __grid_constant__ HadArgs a) {
  extern __shared__ __align__(16) unsigned char had_smem[];
  Acc* sm = reinterpret_cast<Acc*>(had_smem);
  const int r0 = a.m < 4 ? a.m : 4;
  for (int64_t row0 = (int64_t)blockIdx.x * a.rpc; row0 < a.rows; row0 += (int64_t)gridDim.x * a.rpc) {
    const int nr = (int)min((int64_t)a.rpc, a.rows - row0);
    const bool only = (a.K == 1 && a.m == r0);  // the first pass already finishes the transform
    switch (r0) {
      case 4: first_pass<Acc, 4>(a, sm, row0, nr, only); break;
      case 3: first_pass<Acc, 3>(a, sm, row0, nr, only); break;
      case 2: first_pass<Acc, 2>(a, sm, row0, nr, only); break;
      case 1: first_pass<Acc, 1>(a, sm, row0, nr, only); break;
      default: first_pass<Acc, 0>(a, sm, row0, nr, only); break;
    }
    __syncthreads();
    for (int done = r0; done < a.m;) {
      const int r = (a.m - done) < 4 ? (a.m - done) : 4;
      const bool last = (done + r == a.m) && a.K == 1;
      switch (r) {
        case 4: mid_pass<Acc, 4>(a, sm, row0, nr, done, last); break;
        case 3: mid_pass<Acc, 3>(a, sm, row0, nr, done, last); break;
        case 2: mid_pass<Acc, 2>(a, sm, row0, nr, done, last); break;
        default: mid_pass<Acc, 1>(a, sm, row0, nr, done, last); break;
      }
      __syncthreads();
      done += r;
    }
    if (a.K > 1) {
      hadk_pass<Acc, KT>(a, sm, row0, nr);
      __syncthreads();
    }
  }
}

template <typename Acc, int KT>
int launch_had(const HadArgs& a, int grid, size_t smem, cudaStream_t st) {
  LCB_CUDA(cudaFuncSetAttribute(hadamard_rows_kernel<Acc, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hadamard_rows_kernel<Acc, KT><<<grid, HAD_THREADS, smem, st>>>(a);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

template <typename Acc>
int launch_had_k(const HadArgs& a, int grid, size_t smem, cudaStream_t st) {
  switch (a.K) {
    case 12: return launch_had<Acc, 12>(a, grid, smem, st);
    case 20: return launch_had<Acc, 20>(a, grid, smem, st);
    case 28: return launch_had<Acc, 28>(a, grid, smem, st);
    case 40: return launch_had<Acc, 40>(a, grid, smem, st);
    default: return launch_had<Acc, 0>(a, grid, smem, st);
  }
}

}  // namespace
}  // namespace lcb

using namespace lcb;

extern "C" int lcb_hadamard_rows(const void* x, int dtype_in, void* y, int dtype_out, int64_t rows, int64_t n,
                                 const float* signs, const uint64_t* hadk_bits, int K, double divisor, int acc64,
                                 void* stream) {
  LCB_REQUIRE(x && y && rows >= 0 && n > 0 && K >= 1 && K <= HAD_MAXK && divisor != 0.0, "lcb_hadamard_rows: bad arguments");
  LCB_REQUIRE(dtype_in >= LCB_F32 && dtype_in <= LCB_F64 && dtype_out >= LCB_F32 && dtype_out <= LCB_F64,
              "lcb_hadamard_rows: dtype must be LCB_F32, LCB_BF16 or LCB_F64");
  LCB_REQUIRE(K == 1 || hadk_bits != nullptr, "lcb_hadamard_rows: K > 1 needs the H_K sign table");
  LCB_REQUIRE(n % K == 0, "lcb_hadamard_rows: n must be K * 2^m");
  int64_t L = n / K;
  int m = 0;
  while ((1ll << m) < L) ++m;
  LCB_REQUIRE((1ll << m) == L, "lcb_hadamard_rows: n / K must be a power of two");
  if (rows == 0) return LCB_OK;
  const size_t esz = acc64 ? sizeof(double) : sizeof(float);
  const int64_t max_elems = (int64_t)(200 * 1024 / esz) * 16 / 17;
  if (n > max_elems) {
    set_error("lcb_hadamard_rows: n = %lld exceeds the shared-memory tile (%lld)", (long long)n, (long long)max_elems);
    return LCB_ERR_UNSUPPORTED;
  }
  HadArgs a{};
  a.x = x; a.y = y; a.signs = signs; a.rows = rows; a.n = (int)n; a.m = m; a.K = K;
  a.rpc = (int)std::max<int64_t>(1, 4096 / n);
  a.dt_in = dtype_in; a.dt_out = dtype_out; a.divisor = divisor;
  for (int i = 0; i < K && K > 1; ++i) a.hk[i] = hadk_bits[i];
  const int64_t tile = (int64_t)a.rpc * n;
  const size_t smem = (size_t)(tile + (tile >> 4) + 16) * esz;
  const int64_t tiles = ceil_div(rows, a.rpc);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count() * per_sm);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return acc64 ? launch_had_k<double>(a, grid, smem, st) : launch_had_k<float>(a, grid, smem, st);
}

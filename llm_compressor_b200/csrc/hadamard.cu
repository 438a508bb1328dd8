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
#include <cstdlib>
#include <utility>

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
  int raw_bytes;  // tiled kernel: size of the raw staging area in front of the work buffer
  int sgn_off;    // tiled kernel: byte offset of the row of sign XOR masks
  double divisor;  // the reference divides by float32(sqrt(n)) (hadamard_utils.py:111)
  double rcp;      // 1 / divisor
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


// K > 64 (the reference's had108 / had140 / had156 / had172 blocks: Llama-1 / Llama-2 intermediate sizes such as
// 11008 = 172 * 64): the sign table no longer fits one 64-bit word per row; it travels as HAD_BIGW words per row in a larger
// parameter block of its own kernel, so that the launches of the common shapes keep their small parameter block.
constexpr int HAD_BIGK = 172;
constexpr int HAD_BIGW = 3;
struct HadArgsBig {
  HadArgs b;
  unsigned long long hkx[HAD_BIGK * HAD_BIGW];  // bit (a & 63) of hkx[a' * HAD_BIGW + (a >> 6)] set  <=>  H_K[a'][a] == -1
};

template <typename Acc>
__global__ void __launch_bounds__(HAD_THREADS) hadamard_rows_bigk_kernel(const __grid_constant__ HadArgsBig ab) {
  extern __shared__ __align__(16) unsigned char had_smem[];
  Acc* sm = reinterpret_cast<Acc*>(had_smem);
  const HadArgs& a = ab.b;
  const int r0 = a.m < 4 ? a.m : 4;
  const int L = 1 << a.m;
  for (int64_t row0 = (int64_t)blockIdx.x * a.rpc; row0 < a.rows; row0 += (int64_t)gridDim.x * a.rpc) {
    const int nr = (int)min((int64_t)a.rpc, a.rows - row0);
    switch (r0) {
      case 4: first_pass<Acc, 4>(a, sm, row0, nr, false); break;
      case 3: first_pass<Acc, 3>(a, sm, row0, nr, false); break;
      case 2: first_pass<Acc, 2>(a, sm, row0, nr, false); break;
      case 1: first_pass<Acc, 1>(a, sm, row0, nr, false); break;
      default: first_pass<Acc, 0>(a, sm, row0, nr, false); break;
    }
    __syncthreads();
    for (int done = r0; done < a.m;) {
      const int r = (a.m - done) < 4 ? (a.m - done) : 4;
      switch (r) {
        case 4: mid_pass<Acc, 4>(a, sm, row0, nr, done, false); break;
        case 3: mid_pass<Acc, 3>(a, sm, row0, nr, done, false); break;
        case 2: mid_pass<Acc, 2>(a, sm, row0, nr, done, false); break;
        default: mid_pass<Acc, 1>(a, sm, row0, nr, done, false); break;
      }
      __syncthreads();
      done += r;
    }
    // y[a', b] = sum_a H_K[a'][a] v[a, b]
    for (int q = threadIdx.x; q < nr * L; q += HAD_THREADS) {
      const int row = q >> a.m, b = q & (L - 1);
      const int sbase = row * a.n + b;
      const int64_t gbase = (row0 + row) * (int64_t)a.n + b;
      for (int r = 0; r < a.K; ++r) {
        const unsigned long long* bits = ab.hkx + r * HAD_BIGW;
        Acc acc = 0;
        for (int c = 0; c < a.K; ++c) {
          const Acc t = sm[pad(sbase + (c << a.m))];
          acc += ((bits[c >> 6] >> (c & 63)) & 1ull) ? -t : t;
        }
        store_out<Acc>(a.y, a.dt_out, gbase + ((int64_t)r << a.m), acc / (Acc)a.divisor);
      }
    }
    __syncthreads();
  }
}

template <typename Acc>
int launch_had_big(const HadArgsBig& ab, int grid, size_t smem, cudaStream_t st) {
  LCB_CUDA(cudaFuncSetAttribute(hadamard_rows_bigk_kernel<Acc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hadamard_rows_bigk_kernel<Acc><<<grid, HAD_THREADS, smem, st>>>(ab);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Tiled kernel (m >= 5, bf16 / fp32 input, K in {1, 12, 40}): all global traffic is 16-byte vectors and the inner
// loops are (nearly) one instruction per add.  The tile (rpc whole rows, one contiguous chunk) is copied raw into
// shared memory with the sign vector XOR-ed into the sign bits on the way (signs packed to a bit mask once per CTA);
// the stages then run in an order chosen for the memory system, not the reference's -- the factors of T act on
// different index bits and commute (with bf16 inputs every fp64 partial sum is exact, so the order does not change a
// single bit; fp32 accumulation moves within its rounding error):
//   H_K stage first (raw tile -> work buffer; the K strided values of a column in registers, signs of H_K folded at
//   compile time from the Paley construction; H_40 = H_2 (x) H_20 as one butterfly + two H_20), then the radix-2
//   stages from the highest bit down in register-blocked passes of up to 5 bits (b0 >= 5: the padded address is
//   linear in the register index), and the lowest 5 bits last so that every thread ends with 32 CONTIGUOUS outputs.
//   Work-buffer index i lives at i + (i >> 5): every access pattern above (32 consecutive indices per warp, or lane g
//   at 32 g + j) hits 32 distinct banks.
__device__ __forceinline__ int pad32(int i) { return i + (i >> 5); }  // one pad word per 32: see the bank analysis in DESIGN.md 3.7

constexpr int chi_c(int a, int q) {  // Legendre symbol of a mod q (q prime)
  a %= q;
  if (a < 0) a += q;
  if (a == 0) return 0;
  long long r = 1, b = a;
  for (int e = (q - 1) / 2; e > 0; e >>= 1) {
    if (e & 1) r = (r * b) % q;
    b = (b * b) % q;
  }
  return r == 1 ? 1 : -1;
}
// row mask (bit c set <=> entry (r, c) == -1) of the Paley-I matrix of order K = q + 1 the reference tabulates as
// had12 / had20 (hadamard_utils.py; generated the same way in llm_compressor_b200/hadamard.py and checked against
// the reference's tables by tests/test_hadamard.py)
constexpr unsigned long long paley1_row_mask(int K, int r) {
  unsigned long long m = 0;
  for (int c = 0; c < K; ++c) {
    bool neg = false;
    if (r == 0) neg = c > 0;
    else if (c == 0 || c == r) neg = false;
    else neg = chi_c(c - r, K - 1) == 1;  // entry = -chi(c - r)
    if (neg) m |= 1ull << c;
  }
  return m;
}

// H_K v for a Paley-I matrix, K = 4 G: the entries are +-1, so within a group of four inputs (a, b, c, d) every row needs
// one of the eight values a +- b +- c +- d (up to an overall sign).  Build those once per group (12 adds) and each output
// is a signed sum of G of them: 12 G + K (G - 1) adds instead of K (K - 1) -- 60 instead of 132 for H_12, 140 instead of
// 380 for H_20.  Signs and indices fold at compile time (the masks are constexpr, the loops fully unrolled).
template <typename Acc, int KB>
__device__ __forceinline__ void paley_groups(const Acc (&v)[KB], Acc (&cmb)[KB / 4][8]) {
#pragma unroll
  for (int g = 0; g < KB / 4; ++g) {
    const Acc a = v[4 * g], b = v[4 * g + 1], c = v[4 * g + 2], d = v[4 * g + 3];
    const Acc s0 = a + b, d0 = a - b, s1 = c + d, d1 = c - d;
    cmb[g][0] = s0 + s1; cmb[g][1] = d0 + s1;   // index = [b negative] + 2 * (2 * [c negative] + [d negative])
    cmb[g][2] = s0 + d1; cmb[g][3] = d0 + d1;
    cmb[g][4] = s0 - d1; cmb[g][5] = d0 - d1;
    cmb[g][6] = s0 - s1; cmb[g][7] = d0 - s1;
  }
}
template <typename Acc, int KB, int R>
__device__ __forceinline__ Acc paley_dot(const Acc (&cmb)[KB / 4][8]) {
  constexpr unsigned long long M = paley1_row_mask(KB, R);
  Acc acc = 0;
#pragma unroll
  for (int g = 0; g < KB / 4; ++g) {
    const unsigned m4 = (unsigned)((M >> (4 * g)) & 0xFull);
    const bool flip = (m4 & 1u) != 0;                       // a enters negated: use the mirrored pattern, negated
    const unsigned q = flip ? (m4 ^ 0xFu) : m4;
    const int idx = (int)((q >> 1) & 1u) + 2 * (int)(2 * ((q >> 2) & 1u) + ((q >> 3) & 1u));
    const Acc t = cmb[g][idx];
    if (g == 0) acc = flip ? -t : t;
    else acc = flip ? acc - t : acc + t;
  }
  return acc;
}
template <typename Acc, int KB, int... Rs>
__device__ __forceinline__ void paley_store(const Acc (&v)[KB], Acc* p, int pstride, std::integer_sequence<int, Rs...>) {
  Acc cmb[KB / 4][8];
  paley_groups<Acc, KB>(v, cmb);
  ((p[Rs * pstride] = paley_dot<Acc, KB, Rs>(cmb)), ...);
}

template <typename Acc, typename TIn>
__device__ __forceinline__ Acc raw_at(const unsigned char* raw, int idx) {
  if constexpr (sizeof(TIn) == 2) return (Acc)__uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(raw)[idx] << 16);
  else return (Acc) reinterpret_cast<const float*>(raw)[idx];
}

template <typename Acc, int N>
__device__ __forceinline__ void butterfly_n(Acc (&v)[N]) {
#pragma unroll
  for (int h = 1; h < N; h <<= 1) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if ((j & h) == 0) {
        const Acc p = v[j], q = v[j | h];
        v[j] = p + q;
        v[j | h] = p - q;
      }
    }
  }
}

// scale (multiply by the reciprocal: its 2^-53 / 2^-24 error is far below the output rounding; fp64 output divides
// like the reference), convert and store 16 contiguous outputs
template <typename Acc>
__device__ __forceinline__ void store16_out(void* y, int dt, int64_t e0, const Acc (&v)[16], double divisor, double rcpd) {
  if (dt == LCB_BF16) {
    const Acc rcp = (Acc)rcpd;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // torch: double -> float -> bf16; one packed cvt.rn.bf16x2.f32 per pair
      const __nv_bfloat162 pr = __floats2bfloat162_rn((float)(v[2 * i] * rcp), (float)(v[2 * i + 1] * rcp));
      w[i] = *reinterpret_cast<const uint32_t*>(&pr);
    }
    uint4* p = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(y) + e0);
    p[0] = make_uint4(w[0], w[1], w[2], w[3]);
    p[1] = make_uint4(w[4], w[5], w[6], w[7]);
  } else if (dt == LCB_F32) {
    const Acc rcp = (Acc)rcpd;
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(y) + e0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      p[i] = make_float4((float)(v[4 * i] * rcp), (float)(v[4 * i + 1] * rcp), (float)(v[4 * i + 2] * rcp),
                         (float)(v[4 * i + 3] * rcp));
  } else {
    double2* p = reinterpret_cast<double2*>(static_cast<double*>(y) + e0);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = make_double2((double)v[2 * i] / divisor, (double)v[2 * i + 1] / divisor);
  }
}

// Shape known at compile time (NC = n, B0C = first bit of a pass) for the common widths: strides become immediates and
// the per-group index arithmetic folds; NC == 0 / B0C == 0 take the values from the arguments.
// radix-2 stages on bits [b0, b0 + R), b0 >= 5: N = 2^R values per thread at a constant padded stride
template <typename Acc, typename TIn, int R, bool SRC_RAW, int NC = 0, int B0C = 0>
__device__ __forceinline__ void mid_bits(const HadArgs& a, const unsigned char* raw, Acc* sm, int nr, int b0_rt) {
  constexpr int N = 1 << R;
  const int n = NC > 0 ? NC : a.n;
  const int b0 = B0C > 0 ? B0C : b0_rt;
  const int groups = (nr * n) >> R;
  const int lomask = (1 << b0) - 1;
  const int step = 1 << b0, pstep = step + (step >> 5);
  for (int g = threadIdx.x; g < groups; g += HAD_THREADS) {
    const int base = ((g >> b0) << (b0 + R)) | (g & lomask);
    Acc* p = sm + pad32(base);
    Acc v[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if constexpr (SRC_RAW) v[j] = raw_at<Acc, TIn>(raw, base + j * step);
      else v[j] = p[j * pstep];
    }
    butterfly_n<Acc, N>(v);
#pragma unroll
    for (int j = 0; j < N; ++j) p[j * pstep] = v[j];
  }
}

// the lowest 5 bits: 32 contiguous values per thread, straight to global memory
template <typename Acc, typename TIn, bool SRC_RAW, int NC = 0>
__device__ __forceinline__ void low_bits_out(const HadArgs& a, const unsigned char* raw, const Acc* sm, int64_t row0, int nr) {
  const int n = NC > 0 ? NC : a.n;
  const int groups = (nr * n) >> 5;
  for (int g = threadIdx.x; g < groups; g += HAD_THREADS) {
    Acc v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if constexpr (SRC_RAW) v[j] = raw_at<Acc, TIn>(raw, 32 * g + j);
      else v[j] = sm[33 * g + j];  // pad32(32 g + j): lane g, value j -> bank (g + j) % 32
    }
    butterfly_n<Acc, 32>(v);
    Acc lo[16], hi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { lo[j] = v[j]; hi[j] = v[16 + j]; }
    const int64_t e0 = row0 * (int64_t)n + 32 * g;
    store16_out<Acc>(a.y, a.dt_out, e0, lo, a.divisor, a.rcp);
    store16_out<Acc>(a.y, a.dt_out, e0 + 16, hi, a.divisor, a.rcp);
  }
}

// H_K across the leading index (stride L = 2^m), raw tile -> work buffer.  L >= 32, so the padded index of
// sbase + r L is pad32(sbase) + r (L + L / 32): one base address and a constant stride per thread.
template <typename Acc, typename TIn, int KT, int MC = 0>
__device__ __forceinline__ void hadk_first(const HadArgs& a, const unsigned char* raw, Acc* sm, int nr) {
  const int m = MC > 0 ? MC : a.m;
  const int n = MC > 0 ? (KT << MC) : a.n;
  const int L = 1 << m;
  const int pstride = L + (L >> 5);
  for (int q = threadIdx.x; q < nr * L; q += HAD_THREADS) {
    const int row = q >> m, b = q & (L - 1);
    const int sbase = row * n + b;
    const TIn* rp = reinterpret_cast<const TIn*>(raw) + sbase;
    Acc* wp = sm + pad32(sbase);
    if constexpr (KT == 12) {
      Acc v[12];
#pragma unroll
      for (int c = 0; c < 12; ++c) v[c] = raw_at<Acc, TIn>(reinterpret_cast<const unsigned char*>(rp), c * L);
      paley_store<Acc, 12>(v, wp, pstride, std::make_integer_sequence<int, 12>{});
    } else if constexpr (KT == 40) {
      Acc v0[20], v1[20];
#pragma unroll
      for (int c = 0; c < 20; ++c) {
        const Acc p = raw_at<Acc, TIn>(reinterpret_cast<const unsigned char*>(rp), c * L);
        const Acc r = raw_at<Acc, TIn>(reinterpret_cast<const unsigned char*>(rp), (c + 20) * L);
        v0[c] = p + r;
        v1[c] = p - r;
      }
      paley_store<Acc, 20>(v0, wp, pstride, std::make_integer_sequence<int, 20>{});
      paley_store<Acc, 20>(v1, wp + 20 * pstride, pstride, std::make_integer_sequence<int, 20>{});
    }
  }
}

// bits [5, TOP) in register-blocked passes from the top down (the schedule of the run-time loop below, unrolled at
// compile time when the shape is a template argument)
constexpr int pass_bits(int top) {
  const int hi = top - 5;
  int r = hi % 5;
  if (r == 0) r = 5;
  if (hi > 5 && hi < 10) r = (hi + 1) / 2;  // 6..9 -> two balanced passes
  return r;
}
template <typename Acc, typename TIn, int NC, int TOP, bool FROM_RAW>
__device__ __forceinline__ void passes_ct(const HadArgs& a, const unsigned char* raw, Acc* sm, int64_t row0, int nr) {
  if constexpr (TOP > 5) {
    constexpr int R = pass_bits(TOP);
    mid_bits<Acc, TIn, R, FROM_RAW, NC, TOP - R>(a, raw, sm, nr, TOP - R);
    __syncthreads();
    passes_ct<Acc, TIn, NC, TOP - R, false>(a, raw, sm, row0, nr);
  } else {
    low_bits_out<Acc, TIn, FROM_RAW, NC>(a, raw, sm, row0, nr);
  }
}

// MC > 0: n = KT * 2^MC is a compile-time shape (the common hidden sizes); MC == 0: any m >= 5 at run time
template <typename Acc, typename TIn, int KT, int MC>
__global__ void __launch_bounds__(HAD_THREADS, sizeof(Acc) == 4 ? 4 : 2) hadamard_tile_kernel(const __grid_constant__ HadArgs a) {
  extern __shared__ __align__(16) unsigned char had_smem[];
  unsigned char* raw = had_smem;
  Acc* sm = reinterpret_cast<Acc*>(had_smem + a.raw_bytes);
  uint4* sgn = reinterpret_cast<uint4*>(had_smem + a.sgn_off);  // one row of XOR masks (sign bit set <=> sign < 0), raw layout
  constexpr int EPV = 16 / (int)sizeof(TIn);                    // elements per 16-byte vector
  const int nvec_row = (MC > 0 ? (KT << MC) : a.n) / EPV;
  if (a.signs != nullptr) {
    for (int i = threadIdx.x; i < nvec_row; i += HAD_THREADS) {
      uint32_t w[4];
      if constexpr (EPV == 8) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          w[q] = (a.signs[i * 8 + 2 * q] < 0.0f ? 0x8000u : 0u) | (a.signs[i * 8 + 2 * q + 1] < 0.0f ? 0x80000000u : 0u);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = a.signs[i * 4 + q] < 0.0f ? 0x80000000u : 0u;
      }
      sgn[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  __syncthreads();
  int v0 = threadIdx.x;   // vector index within a row of the first vector this thread stages
  while (v0 >= nvec_row) v0 -= nvec_row;
  int vstep = HAD_THREADS;
  while (vstep >= nvec_row) vstep -= nvec_row;   // HAD_THREADS mod nvec_row
  const int64_t stride_rows = (int64_t)gridDim.x * a.rpc;
  // The raw tile of the NEXT iteration is fetched with cp.async as soon as the last reader of the current one is done
  // (the H_K stage, or the first radix-2 pass when K == 1): the HBM latency hides under the remaining passes.
  auto fetch = [&](int64_t r0) {
    if (r0 >= a.rows) return;
    const int cnt = (int)min((int64_t)a.rpc, a.rows - r0) * nvec_row;
    const uint4* src = reinterpret_cast<const uint4*>(static_cast<const TIn*>(a.x) + r0 * (int64_t)a.n);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(raw);
    for (int i = threadIdx.x; i < cnt; i += HAD_THREADS)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + i) : "memory");
  };
  fetch((int64_t)blockIdx.x * a.rpc);
  for (int64_t row0 = (int64_t)blockIdx.x * a.rpc; row0 < a.rows; row0 += stride_rows) {
    const int nr = (int)min((int64_t)a.rpc, a.rows - row0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (a.signs != nullptr) {  // the signs go into the sign bits in place; every thread touches only the vectors it fetched
      uint4* dst = reinterpret_cast<uint4*>(raw);
      const int nvec = nr * nvec_row;
      int v = v0;
      for (int i = threadIdx.x; i < nvec; i += HAD_THREADS) {
        uint4 t = dst[i];
        const uint4 mk = sgn[v];
        t.x ^= mk.x; t.y ^= mk.y; t.z ^= mk.z; t.w ^= mk.w;
        dst[i] = t;
        v += vstep;
        if (v >= nvec_row) v -= nvec_row;
      }
    }
    __syncthreads();
    if constexpr (MC > 0) {
      if constexpr (KT > 1) {
        hadk_first<Acc, TIn, KT, MC>(a, raw, sm, nr);
        __syncthreads();
        fetch(row0 + stride_rows);
        passes_ct<Acc, TIn, (KT << MC), MC, false>(a, raw, sm, row0, nr);
      } else if constexpr (MC > 5) {
        constexpr int R = pass_bits(MC);
        mid_bits<Acc, TIn, R, true, (KT << MC), MC - R>(a, raw, sm, nr, MC - R);
        __syncthreads();
        fetch(row0 + stride_rows);
        passes_ct<Acc, TIn, (KT << MC), MC - R, false>(a, raw, sm, row0, nr);
      } else {
        low_bits_out<Acc, TIn, true, (KT << MC)>(a, raw, sm, row0, nr);
        __syncthreads();
        fetch(row0 + stride_rows);
      }
    } else {
      bool from_raw = true;
      if constexpr (KT > 1) {
        hadk_first<Acc, TIn, KT>(a, raw, sm, nr);
        __syncthreads();
        fetch(row0 + stride_rows);
        from_raw = false;
      }
      int top = a.m;  // bits [5, top) are still to do
      while (top > 5) {
        const int r = pass_bits(top);
        const int b0 = top - r;
#define LCB_HAD_PASS(RR)                                                        \
  case RR:                                                                      \
    if (from_raw) mid_bits<Acc, TIn, RR, true>(a, raw, sm, nr, b0);             \
    else mid_bits<Acc, TIn, RR, false>(a, raw, sm, nr, b0);                     \
    break;
        switch (r) {
          LCB_HAD_PASS(1) LCB_HAD_PASS(2) LCB_HAD_PASS(3) LCB_HAD_PASS(4) LCB_HAD_PASS(5)
        }
#undef LCB_HAD_PASS
        __syncthreads();
        if (from_raw) fetch(row0 + stride_rows);
        from_raw = false;
        top = b0;
      }
      if (from_raw) {
        low_bits_out<Acc, TIn, true>(a, raw, sm, row0, nr);
        __syncthreads();
        fetch(row0 + stride_rows);
      } else {
        low_bits_out<Acc, TIn, false>(a, raw, sm, row0, nr);
      }
    }
    __syncthreads();  // the next tile overwrites sm
  }
}

template <typename Acc, typename TIn, int KT, int MC = 0>
int launch_tile(const HadArgs& a, int grid, size_t smem, cudaStream_t st) {
  LCB_CUDA(cudaFuncSetAttribute(hadamard_tile_kernel<Acc, TIn, KT, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hadamard_tile_kernel<Acc, TIn, KT, MC><<<grid, HAD_THREADS, smem, st>>>(a);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

template <typename Acc, typename TIn>
int launch_tile_k(const HadArgs& a, int grid, size_t smem, cudaStream_t st) {
  // compile-time shapes: the hidden / head / intermediate sizes of the reference's model zoo (3072 = 12 * 2^8,
  // 2560 = 40 * 2^6, 128, 1024, 2048, 4096, 8192); everything else runs the same kernel with run-time strides
  switch (a.K) {
    case 1:
      switch (a.m) {
        case 7: return launch_tile<Acc, TIn, 1, 7>(a, grid, smem, st);
        case 10: return launch_tile<Acc, TIn, 1, 10>(a, grid, smem, st);
        case 11: return launch_tile<Acc, TIn, 1, 11>(a, grid, smem, st);
        case 12: return launch_tile<Acc, TIn, 1, 12>(a, grid, smem, st);
        case 13: return launch_tile<Acc, TIn, 1, 13>(a, grid, smem, st);
        default: return launch_tile<Acc, TIn, 1>(a, grid, smem, st);
      }
    case 12:
      if (a.m == 8) return launch_tile<Acc, TIn, 12, 8>(a, grid, smem, st);
      return launch_tile<Acc, TIn, 12>(a, grid, smem, st);
    default:
      if (a.m == 6) return launch_tile<Acc, TIn, 40, 6>(a, grid, smem, st);
      return launch_tile<Acc, TIn, 40>(a, grid, smem, st);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Register kernel (bf16 in, fp32 accumulation, the row widths of the model zoo).  The tiled kernel above moves every
// element through shared memory once per pass (24 B / element for n = 3072): at 2.7 TB/s of HBM traffic the shared-memory
// pipe is half busy and the barrier-separated passes cannot hide it.  Here a row is held in the REGISTERS of TPR = 32 or 64
// threads and crosses shared memory exactly twice (8 B as fp32 between the two register phases, 2 x the output size on
// the way out):
//   element i = j (n / V) + 8 t + e      j < V = KT 2^X   vector index of thread t (the H_K index and the top X bits)
//                                        t < TPR = 2^LT   thread of the row group
//                                        e < 8            position inside the 16-byte vector
//   phase A  thread t loads its V vectors (16-byte loads, a warp reads 512 contiguous bytes per j), flips the signs, and
//            runs H_K and the radix-2 stages of the X + 3 register-resident bits on its EPT = 8 V values;
//   A -> B   S[t][c] (fp32, row stride EPT + 4: 128-bit stores, conflict-free), thread u reads back the CB = EPT / TPR
//            values c in [u CB, (u + 1) CB) of EVERY t;
//   phase B  the LT stages across t, scaling, conversion;
//   B -> C   output row image O[j][t][e] (8 elements of padding per j: conflict-free scalar stores), read back as 16-byte
//            vectors in the layout of phase A and stored with the access pattern of the loads.
// A group of TPR threads owns a row from load to store: groups only synchronise internally (__syncwarp for TPR = 32, a
// 64-thread named barrier for TPR = 64), never the CTA.
// A shape: EV elements per vector (8 = 16-byte loads, 4 = 8-byte loads), V vectors per thread, TPR = 2^LT threads per row.
//   STRIDED == false: phase-B thread u owns the CB = EPT / TPR columns [u CB, (u + 1) CB)           (EPT % TPR == 0)
//   STRIDED == true : phase-B thread u owns the columns u, u + TPR, ... < EPT (NK = ceil(EPT / TPR) rounds, the last ragged)
template <int KT_, int X, int LT_, int EV_ = 8, bool STRIDED_ = false>
struct RegShape {
  static constexpr int KT = KT_, LT = LT_, EV = EV_;
  static constexpr bool STRIDED = STRIDED_;
  // KT == 40: H_40 = H_2 (x) H_20; the 20 values of H_20 are a thread's vectors, the H_2 bit joins the thread index
  static constexpr int KB = KT == 40 ? 20 : KT;
  static constexpr int V = KB << X, EPT = EV * V, TPR = 1 << LT, N = EPT * TPR, SA = EPT + 4;
  static constexpr int NK = (EPT + TPR - 1) / TPR;      // columns per phase-B thread (rounded up when STRIDED)
  static constexpr int OJ = TPR * EV + 8;               // elements per j of the output image
  static constexpr int GROUP_BYTES = TPR * SA * 4 > V * OJ * 4 ? TPR * SA * 4 : V * OJ * 4;   // S and the output image share it
  static_assert(STRIDED || EPT % TPR == 0, "blocked columns: every phase-B thread holds whole columns");
  static_assert(KT == 1 || X == 0, "H_K needs all K values of a column in one thread");
  static_assert(LT >= 3 && LT <= 6, "row groups of 8 .. 64 threads");
  static_assert(KT != 40 || (LT == 5 && EV == 4) || (LT == 4 && EV == 8), "K = 40: n = 2560, thread = (H_2 bit, upper bits of b)");
  // position (in vectors of EV elements) of vector j of thread t inside the row
  __device__ static __forceinline__ int vec(int j, int t) {
    if constexpr (KT == 40) return (t >> (LT - 1)) * (20 << (LT - 1)) + (j << (LT - 1)) + (t & ((1 << (LT - 1)) - 1));
    else return j * TPR + t;
  }
  // column of phase-B thread u, round k
  __device__ static __forceinline__ int col(int u, int k) { return STRIDED ? u + TPR * k : u * NK + k; }
};

// Packed fp32 pairs (sm_100 add / sub / mul .f32x2 -> FADD2 / FMUL2): one instruction for two lanes' worth of adds.  The
// kernel is bound by instruction issue (13 adds per element for n = 3072), so every stage whose two operands are
// neighbours in the register array runs on pairs.
struct F2 {
  uint64_t v;
  __device__ __forceinline__ F2() {}
  __device__ __forceinline__ F2(int) : v(0ull) {}   // zero (the accumulator start of paley_dot)
  __device__ __forceinline__ F2(float lo, float hi) { asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); }
  __device__ __forceinline__ void unpack(float& lo, float& hi) const { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
  friend __device__ __forceinline__ F2 operator+(F2 a, F2 b) {
    F2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
  }
  friend __device__ __forceinline__ F2 operator-(F2 a, F2 b) {
    F2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
  }
  friend __device__ __forceinline__ F2 operator*(F2 a, F2 b) {
    F2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
  }
  __device__ __forceinline__ F2 operator-() const { return F2(0) - *this; }
};

template <int KB, int... Rs>
__device__ __forceinline__ void paley_out(const F2 (&v)[KB], F2 (&o)[KB], std::integer_sequence<int, Rs...>) {
  F2 cmb[KB / 4][8];
  paley_groups<F2, KB>(v, cmb);
  ((o[Rs] = paley_dot<F2, KB, Rs>(cmb)), ...);
}

template <int N, int H0, int H1>  // radix-2 stages with strides H0, 2 H0, ... < H1 over a register array (N even)
__device__ __forceinline__ void bfly_range(float (&v)[N]) {
#pragma unroll
  for (int h = H0; h < H1; h <<= 1) {
    if (h == 1) {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        const float p = v[j], q = v[j + 1];
        v[j] = p + q;
        v[j + 1] = p - q;
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        if ((j & h) == 0) {
          const F2 p(v[j], v[j + 1]), q(v[j | h], v[(j | h) + 1]);
          (p + q).unpack(v[j], v[j + 1]);
          (p - q).unpack(v[j | h], v[(j | h) + 1]);
        }
      }
    }
  }
}

template <int LT>
__device__ __forceinline__ void group_sync(int grp) {
  if constexpr (LT == 5) __syncwarp();
  else if constexpr (LT < 5)   // several row groups per warp, independent trip counts: only this group's lanes
    __syncwarp((LT == 4 ? 0xffffu : 0xffu) << ((1 << LT) * (grp & ((32 >> LT) - 1))));
  else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(1 << LT) : "memory");
}

template <typename TOut, typename S, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) hadamard_reg_kernel(const __grid_constant__ HadArgs a) {
  constexpr int KT = S::KT, LT = S::LT, EV = S::EV, V = S::V, EPT = S::EPT, TPR = S::TPR, N = S::N, NK = S::NK, SA = S::SA, OJ = S::OJ;
  constexpr int GROUPS = THREADS >> LT;
  constexpr int WPV = EV / 2;   // 32-bit words per vector
  extern __shared__ __align__(16) unsigned char had_smem[];
  uint32_t* sgn = reinterpret_cast<uint32_t*>(had_smem);   // [N / 2] XOR masks, two elements per word, row order
  const int tid = threadIdx.x, grp = tid >> LT, t = tid & (TPR - 1);
  float* Sg = reinterpret_cast<float*>(had_smem + N * 2 + grp * S::GROUP_BYTES);
  TOut* Og = reinterpret_cast<TOut*>(Sg);
  const bool flip = a.signs != nullptr;
  if (flip) {
    for (int i = tid; i < N / 2; i += THREADS)
      sgn[i] = (a.signs[2 * i] < 0.0f ? 0x8000u : 0u) | (a.signs[2 * i + 1] < 0.0f ? 0x80000000u : 0u);
  }
  __syncthreads();
  const float rcp = (float)a.rcp;
  int oofs[NK];      // where this thread's columns go in the output image
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int c = S::col(t, k);
    oofs[k] = (c / EV) * OJ + (c % EV);
  }
  const uint32_t* xin = static_cast<const uint32_t*>(a.x);
  for (int64_t row = (int64_t)blockIdx.x * GROUPS + grp; row < a.rows; row += (int64_t)gridDim.x * GROUPS) {
    float r[EPT];
    {  // ---- phase A: load, sign flip, unpack
      const uint32_t* src = xin + row * (int64_t)(N / 2);
      uint32_t raw[V][WPV];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if constexpr (EV == 8) {
          const uint4 w = __ldcs(reinterpret_cast<const uint4*>(src) + S::vec(j, t));
          raw[j][0] = w.x; raw[j][1] = w.y; raw[j][2] = w.z; raw[j][3] = w.w;
        } else {
          const uint2 w = __ldcs(reinterpret_cast<const uint2*>(src) + S::vec(j, t));
          raw[j][0] = w.x; raw[j][1] = w.y;
        }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        uint32_t mk[WPV];
        if (flip) {
          if constexpr (EV == 8) {
            const uint4 m4 = reinterpret_cast<const uint4*>(sgn)[S::vec(j, t)];
            mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
          } else {
            const uint2 m2 = reinterpret_cast<const uint2*>(sgn)[S::vec(j, t)];
            mk[0] = m2.x; mk[1] = m2.y;
          }
        }
#pragma unroll
        for (int q = 0; q < WPV; ++q) {
          const uint32_t w = flip ? (raw[j][q] ^ mk[q]) : raw[j][q];
          r[EV * j + 2 * q] = __uint_as_float(w << 16);
          r[EV * j + 2 * q + 1] = __uint_as_float(w & 0xffff0000u);
        }
      }
    }
    if constexpr (KT == 1) {
      bfly_range<EPT, 1, EPT>(r);
    } else {
      bfly_range<EPT, 1, EV>(r);
      static_assert(KT == 1 || KT == 12 || KT == 40, "register kernel: H_12 and H_40 = H_2 (x) H_20 only");
      constexpr int KB = S::KB;
#pragma unroll
      for (int e = 0; e < EV; e += 2) {   // two neighbouring columns per packed H_K
        F2 v[KB], o[KB];
#pragma unroll
        for (int c = 0; c < KB; ++c) v[c] = F2(r[EV * c + e], r[EV * c + e + 1]);
        paley_out<KB>(v, o, std::make_integer_sequence<int, KB>{});
#pragma unroll
        for (int c = 0; c < KB; ++c) o[c].unpack(r[EV * c + e], r[EV * c + e + 1]);
      }
    }
    // ---- A -> B
#pragma unroll
    for (int q = 0; q < EPT / 4; ++q)
      *reinterpret_cast<float4*>(Sg + t * SA + 4 * q) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    group_sync<LT>(grp);
    float vb[NK][TPR];
#pragma unroll
    for (int tt = 0; tt < TPR; ++tt) {
      if constexpr (S::STRIDED) {
#pragma unroll
        for (int k = 0; k < NK; ++k)
          if ((k + 1) * TPR <= EPT || t + TPR * k < EPT) vb[k][tt] = Sg[tt * SA + t + TPR * k];
      } else {
        const float* p = Sg + tt * SA + t * NK;
        if constexpr (NK % 4 == 0) {
#pragma unroll
          for (int k = 0; k < NK; k += 4) {
            const float4 f = *reinterpret_cast<const float4*>(p + k);
            vb[k][tt] = f.x; vb[k + 1][tt] = f.y; vb[k + 2][tt] = f.z; vb[k + 3][tt] = f.w;
          }
        } else if constexpr (NK % 2 == 0) {
#pragma unroll
          for (int k = 0; k < NK; k += 2) {
            const float2 f = *reinterpret_cast<const float2*>(p + k);
            vb[k][tt] = f.x; vb[k + 1][tt] = f.y;
          }
        } else {
#pragma unroll
          for (int k = 0; k < NK; ++k) vb[k][tt] = p[k];
        }
      }
    }
    group_sync<LT>(grp);   // every read of S is done: the output image may overwrite it
    // ---- phase B + B -> C
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      if ((k + 1) * TPR <= EPT || !S::STRIDED || t + TPR * k < EPT) {
        bfly_range<TPR, 1, TPR>(vb[k]);
        TOut* o = Og + oofs[k];
        const F2 rcp2(rcp, rcp);
#pragma unroll
        for (int tt = 0; tt < TPR; tt += 2) {
          float y0, y1;
          (F2(vb[k][tt], vb[k][tt + 1]) * rcp2).unpack(y0, y1);
          if constexpr (sizeof(TOut) == 2) {
            o[tt * EV] = __float2bfloat16_rn(y0);
            o[(tt + 1) * EV] = __float2bfloat16_rn(y1);
          } else {
            o[tt * EV] = y0;
            o[(tt + 1) * EV] = y1;
          }
        }
      }
    }
    group_sync<LT>(grp);
    {
      constexpr int OB = EV * (int)sizeof(TOut);   // bytes per output vector: 8, 16 or 32
      unsigned char* dst = static_cast<unsigned char*>(a.y) + row * (int64_t)N * (int64_t)sizeof(TOut);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(Og + j * OJ + t * EV);
        unsigned char* d = dst + (int64_t)S::vec(j, t) * OB;
        if constexpr (OB == 8) {
          __stcs(reinterpret_cast<uint2*>(d), *reinterpret_cast<const uint2*>(p));
        } else {
#pragma unroll
          for (int q = 0; q < OB / 16; ++q) __stcs(reinterpret_cast<uint4*>(d) + q, reinterpret_cast<const uint4*>(p)[q]);
        }
      }
    }
    group_sync<LT>(grp);   // the image is read: the next row may overwrite S
  }
}

template <typename TOut, typename S, int THREADS, int MINB>
int launch_reg(const HadArgs& a, cudaStream_t st) {
  constexpr int GROUPS = THREADS >> S::LT;
  const size_t smem = (size_t)S::N * 2 + (size_t)GROUPS * S::GROUP_BYTES;
  auto kern = hadamard_reg_kernel<TOut, S, THREADS, MINB>;
  LCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(MINB, (227 * 1024) / (smem + 1024)));
  const int grid = (int)std::min<int64_t>(ceil_div(a.rows, GROUPS), (int64_t)sm_count() * per_sm);
  kern<<<grid, THREADS, smem, st>>>(a);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

// n -> register-kernel configuration; returns -1 when the width has none
template <typename TOut>
int launch_reg_n(const HadArgs& a, cudaStream_t st) {
  static int var = -1;   // LCB_HAD_VAR=1: alternative configurations (A/B runs)
  if (var < 0) {
    const char* e = getenv("LCB_HAD_VAR");
    var = (e && e[0] == '1') ? 1 : 0;
  }
  if (a.K == 12 && a.m == 8) return launch_reg<TOut, RegShape<12, 0, 5>, 128, 3>(a, st);   // 3072 (64 threads x 48 values per row: 3.4 TB/s)
  if (a.K == 40 && a.m == 6)                                                                // 2560
    return var ? launch_reg<TOut, RegShape<40, 0, 4>, 128, 2>(a, st) : launch_reg<TOut, RegShape<40, 0, 5, 4, true>, 128, 3>(a, st);
  if (a.K == 1) {
    switch (a.m) {
      case 6: return launch_reg<TOut, RegShape<1, 0, 3>, 256, 4>(a, st);    // 64   (head dims: 8 threads per row)
      case 7: return launch_reg<TOut, RegShape<1, 1, 3>, 256, 4>(a, st);    // 128
      case 8: return launch_reg<TOut, RegShape<1, 1, 4>, 256, 4>(a, st);    // 256
      case 9: return launch_reg<TOut, RegShape<1, 2, 4>, 256, 4>(a, st);    // 512
      case 10: return launch_reg<TOut, RegShape<1, 2, 5>, 256, 4>(a, st);   // 1024
      case 11: return launch_reg<TOut, RegShape<1, 3, 5>, 256, 2>(a, st);   // 2048
      case 12: return launch_reg<TOut, RegShape<1, 3, 6>, 256, 2>(a, st);   // 4096
      case 13:                                                              // 8192: 6 rows in flight per SM (a row's S is 33 KB)
        return var ? launch_reg<TOut, RegShape<1, 4, 6>, 128, 2>(a, st) : launch_reg<TOut, RegShape<1, 4, 6>, 384, 1>(a, st);
      default: break;
    }
  }
  return -1;
}

static int reg_mode() {   // LCB_HAD_REG=0: tiled kernel for every shape (A/B runs)
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("LCB_HAD_REG");
    m = (e && e[0] == '0') ? 0 : 1;
  }
  return m;
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

// The tiled kernel folds the reference's had12 / had20 sign patterns at compile time; a caller-supplied table that
// differs (the transposed one of matmul_hadUt, or any other H_K) takes the generic kernel.
static bool transpose_table(const uint64_t* bits, int K) {
  if (K == 1) return false;
  const int kb = K == 40 ? 20 : K;
  for (int r = 0; r < K; ++r) {
    unsigned long long want = paley1_row_mask(kb, r % kb);
    if (K == 40) want = (r < 20) ? (want | (want << 20)) : (want | ((~want & 0xfffffull) << 20));
    if (bits[r] != want) return true;
  }
  return false;
}

extern "C" int lcb_hadamard_rows(const void* x, int dtype_in, void* y, int dtype_out, int64_t rows, int64_t n,
                                 const float* signs, const uint64_t* hadk_bits, int K, double divisor, int acc64,
                                 void* stream) {
  LCB_REQUIRE(x && y && rows >= 0 && n > 0 && K >= 1 && K <= HAD_BIGK && divisor != 0.0, "lcb_hadamard_rows: bad arguments");
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
  HadArgs a{};
  a.x = x; a.y = y; a.signs = signs; a.rows = rows; a.n = (int)n; a.m = m; a.K = K;
  a.dt_in = dtype_in; a.dt_out = dtype_out; a.divisor = divisor; a.rcp = 1.0 / divisor;
  for (int i = 0; i < K && K > 1 && K <= HAD_MAXK; ++i) a.hk[i] = hadk_bits[i];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t esz_in = dtype_in == LCB_BF16 ? 2 : (dtype_in == LCB_F32 ? 4 : 8);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (!acc64 && aligned && dtype_in == LCB_BF16 && dtype_out != LCB_F64 && (K == 1 || K == 12 || K == 40) && reg_mode() != 0 &&
      !transpose_table(hadk_bits, K)) {
    // register kernel: bf16 rows of the common widths, fp32 accumulation
    const int rc = dtype_out == LCB_BF16 ? launch_reg_n<__nv_bfloat16>(a, st) : launch_reg_n<float>(a, st);
    if (rc >= 0) return rc;
  }
  if (m >= 5 && aligned && dtype_in != LCB_F64 && (K == 1 || K == 12 || K == 40) && !transpose_table(hadk_bits, K)) {
    // tiled kernel: raw tile + padded work buffer + sign words; aim at two resident CTAs per SM
    const int64_t budget = 100 * 1024;
    const int64_t per_row = (int64_t)n * esz_in + (int64_t)(n + (n >> 5)) * esz;
    int64_t rpc = std::max<int64_t>(1, budget / per_row);
    rpc = std::min<int64_t>(rpc, std::max<int64_t>(1, 8192 / n));
    const int64_t tile = rpc * n;
    const size_t raw_bytes = (size_t)((tile * esz_in + 15) / 16 * 16);
    const size_t work_bytes = (size_t)((tile + (tile >> 5) + 16) * esz + 15) / 16 * 16;
    const size_t smem = raw_bytes + work_bytes + (size_t)(n * esz_in + 15) / 16 * 16;   // + one row of sign masks
    if (smem <= 220 * 1024) {
      a.rpc = (int)rpc;
      a.raw_bytes = (int)raw_bytes;
      a.sgn_off = (int)(raw_bytes + work_bytes);
      const int64_t tiles = ceil_div(rows, rpc);
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024)));
      const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count() * per_sm);
      if (dtype_in == LCB_BF16)
        return acc64 ? launch_tile_k<double, __nv_bfloat16>(a, grid, smem, st) : launch_tile_k<float, __nv_bfloat16>(a, grid, smem, st);
      return acc64 ? launch_tile_k<double, float>(a, grid, smem, st) : launch_tile_k<float, float>(a, grid, smem, st);
    }
  }
  // generic kernel: any K * 2^m (m < 4, unaligned pointers, rows too long for raw + work tiles)
  const int64_t max_elems = (int64_t)(200 * 1024 / esz) * 16 / 17;
  if (n > max_elems) {
    set_error("lcb_hadamard_rows: n = %lld exceeds the shared-memory tile (%lld)", (long long)n, (long long)max_elems);
    return LCB_ERR_UNSUPPORTED;
  }
  a.rpc = (int)std::max<int64_t>(1, 4096 / n);
  const int64_t tile = (int64_t)a.rpc * n;
  const size_t smem = (size_t)(tile + (tile >> 4) + 16) * esz;
  const int64_t tiles = ceil_div(rows, a.rpc);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count() * per_sm);
  if (K > HAD_MAXK) {   // HAD_BIGW words per row
    HadArgsBig ab{};
    ab.b = a;
    for (int i = 0; i < K * HAD_BIGW; ++i) ab.hkx[i] = hadk_bits[i];
    return acc64 ? launch_had_big<double>(ab, grid, smem, st) : launch_had_big<float>(ab, grid, smem, st);
  }
  return acc64 ? launch_had_k<double>(a, grid, smem, st) : launch_had_k<float>(a, grid, smem, st);
}

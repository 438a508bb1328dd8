// fp32 SIMT GEMM (128x128x16 tiles, 8x8 register blocking, register-prefetch double buffering)
// for the solver stages that must keep full fp32 accuracy (ref: gptq/core.py:265 lazy update,
// gptaq/core.py:272 P, Cholesky trailing updates).  v1 of these contractions: exact-fp32 FFMA.
#include "linalg.cuh"

namespace lcb {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS = BM + 4;

__device__ __forceinline__ float4 ld4_guard(const float* base, int64_t ld, int r, int c, int R_, int C_, bool fast) {
  // element (r, c..c+3) of a row-major matrix with R_ x C_ valid extent
  if (fast) return *reinterpret_cast<const float4*>(base + (int64_t)r * ld + c);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < R_) {
    const float* p = base + (int64_t)r * ld + c;
    if (c + 0 < C_) v.x = p[0];
    if (c + 1 < C_) v.y = p[1];
    if (c + 2 < C_) v.z = p[2];
    if (c + 3 < C_) v.w = p[3];
  }
  return v;
}

__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];
  const int bm = blockIdx.y, bn = blockIdx.x;
  const int m0 = bm * BM, n0 = bn * BN;
  if ((g.tri & GEMM_LOWER_OUT) && n0 > m0 + BM - 1) return;
  const float* A = g.A + (int64_t)blockIdx.z * g.strideA;
  const float* B = g.B + (int64_t)blockIdx.z * g.strideB;
  float* C = g.C + (int64_t)blockIdx.z * g.strideC;
  int kbeg = 0, kend = g.K;
  if (g.tri & GEMM_A_LOWER) kend = min(kend, m0 + BM);
  if (g.tri & GEMM_B_LOWER) kbeg = (n0 / BK) * BK;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;

  const bool a_al = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (g.lda % 4 == 0);
  const bool b_al = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (g.ldb % 4 == 0);
  const bool a_full = a_al && (m0 + BM <= g.M);
  const bool bt_full = b_al && (n0 + BN <= g.N);  // transB: rows of B are n
  const bool bn_full = b_al && (n0 + BN <= g.N);  // NN: columns of B are n

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto fetch = [&](int k0) {
    const bool kfull = (k0 + BK <= g.K);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = t + 256 * i;
      {  // A tile: [BM x BK], row = idx / 4, k-quad = idx % 4
        const int row = idx >> 2, kq = idx & 3;
        ra[i] = ld4_guard(A, g.lda, m0 + row, k0 + kq * 4, g.M, g.K, a_full && kfull);
      }
      if (g.transB) {  // B is [N, K]
        const int row = idx >> 2, kq = idx & 3;
        rb[i] = ld4_guard(B, g.ldb, n0 + row, k0 + kq * 4, g.N, g.K, bt_full && kfull);
      } else {  // B is [K, N]: krow = idx / 32, n-quad = idx % 32
        const int krow = idx >> 5, nq = idx & 31;
        rb[i] = ld4_guard(B, g.ldb, k0 + krow, n0 + nq * 4, g.K, g.N, bn_full && kfull);
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = t + 256 * i;
      {
        const int row = idx >> 2, kq = idx & 3;
        As[kq * 4 + 0][row] = ra[i].x; As[kq * 4 + 1][row] = ra[i].y;
        As[kq * 4 + 2][row] = ra[i].z; As[kq * 4 + 3][row] = ra[i].w;
      }
      if (g.transB) {
        const int row = idx >> 2, kq = idx & 3;
        Bs[kq * 4 + 0][row] = rb[i].x; Bs[kq * 4 + 1][row] = rb[i].y;
        Bs[kq * 4 + 2][row] = rb[i].z; Bs[kq * 4 + 3][row] = rb[i].w;
      } else {
        const int krow = idx >> 5, nq = idx & 31;
        *reinterpret_cast<float4*>(&Bs[krow][nq * 4]) = rb[i];
      }
    }
  };

  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    __syncthreads();  // previous tile fully consumed
    stash();
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);  // global loads overlap the FMAs below
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  const bool c_al = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (g.ldc % 4 == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= g.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = n0 + (h == 0 ? tx * 4 : 64 + tx * 4);
      float* cp = C + (int64_t)r * g.ldc + c;
      if (c_al && c + 3 < g.N) {
        float4 o;
        if (g.beta != 0.f) {
          const float4 old = *reinterpret_cast<const float4*>(cp);
          o.x = g.alpha * acc[i][h * 4 + 0] + g.beta * old.x; o.y = g.alpha * acc[i][h * 4 + 1] + g.beta * old.y;
          o.z = g.alpha * acc[i][h * 4 + 2] + g.beta * old.z; o.w = g.alpha * acc[i][h * 4 + 3] + g.beta * old.w;
        } else {
          o.x = g.alpha * acc[i][h * 4 + 0]; o.y = g.alpha * acc[i][h * 4 + 1];
          o.z = g.alpha * acc[i][h * 4 + 2]; o.w = g.alpha * acc[i][h * 4 + 3];
        }
        *reinterpret_cast<float4*>(cp) = o;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c + j < g.N) {
            float v = g.alpha * acc[i][h * 4 + j];
            if (g.beta != 0.f) v += g.beta * cp[j];
            cp[j] = v;
          }
        }
      }
    }
  }
}

}  // namespace

int sgemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return LCB_OK;
  dim3 grid((unsigned)ceil_div(g.N, BN), (unsigned)ceil_div(g.M, BM), (unsigned)g.batch);
  sgemm_kernel<<<grid, 256, 0, st>>>(g);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

}  // namespace lcb

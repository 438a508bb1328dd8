// fp32 dense linear algebra building blocks used by the layer solvers (solver.cu) and the
// Cholesky-inverse (chol.cu).  Row-major everywhere.
#pragma once

#include "common.cuh"

namespace lcb {

// C[M,N] = alpha * A[M,K] * op(B) + beta * C       (row-major, leading dimensions in elements)
//   transB == 0: B is [K,N]           transB == 1: B is [N,K]  (C = A * B^T)
// tri (tile skipping for triangular operands / results, all optional):
//   GEMM_LOWER_OUT   : only tiles of C that intersect the lower triangle (incl. diagonal) are
//                      computed (SYRK-style update of a symmetric / lower-triangular matrix)
//   GEMM_A_LOWER     : A is lower triangular (k <= m contributes): k-loop is cut per row tile
//   GEMM_B_LOWER     : B (transB == 0, [K,N]) is lower triangular (k >= n contributes)
// batch: grid.z strides (in elements) for A, B, C.
enum { GEMM_LOWER_OUT = 1, GEMM_A_LOWER = 2, GEMM_B_LOWER = 4 };

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int64_t lda, ldb, ldc;
  float alpha, beta;
  int transB;
  int tri;
  int batch;
  int64_t strideA, strideB, strideC;
};

int sgemm(const GemmArgs& g, cudaStream_t st);

inline GemmArgs gemm_args(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                          int K, float alpha, float beta, int transB, int tri = 0) {
  GemmArgs g{};
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.alpha = alpha; g.beta = beta; g.transB = transB; g.tri = tri;
  g.batch = 1; g.strideA = g.strideB = g.strideC = 0;
  return g;
}

// ---------------------------------------------------------------------------------------------
// fp32-accurate tensor-core GEMM (tgemm.cu): C[M,N] (+)= alpha * A[M,Kd] * B[N,Kd]^T on tcgen05 with
// the 3xTF32 split.  Operands are given as pre-split hi / lo planes (split_tf32), K-major.
//   TG_STORE      C = alpha*A*B^T (default: C += alpha*A*B^T, reduce-add at L2)
//   TG_LOWER_OUT  only tiles of C touching the lower triangle are produced
//   TG_A_LOWER    A[m][k] == 0 for k > m  (k-loop cut per row tile)
//   TG_A_UPPER    A[m][k] == 0 for k < m
// Requirements: all bases 16 B aligned, leading dimensions multiples of 4.
//   TG_PDL        launch with programmatic stream serialization: the kernel's prologue (barriers, TMEM allocation, tensor-map
//                 prefetch) overlaps the tail of the preceding kernel of the stream; it executes griddepcontrol.wait before it
//                 touches global memory
enum { TG_STORE = 1, TG_LOWER_OUT = 2, TG_A_LOWER = 4, TG_A_UPPER = 8, TG_PDL = 16 };
// kchain > 0: the reduction is cut into accumulation chains of at most kchain columns, each reduce-added
// into C separately.  The tensor core's fp32 accumulator truncates, so the error of one chain grows
// linearly with its length (3 * kchain / 8 MMAs); 0 = one chain per tile.
int tgemm_nt(const float* Ah, const float* Al, int64_t lda, const float* Bh, const float* Bl, int64_t ldb, float* C,
             int64_t ldc, int M, int N, int Kd, float alpha, int flags, cudaStream_t st, int kchain = 0);
// hi = tf32(x), lo = x - hi; transpose != 0: out[c][r] = split(in[r][c])
int split_tf32(const float* in, int64_t ld_in, int rows, int cols, float* hi, float* lo, int64_t ld_out, int transpose,
               cudaStream_t st);
inline bool tg_ok(const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0; }

// solver GEMM precision: 0 = exact fp32 FFMA (sgemm), 1 = 3xTF32 on tcgen05 (default)
int gemm_mode();

// Two library-owned side streams + events per host thread and device: look-ahead of the latency-bound chains
// (Cholesky panels, GPTQ block steps) -- the update the next step needs stays on the caller's stream, the rest
// runs underneath the following steps.  nullptr (and lcb_last_error set) when they cannot be created.
struct SideStreams {
  cudaStream_t s[3];  // 0: B(q) panel updates, 1: S_B(J) strip updates, 2: triangular inverse of finished blocks
  cudaEvent_t evP, evB[2], evS[2], evT;
  // the GPTQ block chain itself (quantiser + next-block update) runs on a stream of the GREATEST priority, fenced against the
  // caller's stream by evIn / evOut: a chain GEMM that becomes ready while a look-ahead GEMM of a side stream (priority 0,
  // the lowest there is -- the caller's stream cannot be put above them) still has CTAs to dispatch goes first
  cudaStream_t chain;
  cudaEvent_t evIn, evOut;
};
SideStreams* side_streams();

}  // namespace lcb

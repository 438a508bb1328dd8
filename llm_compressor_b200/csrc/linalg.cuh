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

}  // namespace lcb

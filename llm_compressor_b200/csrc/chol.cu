// (b) Cholesky-inverse stage of the layer solvers.
//
// The reference computes  U = cholesky( cholesky_inverse( cholesky(H) ), upper )
// (ref: gptq/core.py:213-224, gptaq/core.py:258-269, sparsegpt/core.py:179-190), i.e. the upper
// factor of H^-1 = U^T U.  Algebraically U = R^-1 where H = R R^T with R upper triangular, and R is
// the index-reversed lower Cholesky factor of the index-reversed matrix.  So instead of
// potrf + potri + potrf (4K^3/3 flop) this file does ONE hand-written blocked Cholesky of J H J
// and ONE blocked triangular inverse (2K^3/3 flop), writing J L^-1 J back as U:
//   1. gather (optional act-order permutation) + index reversal + damping  -> work matrix
//   2. right-looking blocked potrf, 128-wide panels: diagonal block factored (and inverted) by
//      one CTA in shared memory, panel TRSM and trailing SYRK as fp32 GEMMs
//   3. recursive blocked triangular inverse: batched GEMMs per level (log2(K/128) levels)
#include <algorithm>

#include "linalg.cuh"

namespace lcb {

namespace {

constexpr int NB = 128;

// H[dead,dead] = 1 where diag(H) == 0 (ref: gptq/core.py:175-176); dead flags out.
__global__ void dead_fix_kernel(float* H, int64_t k, uint8_t* dead) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const bool d = H[i * k + i] == 0.0f;
  if (d) H[i * k + i] = 1.0f;
  if (dead) dead[i] = d ? 1 : 0;
}

// out[0] = sum(diag(H)) (single CTA; k <= 16K so this is tiny)
__global__ void diag_sum_kernel(const float* H, int64_t k, float* out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < k; i += blockDim.x) s += H[i * k + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = s;
  }
}

// A[i][j] = H[p(k-1-i)][p(k-1-j)] + (i == j) * damp * mean(diag H)     (perm optional)
__global__ void gather_reverse_damp_kernel(float* __restrict__ A, const float* __restrict__ H,
                                           const int64_t* __restrict__ perm, int64_t k, float damp,
                                           const float* diag_sum) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= k) return;
  int64_t si = k - 1 - i, sj = k - 1 - j;
  if (perm) { si = perm[si]; sj = perm[sj]; }
  float v = H[si * k + sj];
  if (i == j) v += damp * (diag_sum[0] / (float)k);
  A[i * k + j] = v;
}

// U[i][j] = Linv[k-1-i][k-1-j]  (Linv lower triangular with zero upper part -> U upper)
__global__ void reverse_out_kernel(float* __restrict__ U, const float* __restrict__ Linv, int64_t k) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= k) return;
  U[i * k + j] = (j >= i) ? Linv[(k - 1 - i) * k + (k - 1 - j)] : 0.0f;
}

__global__ void zero_strict_upper_kernel(float* A, int64_t k) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j < k && j > i) A[i * k + j] = 0.0f;
}

// One CTA (256 threads): Cholesky of the nb x nb diagonal block at A (ld) and the inverse of the
// factor (`dinv` [NB x NB], turns the panel TRSM into a GEMM).  The 128 x 128 block lives in
// registers, 8 x 8 per thread (thread (ty, tx) owns rows ty*8.., columns tx*8..).  Both phases are
// blocked by 8: per panel one thread factors / inverts an 8 x 8 diagonal tile in registers, the
// panel tiles are exchanged through shared memory and every other thread does 8x8x8 register
// FMAs -- 16 panel steps of ~3 barriers instead of 128 latency-bound column steps.
__global__ void __launch_bounds__(256) potf2_inv_kernel(float* A, int64_t ld, int nb, float* dinv, uint32_t* status) {
  constexpr int PS = 9;                   // padded panel row stride
  __shared__ float P[NB * PS];            // column panel  P[r][m] = L[r][p*8 + m]
  __shared__ float XP[8 * (NB + 4)];      // row panel     XP[m][c] = X[k*8 + m][c]
  __shared__ float dall[16][64];          // inverted diagonal tiles
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  float a[8][8], x[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int r = ty * 8 + i, cc = tx * 8 + c;
      float v = (r == cc) ? 1.0f : 0.0f;  // identity padding for a short last block
      if (r < nb && cc < nb) v = (cc <= r) ? A[(int64_t)r * ld + cc] : 0.0f;
      a[i][c] = v;
      x[i][c] = (r == cc) ? 1.0f : 0.0f;
    }
  bool bad = false;

  // ================= factorisation, right-looking over 16 panels of 8 columns
#pragma unroll 1
  for (int p = 0; p < 16; ++p) {
    if (ty == p && tx == p) {
      // unblocked Cholesky of the 8 x 8 diagonal tile, then its inverse
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = a[j][j];
        if (!(d > 0.0f)) bad = true;
        const float l = sqrtf(d);
        const float rl = __frcp_rn(l);
        a[j][j] = l;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) a[i][j] *= rl;
#pragma unroll
        for (int c = j + 1; c < 8; ++c)
#pragma unroll
          for (int i = c; i < 8; ++i) a[i][c] = fmaf(-a[i][j], a[c][j], a[i][c]);
      }
      float di[8][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int i = 0; i < 8; ++i) di[i][j] = 0.0f;
        const float rj = __frcp_rn(a[j][j]);
        di[j][j] = rj;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) {
          float sacc = 0.f;
#pragma unroll
          for (int m = j; m < i; ++m) sacc = fmaf(a[i][m], di[m][j], sacc);
          di[i][j] = -sacc * __frcp_rn(a[i][i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          dall[p][i * 8 + c] = di[i][c];
          if (c > i) a[i][c] = 0.0f;
        }
    }
    __syncthreads();
    if (tx == p && ty >= p) {
      if (ty > p) {  // L_ip = A_ip * L_pp^-T : new[i][c] = sum_{m <= c} a[i][m] * Dinv[c][m]
        float dv[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) dv[i][c] = dall[p][i * 8 + c];
        float nw[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m <= c; ++m) acc = fmaf(a[i][m], dv[c][m], acc);
            nw[i][c] = acc;
          }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) a[i][c] = nw[i][c];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) P[(ty * 8 + i) * PS + c] = a[i][c];
    }
    __syncthreads();
    if (ty > p && tx > p && tx <= ty) {  // trailing update of the lower tiles
      float lr[8][8], lc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          lr[i][m] = P[(ty * 8 + i) * PS + m];
          lc[i][m] = P[(tx * 8 + i) * PS + m];
        }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float acc = a[i][c];
#pragma unroll
          for (int m = 0; m < 8; ++m) acc = fmaf(-lr[i][m], lc[c][m], acc);
          a[i][c] = acc;
        }
    }
    // the next panel's P is written only after the next iteration's first barrier
  }
  if (bad && status) atomicOr(status, LCB_ST_NOT_SPD);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int r = ty * 8 + i, cc = tx * 8 + c;
      if (r < nb && cc < nb) A[(int64_t)r * ld + cc] = (cc <= r) ? a[i][c] : 0.0f;
    }
  __syncthreads();

  // ================= X = L^-1, blocked forward substitution over the 16 row panels
#pragma unroll 1
  for (int k = 0; k < 16; ++k) {
    if (ty == k && tx <= k) {  // row panel k becomes final: X_k,: = Dinv_k * (accumulated tile)
      float dv[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) dv[i][c] = dall[k][i * 8 + c];
      float nw[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int m = 0; m <= i; ++m) acc = fmaf(dv[i][m], x[m][c], acc);
          nw[i][c] = acc;
        }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          x[i][c] = nw[i][c];
          XP[i * (NB + 4) + tx * 8 + c] = nw[i][c];
        }
    }
    if (tx == k && ty > k) {  // column panel k of L below the diagonal tile
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) P[(ty * 8 + i) * PS + c] = a[i][c];
    }
    __syncthreads();
    if (ty > k && tx <= k) {  // X_i,: -= L_ik * X_k,:
      float lr[8][8], xr[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          lr[i][m] = P[(ty * 8 + i) * PS + m];
          xr[i][m] = XP[i * (NB + 4) + tx * 8 + m];  // xr[m'][c]: row m' of the panel, columns of my tile
        }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float acc = x[i][c];
#pragma unroll
          for (int m = 0; m < 8; ++m) acc = fmaf(-lr[i][m], xr[m][c], acc);
          x[i][c] = acc;
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int r = ty * 8 + i, cc = tx * 8 + c;
      if (r < nb) dinv[r * NB + cc] = (cc <= r && cc < nb) ? x[i][c] : 0.0f;
    }
}

// copy the inverted diagonal blocks (dinv_all: [nblk][NB][NB]) onto the diagonal of A
__global__ void put_diag_blocks_kernel(float* A, int64_t k, const float* dinv_all) {
  const int b = blockIdx.x;
  const int nb = (int)min((int64_t)NB, k - (int64_t)b * NB);
  for (int e = threadIdx.x; e < nb * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    if (c < nb) A[((int64_t)b * NB + r) * k + (int64_t)b * NB + c] = dinv_all[(int64_t)b * NB * NB + r * NB + c];
  }
}

}  // namespace

constexpr int SW = 512;        // super-panel width of the tensor-core path (Kd of the big SYRK)
constexpr int TRI_TG_MIN = 1024;  // trtri levels with node size >= this run on the tensor cores
constexpr int TG_CHAIN = 256;     // accumulation chain length (columns) of the tensor-core contractions

static size_t chol_ws_floats(int64_t k) {
  const int64_t nblk = ceil_div(k, NB);
  const int64_t kp = ceil_div(k, 4) * 4;
  const size_t t_exact = (size_t)(k * k / 2 + k * NB);
  const size_t t_tg = (size_t)(9 * (kp / 2 + NB) * (kp / 2 + NB));  // planes of one trtri node
  return (size_t)(k * k)                 // work matrix
         + std::max(t_exact, t_tg)       // T of the trtri recursion / operand planes
         + (size_t)(nblk * NB * NB)      // inverted diagonal blocks
         + (size_t)(3 * kp * NB)         // TRSM panel + its hi / lo planes
         + (size_t)(2 * kp * SW)         // hi / lo planes of a super-panel strip
         + 64;
}

}  // namespace lcb

using namespace lcb;

extern "C" size_t lcb_chol_ws_bytes(int64_t k) { return chol_ws_floats(k) * sizeof(float); }

extern "C" int lcb_hessian_dead_fix(float* H, int64_t k, uint8_t* dead, void* stream) {
  LCB_REQUIRE(H != nullptr && k > 0, "lcb_hessian_dead_fix: bad arguments");
  dead_fix_kernel<<<(unsigned)ceil_div(k, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(H, k, dead);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_chol_inv_upper(const float* H, float* U, int64_t k, const int64_t* perm, float damp, void* ws,
                                  size_t ws_bytes, uint32_t* status, void* stream) {
  LCB_REQUIRE(H != nullptr && U != nullptr && k > 0, "lcb_chol_inv_upper: bad arguments");
  if (ws == nullptr || ws_bytes < lcb_chol_ws_bytes(k)) {
    set_error("lcb_chol_inv_upper: workspace of %zu bytes needed, %zu given", lcb_chol_ws_bytes(k), ws_bytes);
    return LCB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nblk = ceil_div(k, NB);
  const int64_t kp = ceil_div(k, 4) * 4;
  float* A = static_cast<float*>(ws);
  float* T = A + k * k;
  const size_t t_floats = std::max((size_t)(k * k / 2 + k * NB), (size_t)(9 * (kp / 2 + NB) * (kp / 2 + NB)));
  float* dinv_all = T + t_floats;
  float* panel = dinv_all + nblk * NB * NB;
  float* panelH = panel + kp * NB;
  float* panelL = panelH + kp * NB;
  float* stripH = panelL + kp * NB;
  float* stripL = stripH + kp * SW;
  float* dsum = stripL + kp * SW;
  const bool tg = gemm_mode() == 1 && k % 4 == 0 && tg_ok(ws, 4) && k > NB;

  diag_sum_kernel<<<1, 1024, 0, st>>>(H, k, dsum);
  LCB_LAUNCH_CHECK();
  dim3 g2((unsigned)ceil_div(k, 256), (unsigned)k);
  gather_reverse_damp_kernel<<<g2, 256, 0, st>>>(A, H, perm, k, damp, dsum);
  LCB_LAUNCH_CHECK();

  int rc;
  if (!tg) {
    // ---- blocked right-looking Cholesky (lower) of A, exact fp32
    for (int64_t b = 0; b < nblk; ++b) {
      const int64_t j = b * NB;
      const int nb = (int)std::min<int64_t>(NB, k - j);
      float* Ajj = A + j * k + j;
      float* dinv = dinv_all + b * NB * NB;
      potf2_inv_kernel<<<1, 256, 0, st>>>(Ajj, k, nb, dinv, status);
      LCB_LAUNCH_CHECK();
      const int64_t m2 = k - j - nb;
      if (m2 <= 0) break;
      float* A21 = A + (j + nb) * k + j;
      // L21 = A21 * L11^-T   (panel, ld NB)
      rc = sgemm(gemm_args(A21, k, dinv, NB, panel, NB, (int)m2, nb, nb, 1.0f, 0.0f, /*transB=*/1), st);
      if (rc != LCB_OK) return rc;
      LCB_CUDA(cudaMemcpy2DAsync(A21, (size_t)k * sizeof(float), panel, NB * sizeof(float), (size_t)nb * sizeof(float),
                                 (size_t)m2, cudaMemcpyDeviceToDevice, st));
      // A22 -= L21 * L21^T   (lower tiles only)
      rc = sgemm(gemm_args(panel, NB, panel, NB, A + (j + nb) * k + (j + nb), k, (int)m2, (int)m2, nb, -1.0f, 1.0f, 1,
                           GEMM_LOWER_OUT), st);
      if (rc != LCB_OK) return rc;
    }
  } else {
    // ---- tensor-core path: two-level right-looking Cholesky.  Inside a super-panel of SW columns the
    // 128-wide panel updates touch only the rest of the super-panel strip (Kd = 128, small N); the
    // trailing matrix beyond the strip gets one SYRK with Kd = SW per super-panel.
    for (int64_t j0 = 0; j0 < k; j0 += SW) {
      const int64_t j1 = std::min<int64_t>(j0 + SW, k);
      for (int64_t j = j0; j < j1; j += NB) {
        const int nb = (int)std::min<int64_t>(NB, k - j);
        float* Ajj = A + j * k + j;
        float* dinv = dinv_all + (j / NB) * NB * NB;
        potf2_inv_kernel<<<1, 256, 0, st>>>(Ajj, k, nb, dinv, status);
        LCB_LAUNCH_CHECK();
        const int64_t m2 = k - j - nb;
        if (m2 <= 0) break;
        float* A21 = A + (j + nb) * k + j;
        // L21 = A21 * L11^-T in place (a CTA reads its 128 rows completely before it writes them)
        rc = sgemm(gemm_args(A21, k, dinv, NB, A21, k, (int)m2, nb, nb, 1.0f, 0.0f, /*transB=*/1), st);
        if (rc != LCB_OK) return rc;
        const int64_t nrest = j1 - (j + nb);  // columns of the strip still to be updated
        if (nrest > 0) {
          if ((rc = split_tf32(A21, k, (int)m2, nb, panelH, panelL, NB, 0, st)) != LCB_OK) return rc;
          rc = tgemm_nt(panelH, panelL, NB, panelH, panelL, NB, A + (j + nb) * k + (j + nb), k, (int)m2, (int)nrest, nb,
                        -1.0f, TG_LOWER_OUT, st);
          if (rc != LCB_OK) return rc;
        }
      }
      const int64_t m3 = k - j1;
      if (m3 > 0) {
        const int kd = (int)(j1 - j0);
        if ((rc = split_tf32(A + j1 * k + j0, k, (int)m3, kd, stripH, stripL, SW, 0, st)) != LCB_OK) return rc;
        rc = tgemm_nt(stripH, stripL, SW, stripH, stripL, SW, A + j1 * k + j1, k, (int)m3, (int)m3, kd, -1.0f,
                      TG_LOWER_OUT, st, TG_CHAIN);
        if (rc != LCB_OK) return rc;
      }
    }
  }
  zero_strict_upper_kernel<<<g2, 256, 0, st>>>(A, k);
  LCB_LAUNCH_CHECK();
  put_diag_blocks_kernel<<<(unsigned)nblk, 256, 0, st>>>(A, k, dinv_all);
  LCB_LAUNCH_CHECK();

  // ---- recursive triangular inverse, in place: [[A,0],[C,B]]^-1 = [[A^-1,0],[-B^-1 C A^-1, B^-1]]
  for (int64_t s = NB; s < k; s *= 2) {
    const int64_t nodes = ceil_div(k, 2 * s);
    if (tg && s >= TRI_TG_MIN) {
      // tensor cores, node by node:  T^T = A^-T C^T  and  C <- -B^-1 T  as two NT GEMMs on hi/lo planes
      for (int64_t t0 = 0; t0 < nodes; ++t0) {
        const int64_t base = t0 * 2 * s;
        const int64_t sB = std::min<int64_t>(s, k - base - s);
        if (sB <= 0) break;
        float* Ainv = A + base * k + base;
        float* Cblk = A + (base + s) * k + base;
        float* Binv = A + (base + s) * k + (base + s);
        float* ATh = T;                 // [s, s]   planes of Ainv^T (upper triangular)
        float* ATl = ATh + s * s;
        float* Ch = ATl + s * s;        // [sB, s]  planes of C
        float* Cl = Ch + sB * s;
        float* TT = Cl + sB * s;        // [s, sB]  T^T
        float* TTh = TT + s * sB;
        float* TTl = TTh + s * sB;
        float* Bh = TTl + s * sB;       // [sB, sB] planes of Binv (lower triangular)
        float* Bl = Bh + sB * sB;
        if ((rc = split_tf32(Ainv, k, (int)s, (int)s, ATh, ATl, s, 1, st)) != LCB_OK) return rc;
        if ((rc = split_tf32(Cblk, k, (int)sB, (int)s, Ch, Cl, s, 0, st)) != LCB_OK) return rc;
        if ((rc = split_tf32(Binv, k, (int)sB, (int)sB, Bh, Bl, sB, 0, st)) != LCB_OK) return rc;
        rc = tgemm_nt(ATh, ATl, s, Ch, Cl, s, TT, sB, (int)s, (int)sB, (int)s, 1.0f, TG_STORE | TG_A_UPPER, st, TG_CHAIN);
        if (rc != LCB_OK) return rc;
        if ((rc = split_tf32(TT, sB, (int)s, (int)sB, TTh, TTl, sB, 0, st)) != LCB_OK) return rc;
        rc = tgemm_nt(Bh, Bl, sB, TTh, TTl, sB, Cblk, k, (int)sB, (int)s, (int)sB, -1.0f, TG_STORE | TG_A_LOWER, st, TG_CHAIN);
        if (rc != LCB_OK) return rc;
      }
      continue;
    }
    for (int64_t t0 = 0; t0 < nodes;) {
      // batch consecutive nodes with the same (full) size; the ragged last node goes alone
      const int64_t base = t0 * 2 * s;
      const int64_t sB0 = std::min<int64_t>(s, k - base - s);
      if (sB0 <= 0) break;
      int64_t cnt = 1;
      if (sB0 == s) {
        while (t0 + cnt < nodes && (k - (t0 + cnt) * 2 * s - s) >= s) ++cnt;
      }
      float* Ainv = A + base * k + base;
      float* Cblk = A + (base + s) * k + base;
      float* Binv = A + (base + s) * k + (base + s);
      GemmArgs g1 = gemm_args(Cblk, k, Ainv, k, T, s, (int)sB0, (int)s, (int)s, 1.0f, 0.0f, 0, GEMM_B_LOWER);
      g1.batch = (int)cnt; g1.strideA = g1.strideB = 2 * s * (k + 1); g1.strideC = s * s;
      rc = sgemm(g1, st);
      if (rc != LCB_OK) return rc;
      GemmArgs g2a = gemm_args(Binv, k, T, s, Cblk, k, (int)sB0, (int)s, (int)sB0, -1.0f, 0.0f, 0, GEMM_A_LOWER);
      g2a.batch = (int)cnt; g2a.strideA = g2a.strideC = 2 * s * (k + 1); g2a.strideB = s * s;
      rc = sgemm(g2a, st);
      if (rc != LCB_OK) return rc;
      t0 += cnt;
    }
  }
  reverse_out_kernel<<<g2, 256, 0, st>>>(U, A, k);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

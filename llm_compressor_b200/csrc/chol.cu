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

__global__ void zero_strict_upper_kernel(float* A, int64_t k, int64_t i0) {  // rows i0 .. i0 + gridDim.y - 1
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = i0 + blockIdx.y;
  if (j < k && j > i) A[i * k + j] = 0.0f;
}

// Register-tiled 128 x 128 building blocks shared by the Cholesky kernels.  256 threads; thread (ty, tx)
// owns the 8 x 8 tile at rows ty*8.., columns tx*8.. in registers.  Both phases are blocked by 8: per
// panel one thread factors / inverts an 8 x 8 diagonal tile, the panel tiles are exchanged through
// shared memory and every other thread does 8x8x8 register FMAs -- 16 panel steps of ~3 barriers
// instead of 128 latency-bound column steps.
constexpr int PS = 9;  // padded panel row stride

// in: a = lower triangle (incl. diagonal) of an SPD block; out: a = its Cholesky factor L (upper tiles
// zero), dall[p] = inverse of the p-th 8 x 8 diagonal tile of L.  P: NB * PS floats of scratch.
__device__ __forceinline__ void factor_block(float (&a)[8][8], float* P, float (*dall)[64], bool& bad, int ty, int tx) {
  // ================= factorisation, right-looking over 16 panels of 8 columns
#pragma unroll 1
  for (int p = 0; p < 16; ++p) {
    if (ty == p && tx == p) {
      // unblocked Cholesky of the 8 x 8 diagonal tile, then its inverse.  One rsqrt (MUFU + one Newton
      // step, ~1 ulp) per pivot gives both l = d * rsqrt(d) and 1/l: the pivot chain is the critical path
      // of the whole factorisation, so sqrt-then-reciprocal (two long-latency sequences) is avoided.
      float rl[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = a[j][j];
        if (!(d > 0.0f)) bad = true;
        float r = rsqrtf(d);
        r = fmaf(r, fmaf(-0.5f * d * r, r, 0.5f), r);  // r += r * (0.5 - 0.5 * d * r^2)
        rl[j] = r;
        a[j][j] = d * r;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) a[i][j] *= r;
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
        di[j][j] = rl[j];
#pragma unroll
        for (int i = j + 1; i < 8; ++i) {
          float sacc = 0.f;
#pragma unroll
          for (int m = j; m < i; ++m) sacc = fmaf(a[i][m], di[m][j], sacc);
          di[i][j] = -sacc * rl[i];
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
}

constexpr int LS = NB + 1;  // row stride of L11 in shared memory: 8 rows apart = 8 banks apart (TRSM half of chol_panel_kernel)

// ---------------------------------------------------------------------------------------------------------------
// Balanced factorisation for the panel kernel.  ncu on the register-tiled version (potf2-only launch, 35 us): 9 of 32
// lanes active on average -- the trailing update keeps one 8 x 8 tile per thread, so the triangular shape and the
// shrinking trailing matrix leave most lanes idle while every warp still issues the full 512-FMA tile update.
// Here the block lives in shared memory as 4 x 4 micro-tiles in tile-major lower-triangular order (tile (ti, tk),
// tk <= ti, at T + (ti (ti + 1) / 2 + tk) * 16: a thread moves its tile with four 128-bit accesses, consecutive
// threads touch consecutive 64-byte chunks), and each of the 16 panel steps hands out exactly the work that exists:
//   P1  thread 0 factors the 8 x 8 diagonal tile (3 micro-tiles) and publishes L_pp and the pivot reciprocals;
//   P2  one thread per row below solves its 8 entries by substitution and writes them back and to the panel buffer
//       Pn[m][row];
//   P3  the g (g + 1) / 2 micro-tiles of the trailing lower triangle (g = (120 - 8 p) / 4) are dealt round-robin to the
//       256 threads: 16 + 8 128-bit shared accesses for 128 FMAs.
// ~4x fewer warp instructions than the register-tiled steps.  Outputs: Lsm (row-major, stride LS, upper part zero) and
// dall (inverses of the 8 x 8 diagonal tiles).
__device__ __forceinline__ float* mtile(float* T, int ti, int tk) { return T + (((ti * (ti + 1)) >> 1) + tk) * 16; }

__device__ __forceinline__ void tri_index(int t, int& u, int& v) {  // t -> (u, v), v <= u, t = u (u + 1) / 2 + v
  u = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (((u + 1) * (u + 2)) / 2 <= t) ++u;
  while ((u * (u + 1)) / 2 > t) --u;
  v = t - (u * (u + 1)) / 2;
}

__device__ __forceinline__ void factor_block_tiles(const float* Ajj, int64_t ld, int nb, float* T, float* Pn, float* D,
                                                   float* Lsm, float (*dall)[64], bool& bad, int tid) {
  constexpr int NT = NB / 4;                 // 32 micro-tile rows
  constexpr int TILES = NT * (NT + 1) / 2;   // 528
  // ---- load the lower triangle (identity padding for a short last block)
  for (int q = tid; q < TILES; q += 256) {
    int ti, tk;
    tri_index(q, ti, tk);
    float* dst = T + q * 16;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = 4 * ti + a, c0 = 4 * tk;
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = (r == c0 + c) ? 1.0f : 0.0f;
      if (r < nb) {
        if (c0 + 3 < nb) {
          const float4 g = *reinterpret_cast<const float4*>(Ajj + (int64_t)r * ld + c0);
          v[0] = g.x; v[1] = g.y; v[2] = g.z; v[3] = g.w;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) if (c0 + c < nb) v[c] = Ajj[(int64_t)r * ld + c0 + c];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) if (c0 + c > r) v[c] = 0.0f;
      }
      *reinterpret_cast<float4*>(dst + 4 * a) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int p = 0; p < 16; ++p) {
    const int b = 8 * p + 8;     // first trailing row / column
    const int R = NB - b;        // trailing size
    if (tid == 0) {              // ---- P1: the 8 x 8 diagonal tile
      float d[8][8];
      float* t00 = mtile(T, 2 * p, 2 * p);
      float* t10 = mtile(T, 2 * p + 1, 2 * p);
      float* t11 = mtile(T, 2 * p + 1, 2 * p + 1);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          d[i][c] = t00[4 * i + c];
          d[i][c + 4] = 0.0f;
          d[i + 4][c] = t10[4 * i + c];
          d[i + 4][c + 4] = t11[4 * i + c];
        }
      float rl[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float dv = d[jj][jj];
        if (!(dv > 0.0f)) bad = true;
        float r = rsqrtf(dv);
        r = fmaf(r, fmaf(-0.5f * dv * r, r, 0.5f), r);
        rl[jj] = r;
        d[jj][jj] = dv * r;
#pragma unroll
        for (int i = jj + 1; i < 8; ++i) d[i][jj] *= r;
#pragma unroll
        for (int c = jj + 1; c < 8; ++c)
#pragma unroll
          for (int i = c; i < 8; ++i) d[i][c] = fmaf(-d[i][jj], d[c][jj], d[i][c]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c > i) d[i][c] = 0.0f;
          D[i * 8 + c] = d[i][c];
        }
        D[64 + i] = rl[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          t00[4 * i + c] = d[i][c];
          t10[4 * i + c] = d[i + 4][c];
          t11[4 * i + c] = d[i + 4][c + 4];
        }
    }
    __syncthreads();
    if (tid < R) {               // ---- P2: row i of the panel, x = a * L_pp^-T by substitution
      const int i = b + tid;
      const int ti = i >> 2, a = i & 3;
      float* r0 = mtile(T, ti, 2 * p) + 4 * a;
      float* r1 = mtile(T, ti, 2 * p + 1) + 4 * a;
      const float4 x0 = *reinterpret_cast<const float4*>(r0), x1 = *reinterpret_cast<const float4*>(r1);
      float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float acc = x[c];
#pragma unroll
        for (int m = 0; m < c; ++m) acc = fmaf(-x[m], D[c * 8 + m], acc);
        x[c] = acc * D[64 + c];
      }
      *reinterpret_cast<float4*>(r0) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(r1) = make_float4(x[4], x[5], x[6], x[7]);
#pragma unroll
      for (int m = 0; m < 8; ++m) Pn[m * NB + i] = x[m];
    }
    __syncthreads();
    {                            // ---- P3: trailing micro-tiles
      const int g = R >> 2;
      const int items = (g * (g + 1)) >> 1;
      for (int t = tid; t < items; t += 256) {
        int u, v;
        tri_index(t, u, v);
        const int ti = 2 * p + 2 + u, tk = 2 * p + 2 + v;
        float* ct = mtile(T, ti, tk);
        float4 c4[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) c4[a] = *reinterpret_cast<const float4*>(ct + 4 * a);
        float acc[4][4] = {{c4[0].x, c4[0].y, c4[0].z, c4[0].w}, {c4[1].x, c4[1].y, c4[1].z, c4[1].w},
                           {c4[2].x, c4[2].y, c4[2].z, c4[2].w}, {c4[3].x, c4[3].y, c4[3].z, c4[3].w}};
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const float4 ra = *reinterpret_cast<const float4*>(Pn + m * NB + 4 * ti);
          const float4 ca = *reinterpret_cast<const float4*>(Pn + m * NB + 4 * tk);
          const float rr[4] = {ra.x, ra.y, ra.z, ra.w}, cc[4] = {ca.x, ca.y, ca.z, ca.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(-rr[a], cc[c], acc[a][c]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) *reinterpret_cast<float4*>(ct + 4 * a) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      }
    }
    __syncthreads();
  }
  if (tid < 16) {  // inverses of the 8 x 8 diagonal tiles, all 16 in parallel
    float l[8][8], di[8][8];
    const float* t00 = mtile(T, 2 * tid, 2 * tid);
    const float* t10 = mtile(T, 2 * tid + 1, 2 * tid);
    const float* t11 = mtile(T, 2 * tid + 1, 2 * tid + 1);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        l[i][c] = t00[4 * i + c];
        l[i][c + 4] = 0.0f;
        l[i + 4][c] = t10[4 * i + c];
        l[i + 4][c + 4] = t11[4 * i + c];
      }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
      for (int i = 0; i < 8; ++i) di[i][jj] = 0.0f;
      di[jj][jj] = __frcp_rn(l[jj][jj]);
#pragma unroll
      for (int i = jj + 1; i < 8; ++i) {
        float sacc = 0.f;
#pragma unroll
        for (int m = jj; m < i; ++m) sacc = fmaf(l[i][m], di[m][jj], sacc);
        di[i][jj] = -sacc * __frcp_rn(l[i][i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) dall[tid][i * 8 + c] = di[i][c];
  }
  // row-major copy for the TRSM half / the side buffer (upper part zero)
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e >> 7, c = e & (NB - 1);
    Lsm[r * LS + c] = (c <= r) ? mtile(T, r >> 2, c >> 2)[4 * (r & 3) + (c & 3)] : 0.0f;
  }
  __syncthreads();
}

// x = L^-1 for the factor held in `a` (dall from factor_block); x must enter as the identity tiles.
// P: NB * PS floats, XP: 8 * (NB + 4) floats of scratch.
__device__ __forceinline__ void invert_block(const float (&a)[8][8], float (&x)[8][8], float* P, float* XP,
                                             float (*dall)[64], int ty, int tx) {
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
}

// One CTA: Cholesky of the nb x nb diagonal block at A (ld) in place and the inverse of the factor
// (`dinv` [NB x NB], turns the panel TRSM into a GEMM).  Exact-fp32 path of lcb_chol_inv_upper.
__global__ void __launch_bounds__(256) potf2_inv_kernel(float* A, int64_t ld, int nb, float* dinv, uint32_t* status) {
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
  factor_block(a, P, dall, bad, ty, tx);
  if (bad && status) atomicOr(status, LCB_ST_NOT_SPD);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int r = ty * 8 + i, cc = tx * 8 + c;
      if (r < nb && cc < nb) A[(int64_t)r * ld + cc] = (cc <= r) ? a[i][c] : 0.0f;
    }
  __syncthreads();
  invert_block(a, x, P, XP, dall, ty, tx);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int r = ty * 8 + i, cc = tx * 8 + c;
      if (r < nb) dinv[r * NB + cc] = (cc <= r && cc < nb) ? x[i][c] : 0.0f;
    }
}

// Tensor-core path, one launch per 128-column panel: every CTA factors the diagonal block itself
// (redundantly -- the other SMs would idle otherwise, and it saves a launch plus a grid-wide dependency),
// keeps L11 in shared memory and solves its own 128-row tile of the panel, L21 = A21 * L11^-T, with the same
// 8-wide blocked steps (the 8 x 8 diagonal inverses come out of the factorisation).  The tile is written
// back in place and as tf32 hi / lo planes (ld NB) for the SYRK GEMMs that follow.
constexpr int PANEL_SMEM = (NB * LS + 2 * NB * PS + 16 * 64 + 528 * 16) * (int)sizeof(float);

__global__ void __launch_bounds__(256) chol_panel_kernel(float* A, int64_t ld, int64_t j, int nb, int64_t m2,
                                                         float* Lblk, float* panelH, float* panelL, uint32_t* status) {
  extern __shared__ __align__(16) float psm[];
  float* Lsm = psm;                      // [NB][LS]  L11
  float* P0 = Lsm + NB * LS;             // two column-panel buffers
  float* P1 = P0 + NB * PS;
  float(*dall)[64] = reinterpret_cast<float(*)[64]>(P1 + NB * PS);
  const int tid = threadIdx.x;
  float* Ajj = A + j * ld + j;
  float* Tt = reinterpret_cast<float*>(dall) + 16 * 64;   // [528][16] micro-tiles of the diagonal block
  bool bad = false;
  factor_block_tiles(Ajj, ld, nb, Tt, P0, P1, Lsm, dall, bad, tid);  // P0: [8][NB] panel, P1: diagonal tile + pivots
  if (blockIdx.x == 0) {
    // L11 goes to a side buffer [NB x NB]: the other CTAs of this launch are still reading Ajj
    if (bad && status) atomicOr(status, LCB_ST_NOT_SPD);
    for (int e = tid; e < NB * NB; e += 256) Lblk[e] = Lsm[(e >> 7) * LS + (e & (NB - 1))];
  }
  if (m2 <= 0) return;

  // ---- my row tile of A21:  X = A21 L11^-T.  Rows are independent, so this half needs no block barrier: four
  // adjacent lanes share a pair of rows, lane `qd` owns the 8-column panels qd, qd + 4, qd + 8, qd + 12 of both rows
  // (interleaved: balanced trailing work).  Per 8-column step the owner multiplies its panel by the inverse of the
  // 8 x 8 diagonal tile (from the factorisation), the 16 results go to the other three lanes by shuffle, and every lane
  // updates its own panels to the right.  L11 is read from shared memory (row stride 129: the four panels a warp
  // touches in one load sit 8 banks apart).
  const int rp = tid >> 2, qd = tid & 3;
  const int64_t r0 = (int64_t)blockIdx.x * NB;  // first row of the tile inside the panel
  float* A21 = A + (j + nb + r0) * ld + j;
  float x[2][4][8];
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const bool ok = r0 + 2 * rp + rr < m2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* rpnt = A21 + (int64_t)(2 * rp + rr) * ld + 8 * (qd + 4 * k);
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (ok) { v0 = *reinterpret_cast<const float4*>(rpnt); v1 = *reinterpret_cast<const float4*>(rpnt + 4); }
      x[rr][k][0] = v0.x; x[rr][k][1] = v0.y; x[rr][k][2] = v0.z; x[rr][k][3] = v0.w;
      x[rr][k][4] = v1.x; x[rr][k][5] = v1.y; x[rr][k][6] = v1.z; x[rr][k][7] = v1.w;
    }
  }
  __syncthreads();  // Lsm and dall complete
  // The loops stay ROLLED (slot 0 always holds the panel being finished; finished panels are stored and the register
  // panels shift down): fully unrolled this half was ~200 KB of SASS and ran out of the instruction cache.
#pragma unroll 1
  for (int s = 0; s < 4; ++s) {
#pragma unroll 1
    for (int o = 0; o < 4; ++o) {
      const int p = 4 * s + o;
      if (qd == o) {  // X_p = A_p * L_pp^-T : new[c] = sum_{m <= c} a[m] * Dinv[c][m]
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          float nw[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m <= c; ++m) acc = fmaf(x[rr][0][m], dall[p][c * 8 + m], acc);
            nw[c] = acc;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) x[rr][0][c] = nw[c];
        }
      }
      float xp[2][8];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int c = 0; c < 8; ++c) xp[rr][c] = __shfl_sync(0xffffffffu, x[rr][0][c], o, 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // register panel k holds panel q = qd + 4 (s + k); it lies to the right of p iff k > 0 or qd > o
        if (k < 4 - s && (k > 0 || qd > o)) {
          const float* lbase = Lsm + (8 * (qd + 4 * (s + k))) * LS + 8 * p;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float* lrow = lbase + c * LS;
            float a0 = x[0][k][c], a1 = x[1][k][c];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
              const float l = lrow[m];
              a0 = fmaf(-xp[0][m], l, a0);
              a1 = fmaf(-xp[1][m], l, a1);
            }
            x[0][k][c] = a0;
            x[1][k][c] = a1;
          }
        }
      }
    }
    // panel qd + 4 s is final for both rows: store it (in place + tf32 hi / lo planes), shift the register panels
    const int col = 8 * (qd + 4 * s);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int64_t rl = r0 + 2 * rp + rr;
      if (rl < m2) {
        float* rpnt = A21 + (int64_t)(2 * rp + rr) * ld + col;
        *reinterpret_cast<float4*>(rpnt) = make_float4(x[rr][0][0], x[rr][0][1], x[rr][0][2], x[rr][0][3]);
        *reinterpret_cast<float4*>(rpnt + 4) = make_float4(x[rr][0][4], x[rr][0][5], x[rr][0][6], x[rr][0][7]);
        float h[8], l[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t hb;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[rr][0][c]));
          h[c] = __uint_as_float(hb);
          l[c] = __fsub_rn(x[rr][0][c], h[c]);
        }
        float* hp = panelH + rl * NB + col;
        float* lp = panelL + rl * NB + col;
        *reinterpret_cast<float4*>(hp) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(hp + 4) = make_float4(h[4], h[5], h[6], h[7]);
        *reinterpret_cast<float4*>(lp) = make_float4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<float4*>(lp + 4) = make_float4(l[4], l[5], l[6], l[7]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 8; ++c) x[rr][k][c] = x[rr][k + 1][c];
    }
  }
}

// Inverse of every 128 x 128 diagonal block of the factor, in place (base of the trtri recursion): one CTA
// per block, all blocks in one launch.
__global__ void __launch_bounds__(256) diag_inv_kernel(float* A, int64_t k, const float* Lblk_all, int b0) {
  __shared__ float P[NB * PS];
  __shared__ float XP[8 * (NB + 4)];
  __shared__ float dall[16][64];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t blk = (int64_t)b0 + blockIdx.x;
  const int64_t j = blk * NB;
  const int nb = (int)min((int64_t)NB, k - j);
  float* Ajj = A + j * k + j;
  const float* Lb = Lblk_all + blk * NB * NB;  // factor block (identity padded) from chol_panel_kernel
  float a[8][8], x[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int r = ty * 8 + i, cc = tx * 8 + c;
      a[i][c] = Lb[r * NB + cc];
      x[i][c] = (r == cc) ? 1.0f : 0.0f;
    }
  if (ty == tx) {  // inverse of my 8 x 8 diagonal tile of L
    float di[8][8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
      for (int i = 0; i < 8; ++i) di[i][jj] = 0.0f;
      di[jj][jj] = __frcp_rn(a[jj][jj]);
#pragma unroll
      for (int i = jj + 1; i < 8; ++i) {
        float sacc = 0.f;
#pragma unroll
        for (int m = jj; m < i; ++m) sacc = fmaf(a[i][m], di[m][jj], sacc);
        di[i][jj] = -sacc * __frcp_rn(a[i][i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) dall[ty][i * 8 + c] = di[i][c];
  }
  __syncthreads();
  invert_block(a, x, P, XP, dall, ty, tx);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int r = ty * 8 + i, cc = tx * 8 + c;
      if (r < nb && cc < nb) Ajj[(int64_t)r * k + cc] = (cc <= r) ? x[i][c] : 0.0f;
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

SideStreams* side_streams() {
  static thread_local SideStreams cache[16];
  static thread_local bool ready[16] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) {
    set_error("side_streams: cudaGetDevice failed or device index >= 16");
    return nullptr;
  }
  if (!ready[dev]) {
    SideStreams& c = cache[dev];
    bool ok = true;
    for (int i = 0; i < 3; ++i) ok = ok && cudaStreamCreateWithFlags(&c.s[i], cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c.evP, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c.evT, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2; ++i) {
      ok = ok && cudaEventCreateWithFlags(&c.evB[i], cudaEventDisableTiming) == cudaSuccess;
      ok = ok && cudaEventCreateWithFlags(&c.evS[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
      set_error("side_streams: cannot create CUDA streams / events");
      return nullptr;
    }
    ready[dev] = true;
  }
  return &cache[dev];
}

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
         + (size_t)(4 * kp * NB)         // TRSM panel (exact path) / two sets of panel hi + lo planes
         + (size_t)(4 * kp * SW)         // two sets of hi / lo planes of a super-panel strip
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
  float* panel = dinv_all + nblk * NB * NB;   // exact path: [k, NB]; tensor-core path: planes [set][hi|lo][kp, NB]
  float* strip = panel + 4 * kp * NB;         // planes [set][hi|lo][kp, SW]
  float* dsum = strip + 4 * kp * SW;
  const bool tg = gemm_mode() == 1 && k % 4 == 0 && tg_ok(ws, 4) && k > NB;
  if (tg) LCB_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM));

  diag_sum_kernel<<<1, 1024, 0, st>>>(H, k, dsum);
  LCB_LAUNCH_CHECK();
  dim3 g2((unsigned)ceil_div(k, 256), (unsigned)k);
  gather_reverse_damp_kernel<<<g2, 256, 0, st>>>(A, H, perm, k, damp, dsum);
  LCB_LAUNCH_CHECK();

  // ---- one level of the recursive triangular inverse, in place, nodes [t_begin, t_end) of size 2 s:
  //      [[A,0],[C,B]]^-1 = [[A^-1,0],[-B^-1 C A^-1, B^-1]]   (A^-1, B^-1 already in place)
  auto trtri_nodes = [&](int64_t s, int64_t t_begin, int64_t t_end, cudaStream_t ts) -> int {
    int r;
    if (tg && s >= TRI_TG_MIN) {
      // tensor cores, node by node:  T^T = A^-T C^T  and  C <- -B^-1 T  as two NT GEMMs on hi/lo planes
      for (int64_t t0 = t_begin; t0 < t_end; ++t0) {
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
        if ((r = split_tf32(Ainv, k, (int)s, (int)s, ATh, ATl, s, 1, ts)) != LCB_OK) return r;
        if ((r = split_tf32(Cblk, k, (int)sB, (int)s, Ch, Cl, s, 0, ts)) != LCB_OK) return r;
        if ((r = split_tf32(Binv, k, (int)sB, (int)sB, Bh, Bl, sB, 0, ts)) != LCB_OK) return r;
        r = tgemm_nt(ATh, ATl, s, Ch, Cl, s, TT, sB, (int)s, (int)sB, (int)s, 1.0f, TG_STORE | TG_A_UPPER, ts, TG_CHAIN);
        if (r != LCB_OK) return r;
        if ((r = split_tf32(TT, sB, (int)s, (int)sB, TTh, TTl, sB, 0, ts)) != LCB_OK) return r;
        r = tgemm_nt(Bh, Bl, sB, TTh, TTl, sB, Cblk, k, (int)sB, (int)s, (int)sB, -1.0f, TG_STORE | TG_A_LOWER, ts, TG_CHAIN);
        if (r != LCB_OK) return r;
      }
      return LCB_OK;
    }
    for (int64_t t0 = t_begin; t0 < t_end;) {
      // batch consecutive nodes with the same (full) size; the ragged last node goes alone
      const int64_t base = t0 * 2 * s;
      const int64_t sB0 = std::min<int64_t>(s, k - base - s);
      if (sB0 <= 0) break;
      int64_t cnt = 1;
      if (sB0 == s) {
        while (t0 + cnt < t_end && (k - (t0 + cnt) * 2 * s - s) >= s) ++cnt;
      }
      float* Ainv = A + base * k + base;
      float* Cblk = A + (base + s) * k + base;
      float* Binv = A + (base + s) * k + (base + s);
      GemmArgs g1 = gemm_args(Cblk, k, Ainv, k, T, s, (int)sB0, (int)s, (int)s, 1.0f, 0.0f, 0, GEMM_B_LOWER);
      g1.batch = (int)cnt; g1.strideA = g1.strideB = 2 * s * (k + 1); g1.strideC = s * s;
      if ((r = sgemm(g1, ts)) != LCB_OK) return r;
      GemmArgs g2a = gemm_args(Binv, k, T, s, Cblk, k, (int)sB0, (int)s, (int)sB0, -1.0f, 0.0f, 0, GEMM_A_LOWER);
      g2a.batch = (int)cnt; g2a.strideA = g2a.strideC = 2 * s * (k + 1); g2a.strideB = s * s;
      if ((r = sgemm(g2a, ts)) != LCB_OK) return r;
      t0 += cnt;
    }
    return LCB_OK;
  };

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
    // ---- tensor-core path: two-level right-looking Cholesky with look-ahead.
    // Inside a super-panel of SW columns the 128-wide panel updates touch only the rest of the super-panel
    // strip (Kd = 128); the trailing matrix beyond the strip gets one SYRK with Kd = SW per super-panel.
    // The panel kernels are a latency-bound chain, so every update is split in two: the part the NEXT
    // panel / super-panel needs (one column block / one strip) stays on the caller's stream, the rest runs
    // on a side stream underneath the following panel kernels.  All updates are L2 reduce-adds, so they
    // commute; events order them against the readers, and the hi / lo planes are double-buffered.
    SideStreams* ss = side_streams();
    if (ss == nullptr) return LCB_ERR_CUDA;
    auto fail = [&](int code) {  // nothing may outlive the workspace
      cudaStreamSynchronize(ss->s[0]);
      cudaStreamSynchronize(ss->s[1]);
      cudaStreamSynchronize(ss->s[2]);
      return code;
    };
    int64_t q = 0;  // panel counter
    bool evB_live[2] = {false, false}, evS_live[2] = {false, false};
    int64_t J = 0;
    for (int64_t j0 = 0; j0 < k; j0 += SW, ++J) {
      const int64_t j1 = std::min<int64_t>(j0 + SW, k);
      // strip J was updated by S_A(J-1) (this stream) and S_B(J-2) (side stream 1); S_B(J-2) also used the
      // strip planes of parity J & 1 that this super-panel will overwrite
      if (evS_live[J & 1]) { LCB_CUDA(cudaStreamWaitEvent(st, ss->evS[J & 1], 0)); evS_live[J & 1] = false; }
      for (int64_t j = j0; j < j1; j += NB, ++q) {
        const int nb = (int)std::min<int64_t>(NB, k - j);
        const int64_t m2 = k - j - nb;
        float* pH = panel + (q & 1) * 2 * kp * NB;
        float* pL = pH + kp * NB;
        // column block j was updated by A(q-1) (this stream) and B(q-2) (side stream 0), which also read the
        // panel planes of parity q & 1
        if (evB_live[q & 1]) { LCB_CUDA(cudaStreamWaitEvent(st, ss->evB[q & 1], 0)); evB_live[q & 1] = false; }
        // potf2 of the diagonal block + TRSM of the panel rows + tf32 split, one launch
        const unsigned ctas = m2 > 0 ? (unsigned)ceil_div(m2, NB) : 1u;
        chol_panel_kernel<<<ctas, 256, PANEL_SMEM, st>>>(A, k, j, nb, m2, dinv_all + (j / NB) * NB * NB, pH, pL, status);
        LCB_LAUNCH_CHECK();
        if (m2 <= 0) break;
        const int64_t nrest = j1 - (j + nb);  // columns of the strip still to be updated
        if (nrest > 0) {
          const int64_t nB = nrest - nb;       // columns beyond the next block
          if (nB > 0) {                        // B(q): rows / columns from j + 2 nb on, side stream 0
            LCB_CUDA(cudaEventRecord(ss->evP, st));
            LCB_CUDA(cudaStreamWaitEvent(ss->s[0], ss->evP, 0));
            rc = tgemm_nt(pH + nb * NB, pL + nb * NB, NB, pH + nb * NB, pL + nb * NB, NB,
                          A + (j + 2 * nb) * k + (j + 2 * nb), k, (int)(m2 - nb), (int)nB, nb, -1.0f, TG_LOWER_OUT, ss->s[0]);
            if (rc != LCB_OK) return fail(rc);
            LCB_CUDA(cudaEventRecord(ss->evB[q & 1], ss->s[0]));
            evB_live[q & 1] = true;
          }
          // A(q): the next column block, this stream
          rc = tgemm_nt(pH, pL, NB, pH, pL, NB, A + (j + nb) * k + (j + nb), k, (int)m2, (int)std::min<int64_t>(nb, nrest),
                        nb, -1.0f, TG_LOWER_OUT, st);
          if (rc != LCB_OK) return fail(rc);
        }
      }
      // ---- the square [0, j1) x [0, j1) of the factor is final: its part of the triangular inverse runs NOW on side
      // stream 2, underneath the rest of the (latency-bound) panel chain, instead of after it.  Nodes are emitted in
      // post-order: everything inside this super-panel (diagonal blocks, levels 128 and 256), then every larger node
      // that ends exactly here; what is left for after the chain are the ragged nodes at the matrix edge.
      {
        cudaStream_t ts = ss->s[2];
        LCB_CUDA(cudaEventRecord(ss->evT, st));
        LCB_CUDA(cudaStreamWaitEvent(ts, ss->evT, 0));
        const int64_t rows = j1 - j0;
        dim3 gz((unsigned)ceil_div(k, 256), (unsigned)rows);
        zero_strict_upper_kernel<<<gz, 256, 0, ts>>>(A, k, j0);
        LCB_LAUNCH_CHECK();
        diag_inv_kernel<<<(unsigned)ceil_div(rows, NB), 256, 0, ts>>>(A, k, dinv_all, (int)(j0 / NB));
        LCB_LAUNCH_CHECK();
        for (int64_t s = NB; 2 * s <= SW; s *= 2)
          if ((rc = trtri_nodes(s, j0 / (2 * s), ceil_div(j1, 2 * s), ts)) != LCB_OK) return fail(rc);
        for (int64_t s = SW; s < k; s *= 2)
          if (j1 % (2 * s) == 0 && (rc = trtri_nodes(s, j1 / (2 * s) - 1, j1 / (2 * s), ts)) != LCB_OK) return fail(rc);
      }
      const int64_t m3 = k - j1;
      if (m3 > 0) {
        // the strip is final once the last B of this super-panel is done
        for (int e = 0; e < 2; ++e)
          if (evB_live[e]) { LCB_CUDA(cudaStreamWaitEvent(st, ss->evB[e], 0)); evB_live[e] = false; }
        const int kd = (int)(j1 - j0);
        float* sH = strip + (J & 1) * 2 * kp * SW;
        float* sL = sH + kp * SW;
        if ((rc = split_tf32(A + j1 * k + j0, k, (int)m3, kd, sH, sL, SW, 0, st)) != LCB_OK) return fail(rc);
        const int64_t nS = m3 - SW;  // rows / columns beyond the next strip
        if (nS > 0) {                // S_B(J), side stream 1
          LCB_CUDA(cudaEventRecord(ss->evP, st));
          LCB_CUDA(cudaStreamWaitEvent(ss->s[1], ss->evP, 0));
          rc = tgemm_nt(sH + SW * SW, sL + SW * SW, SW, sH + SW * SW, sL + SW * SW, SW, A + (j1 + SW) * k + (j1 + SW), k,
                        (int)nS, (int)nS, kd, -1.0f, TG_LOWER_OUT, ss->s[1], TG_CHAIN);
          if (rc != LCB_OK) return fail(rc);
          LCB_CUDA(cudaEventRecord(ss->evS[J & 1], ss->s[1]));
          evS_live[J & 1] = true;
        }
        // S_A(J): the next strip, this stream
        rc = tgemm_nt(sH, sL, SW, sH, sL, SW, A + j1 * k + j1, k, (int)m3, (int)std::min<int64_t>(SW, m3), kd, -1.0f,
                      TG_LOWER_OUT, st, TG_CHAIN);
        if (rc != LCB_OK) return fail(rc);
      }
    }
    for (int e = 0; e < 2; ++e) {
      if (evB_live[e]) LCB_CUDA(cudaStreamWaitEvent(st, ss->evB[e], 0));
      if (evS_live[e]) LCB_CUDA(cudaStreamWaitEvent(st, ss->evS[e], 0));
    }
    // ragged nodes at the matrix edge (size 2 s does not divide k), children first; then join stream 2
    for (int64_t s = SW; s < k; s *= 2) {
      const int64_t t_last = k / (2 * s);
      if (k % (2 * s) != 0 && k - t_last * 2 * s > s && (rc = trtri_nodes(s, t_last, t_last + 1, ss->s[2])) != LCB_OK)
        return fail(rc);
    }
    LCB_CUDA(cudaEventRecord(ss->evT, ss->s[2]));
    LCB_CUDA(cudaStreamWaitEvent(st, ss->evT, 0));
  }
  if (!tg) {
    zero_strict_upper_kernel<<<g2, 256, 0, st>>>(A, k, 0);
    LCB_LAUNCH_CHECK();
    put_diag_blocks_kernel<<<(unsigned)nblk, 256, 0, st>>>(A, k, dinv_all);
    LCB_LAUNCH_CHECK();
    for (int64_t s = NB; s < k; s *= 2)
      if ((rc = trtri_nodes(s, 0, ceil_div(k, 2 * s), st)) != LCB_OK) return rc;
  }
  reverse_out_kernel<<<g2, 256, 0, st>>>(U, A, k);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

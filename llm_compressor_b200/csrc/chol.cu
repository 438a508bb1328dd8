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
#include <cstdlib>

#include "linalg.cuh"
#include "tc_tile.cuh"

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

__device__ __forceinline__ void load_block_tiles(const float* Ajj, int64_t ld, int nb, float* T, int tid) {
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
}

// factorisation of the block held in T (micro-tile layout, see above); all 256 threads, T complete and visible on entry
__device__ __forceinline__ void factor_tiles_core(float* T, float* Pn, float* D, float* Lsm, float (*dall)[64], bool& bad,
                                                  int tid) {
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

__device__ __forceinline__ void factor_block_tiles(const float* Ajj, int64_t ld, int nb, float* T, float* Pn, float* D,
                                                   float* Lsm, float (*dall)[64], bool& bad, int tid) {
  load_block_tiles(Ajj, ld, nb, T, tid);
  factor_tiles_core(T, Pn, D, Lsm, dall, bad, tid);
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


// ===============================================================================================================
// Tile-task Cholesky + triangular inverse: ONE persistent launch instead of K/128 dependent panel launches.
//
// The lower triangle of the work matrix is cut into 128 x 128 tiles.  A task owns one tile from start to finish
// with the tile in registers (8 x 8 values per thread):
//   L(i, c), i >= c:  acc = A(i,c) - sum_{j in [c0, c)} L(i,j) L(c,j)^T           (left-looking, 3xTF32 tcgen05)
//                     i == c: potf2 of acc fused with Linv(c) = L(c,c)^-1 -> dinv[c] (factor_invert_la, fp32 SIMT)
//                     i >  c: L(i,c) = acc Linv(c)^T                                 (the panel TRSM as a GEMM)
//   Y(k, c), k <  c:  acc = -sum_{j in [k, c)} Y(k,j) L(c,j)^T ;  Y(k,c) = acc Linv(c)^T     with Y = L^-T (the
//                     transposed triangular inverse; Y(c,c) = Linv(c)^T is written by the diagonal task)
// Tasks are numbered column by column (all L(., c), then all Y(., c)) and handed out by an atomic counter; every
// operand of a task is the result of a LOWER-numbered task, published through a release / acquire flag per tile.
// A CTA that holds a task is running, so whatever a task waits for is being computed: no deadlock, whatever part of
// the grid is resident (no cooperative launch needed).  Tasks accumulate eagerly -- term j is consumed as soon as
// its two tiles are flagged -- so only the last term, the TRSM and the potf2 of the diagonal tile are on the
// critical path of the panel chain; everything else hides underneath it on the other SMs.
// Operand tiles stream L2 -> shared memory with cp.async.cg in [128 rows][32 k] chunks, 16-byte pieces XOR-swizzled
// by (row & 7) -- the K-major SWIZZLE_128B layout the tensor core reads (chol_tiles_tc_kernel below).
struct TileArgs {
  // Every tile a task publishes is stored pre-split for the 3xTF32 products of its consumers: the tf32-rounded value
  // in the matrix itself (hi), the exact remainder x - hi in the companion *lo matrix (hi + lo == x exactly).
  float* A;          // [k, k] work matrix, lower triangle (in: SPD block, out: hi(L) below the diagonal tiles)
  float* Alo;        // [k, k] lo(L)
  float* Y;          // [k, k] hi of the transposed inverse Y = L^-T (upper-triangular tiles)
  float* Ylo;        // [k, k] lo(Y)
  float* dinv;       // [nblk][NB * NB] hi of the inverses of the diagonal tiles of L
  float* dinv_lo;    // [nblk][NB * NB] lo
  uint32_t* flagL;   // [nblk * nblk]
  uint32_t* flagY;   // [nblk * nblk]
  uint32_t* counter;
  uint32_t* status;
  unsigned long long* trace;  // debug (LCB_CHOL_TRACE=1): per task {type|i|c, t_fetch, t_acc, t_diag_flag, t_factor, t_end} in ns
  int64_t ld;
  int k, nblk, c0, c1, with_inv, ntasks;
  int fi_nt;         // T-worker warps of factor_invert_split (1..6); 0 = the joint-worker factor_invert_la (A/B runs)
  int gcols;         // columns per task group (1 = plain column-major order)
  int use_tma;       // operand chunks by TMA (one producer thread) instead of cp.async from 224 threads
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused potf2 + inverse of a 128 x 128 diagonal tile with look-ahead (the serial heart of the panel chain).
// In: T = lower triangle in micro-tile layout (see factor_block_tiles).  Out: W = L^-1 in the same layout.
// L itself is never needed downstream (panel TRSMs multiply by L^-1), so it is not assembled.
// Sixteen 8-column steps.  Warp 0 is the CHAIN warp: lane 0 factors the 8 x 8 diagonal tile of step p and inverts
// it (Dinv), then the warp solves the next 8 rows of the panel and updates the NEXT diagonal tile itself, so
// that step p+1's pivots start ~0.2 us after step p's instead of after the whole trailing update.  Warps 1..7 are
// WORKERS: panel solve x = T(i,p) Dinv^T for the other rows, the inverse's row panel W_p = Dinv W_p, then the
// trailing micro-tiles of T and of W (forward substitution on the identity: W(i,:) -= L(i,p) W_p), the tiles the
// chain warp needs next first.  Hand-offs are named barriers (arrive on the producer side, sync on the consumer
// side); panel buffers are double-buffered by step parity.
// Micro-tile storage of factor_invert_la: row a (0..3) of micro-tile t = ti (ti + 1) / 2 + tk lives in plane a at
// float4 index t, so the lanes of a warp that work on consecutive micro-tiles touch consecutive 16-byte words (the
// tile-major layout of factor_block_tiles puts them 64 bytes apart: 4-way bank conflicts, and the trailing update is
// bound by shared-memory wavefronts, not by FMAs).  Planes are 32 bytes out of phase so that the four rows of one
// micro-tile sit in different banks as well.
constexpr int FI_NT = 528;
constexpr int FI_PLANE = FI_NT * 4 + 8;          // floats per plane
constexpr int FI_T_FLOATS = 4 * FI_PLANE;
constexpr int FI_SMEM_FLOATS = 2 * FI_T_FLOATS + 2 * 8 * NB + 2 * 8 * NB + 2 * 64;
enum { FI_BAR_P1 = 1, FI_BAR_LA = 2, FI_BAR_P3A = 3, FI_BAR_W = 4, FI_BAR_P1B = 5 };  // P1 alternates ids by step parity:
// the chain warp may arrive for step p+1 before every worker has arrived for step p
__device__ __forceinline__ int fi_tidx(int ti, int tk) { return ((ti * (ti + 1)) >> 1) + tk; }
__device__ __forceinline__ float* fi_row(float* T, int ti, int tk, int a) { return T + a * FI_PLANE + (fi_tidx(ti, tk) << 2); }

// C(t) -= sum_m rowp[m][4 ti ..] * colp[m][4 tk ..]
__device__ __forceinline__ void fi_tile_update(float* M, int t, const float* rowp, const float* colp, int ti, int tk) {
  float* ct = M + (t << 2);
  float4 c4[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) c4[a] = *reinterpret_cast<const float4*>(ct + a * FI_PLANE);
  float acc[4][4] = {{c4[0].x, c4[0].y, c4[0].z, c4[0].w}, {c4[1].x, c4[1].y, c4[1].z, c4[1].w},
                     {c4[2].x, c4[2].y, c4[2].z, c4[2].w}, {c4[3].x, c4[3].y, c4[3].z, c4[3].w}};
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float4 ra = *reinterpret_cast<const float4*>(rowp + m * NB + 4 * ti);
    const float4 ca = *reinterpret_cast<const float4*>(colp + m * NB + 4 * tk);
    const float rr[4] = {ra.x, ra.y, ra.z, ra.w}, cc[4] = {ca.x, ca.y, ca.z, ca.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(-rr[a], cc[c], acc[a][c]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
    *reinterpret_cast<float4*>(ct + a * FI_PLANE) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
}

__device__ __forceinline__ void factor_invert_la(float* T, float* W, float* Pn, float* Wp, float* D, bool& bad, int tid,
                                                 long long* dbg = nullptr) {
#define FI_DBG(slot) do { if (dbg) dbg[p * 12 + (slot)] = clock64(); } while (0)
  // W = identity
  for (int q = tid; q < FI_NT; q += 256) {
    int ti, tk;
    tri_index(q, ti, tk);
    const float dg = (ti == tk) ? 1.0f : 0.0f;
    *reinterpret_cast<float4*>(W + 0 * FI_PLANE + 4 * q) = make_float4(dg, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(W + 1 * FI_PLANE + 4 * q) = make_float4(0.f, dg, 0.f, 0.f);
    *reinterpret_cast<float4*>(W + 2 * FI_PLANE + 4 * q) = make_float4(0.f, 0.f, dg, 0.f);
    *reinterpret_cast<float4*>(W + 3 * FI_PLANE + 4 * q) = make_float4(0.f, 0.f, 0.f, dg);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    // ======================================================================= chain warp
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
      float* Dv = D + (p & 1) * 64;      // Dinv of this step, row-major 8 x 8
      float* Pp = Pn + (p & 1) * 8 * NB;
      if (lane == 0) FI_DBG(0);
      if (lane == 0) {
        float d[8][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(fi_row(T, 2 * p, 2 * p, i));
          const float4 b = *reinterpret_cast<const float4*>(fi_row(T, 2 * p + 1, 2 * p, i));
          const float4 c = *reinterpret_cast<const float4*>(fi_row(T, 2 * p + 1, 2 * p + 1, i));
          d[i][0] = a.x; d[i][1] = a.y; d[i][2] = a.z; d[i][3] = a.w;
          d[i][4] = 0.f; d[i][5] = 0.f; d[i][6] = 0.f; d[i][7] = 0.f;
          d[i + 4][0] = b.x; d[i + 4][1] = b.y; d[i + 4][2] = b.z; d[i + 4][3] = b.w;
          d[i + 4][4] = c.x; d[i + 4][5] = c.y; d[i + 4][6] = c.z; d[i + 4][7] = c.w;
        }
        float rl[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float dv = d[jj][jj];
          if (!(dv > 0.0f)) bad = true;
          float r = rsqrtf(dv);
          r = fmaf(r, fmaf(-0.5f * dv * r, r, 0.5f), r);
          rl[jj] = r;
#pragma unroll
          for (int i = jj + 1; i < 8; ++i) d[i][jj] *= r;
#pragma unroll
          for (int c = jj + 1; c < 8; ++c)
#pragma unroll
            for (int i = c; i < 8; ++i) d[i][c] = fmaf(-d[i][jj], d[c][jj], d[i][c]);
        }
        // inverse of the 8 x 8 factor (column by column; the columns are independent chains)
        float di[8][8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
          for (int i = 0; i < 8; ++i) di[i][jj] = 0.0f;
          di[jj][jj] = rl[jj];
#pragma unroll
          for (int i = jj + 1; i < 8; ++i) {
            float sacc = 0.f;
#pragma unroll
            for (int m = jj; m < i; ++m) sacc = fmaf(d[i][m], di[m][jj], sacc);
            di[i][jj] = -sacc * rl[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          *reinterpret_cast<float4*>(Dv + 8 * i) = make_float4(di[i][0], di[i][1], di[i][2], di[i][3]);
          *reinterpret_cast<float4*>(Dv + 8 * i + 4) = make_float4(di[i][4], di[i][5], di[i][6], di[i][7]);
        }
      }
      __syncwarp();
      nbar_arrive((p & 1) ? FI_BAR_P1B : FI_BAR_P1, 256);
      if (lane == 0) FI_DBG(1);
      if (p < 15) {
        if (p >= 1) nbar_sync(FI_BAR_P3A, 256);   // T(p+1,p) and T(p+1,p+1) carry the updates of steps < p
        // ---- look-ahead: the 8 rows below the diagonal tile, then the next diagonal tile
        const int r = lane >> 2, cg = lane & 3;          // row r of the 8, columns 2 cg and 2 cg + 1
        const int gi = 8 * (p + 1) + r;
        const float4 x0 = *reinterpret_cast<const float4*>(fi_row(T, gi >> 2, 2 * p, gi & 3));
        const float4 x1 = *reinterpret_cast<const float4*>(fi_row(T, gi >> 2, 2 * p + 1, gi & 3));
        if (lane == 0) FI_DBG(2);
        const float xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = 2 * cg + q;
          const float4 d0 = *reinterpret_cast<const float4*>(Dv + 8 * c), d1 = *reinterpret_cast<const float4*>(Dv + 8 * c + 4);
          float acc = xr[0] * d0.x;                       // Dinv[c][m] == 0 for m > c
          acc = fmaf(xr[1], d0.y, acc); acc = fmaf(xr[2], d0.z, acc); acc = fmaf(xr[3], d0.w, acc);
          acc = fmaf(xr[4], d1.x, acc); acc = fmaf(xr[5], d1.y, acc); acc = fmaf(xr[6], d1.z, acc); acc = fmaf(xr[7], d1.w, acc);
          Pp[c * NB + gi] = acc;
        }
        __syncwarp();
        {
          float li[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) li[m] = Pp[m * NB + gi];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = 2 * cg + q;
            if (c <= r) {
              const int gc = 8 * (p + 1) + c;
              float acc = 0.f;
#pragma unroll
              for (int m = 0; m < 8; ++m) acc = fmaf(li[m], Pp[m * NB + gc], acc);
              float* e = fi_row(T, gi >> 2, gc >> 2, gi & 3) + (gc & 3);
              *e -= acc;
            }
          }
        }
        __syncwarp();
        nbar_arrive(FI_BAR_LA, 256);
        if (lane == 0) FI_DBG(3);
      }
    }
  } else {
    // ======================================================================= workers
    const int wt = tid - 32;   // 0 .. 223
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
      const float* Dv = D + (p & 1) * 64;
      float* Pp = Pn + (p & 1) * 8 * NB;
      float* Wq = Wp + (p & 1) * 8 * NB;
      nbar_sync((p & 1) ? FI_BAR_P1B : FI_BAR_P1, 256);
      float dv[8][8];   // Dinv (lower triangle), broadcast loads
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 d0 = *reinterpret_cast<const float4*>(Dv + 8 * c), d1 = *reinterpret_cast<const float4*>(Dv + 8 * c + 4);
        dv[c][0] = d0.x; dv[c][1] = d0.y; dv[c][2] = d0.z; dv[c][3] = d0.w;
        dv[c][4] = d1.x; dv[c][5] = d1.y; dv[c][6] = d1.z; dv[c][7] = d1.w;
      }
      if (wt == 0) FI_DBG(4);
      {  // ---- panel solve for row i (the chain warp does rows 8(p+1) .. 8(p+1)+7)
        const int i = 8 * (p + 2) + wt;
        if (i < NB) {
          const float4 x0 = *reinterpret_cast<const float4*>(fi_row(T, i >> 2, 2 * p, i & 3));
          const float4 x1 = *reinterpret_cast<const float4*>(fi_row(T, i >> 2, 2 * p + 1, i & 3));
          const float xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m <= c; ++m) acc = fmaf(xr[m], dv[c][m], acc);
            Pp[c * NB + i] = acc;
          }
        }
      }
      {  // ---- row panel p of the inverse: W_p = Dinv W_p (column `col` of the 8 rows), final
        const int col = wt;
        if (col < 8 * p + 8) {
          float w[8];
          const int tk = col >> 2, cc = col & 3;
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const int rr = 8 * p + m;
            w[m] = (tk <= (rr >> 2)) ? fi_row(W, rr >> 2, tk, rr & 3)[cc] : 0.0f;
          }
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q <= m; ++q) acc = fmaf(dv[m][q], w[q], acc);
            const int rr = 8 * p + m;
            if (tk <= (rr >> 2)) fi_row(W, rr >> 2, tk, rr & 3)[cc] = acc;
            Wq[m * NB + col] = acc;
          }
        }
      }
      if (wt == 0) FI_DBG(5);
      if (p < 15) nbar_sync(FI_BAR_LA, 256);   // every panel entry of this step (workers' and the chain warp's) is written
      else nbar_sync(FI_BAR_W, 224);
      if (p < 15) {
        // ---- trailing micro-tiles: T part (lower triangle from micro-row 2(p+1), minus the next diagonal tile that the
        // chain warp updated itself: items 0..2), then W part (micro-rows >= 2(p+1), micro-columns <= 2p+1)
        const int base = 2 * (p + 1);
        const int g = 32 - base;
        const int nA = (g * (g + 1)) >> 1;
        const int nWc = base;                 // micro-columns of the W part
        const int total = (nA - 3) + g * nWc;
        bool first = true;
        for (int it = wt; it < total || first; it += 224) {
          if (it < total) {
            if (it < nA - 3) {
              int u, v;
              tri_index(it + 3, u, v);
              fi_tile_update(T, fi_tidx(base + u, base + v), Pp, Pp, base + u, base + v);
            } else {
              const int w2 = it - (nA - 3);
              const int u = w2 / nWc, v = w2 - u * nWc;
              fi_tile_update(W, fi_tidx(base + u, v), Pp, Wq, base + u, v);
            }
          }
          if (first) {
            first = false;
            if (p < 14) nbar_arrive(FI_BAR_P3A, 256);   // items 3..9 (the chain warp's next inputs) are in the first round
            if (wt == 0) FI_DBG(7);
          }
        }
      }
      if (wt == 0) FI_DBG(8);
      nbar_sync(FI_BAR_W, 224);
    }
  }
  __syncthreads();
#undef FI_DBG
}

// ---------------------------------------------------------------------------------------------------------------
// Split-role variant of factor_invert_la (the default).  In factor_invert_la every worker does its share of BOTH trailing
// updates (T, the factor, and W, the inverse riding along) and all of them meet at the end of each step, so the
// pivot chain waits for the inverse's work too: 3400 cycles per 8-column step where the chain itself needs ~1300.
// Here the inverse is taken OFF the pivot chain:
//   warp 0            chain warp, as above;
//   warps 1 .. nt     T workers: panel solve + trailing update of the factor only (what the next pivots wait for);
//   warps nt+1 .. 7   W workers: row panel of the inverse W_p = Dinv W_p and its trailing update W(i,:) -= L(i,p) W_p.
// The W workers only CONSUME what the other two groups produce (Dinv(p) and the panel L(:, p)); every step's panel and
// Dinv get their own buffer (16 panels = 64 KB), so the factor side never waits for the inverse.  The hand-over is a
// shared-memory counter (fence + atomicAdd by the producers' lane 0, volatile poll by the consumers): named barriers
// cannot express a producer that runs several steps ahead.  Inside a group the steps are ordered by named barriers.
constexpr int FS_PN_FLOATS = 16 * 8 * NB;
constexpr int FS_SMEM_FLOATS = 2 * FI_T_FLOATS + FS_PN_FLOATS + 2 * 8 * NB + 16 * 64 + 8;
enum { FS_BAR_TEND = 4, FS_BAR_WSYNC = 10 };   // P1 / P1B / LA / P3A keep the ids of factor_invert_la

__device__ __forceinline__ void fs_signal(int* cnt, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicAdd(cnt, 1);
  }
}

__device__ __forceinline__ void factor_invert_split(float* T, float* W, float* Pn, float* Wp, float* D, int* cnt, bool& bad,
                                                    int tid, int nt, long long* dbg = nullptr) {
#define FS_DBG(slot) do { if (dbg) dbg[p * 12 + (slot)] = clock64(); } while (0)
  const int TTH = 32 * nt;             // T workers
  const int WTH = 224 - TTH;           // W workers
  const int CT = 32 + TTH;             // chain + T workers: participants of P1 / LA / P3A
  for (int q = tid; q < FI_NT; q += 256) {
    int ti, tk;
    tri_index(q, ti, tk);
    const float dg = (ti == tk) ? 1.0f : 0.0f;
    *reinterpret_cast<float4*>(W + 0 * FI_PLANE + 4 * q) = make_float4(dg, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(W + 1 * FI_PLANE + 4 * q) = make_float4(0.f, dg, 0.f, 0.f);
    *reinterpret_cast<float4*>(W + 2 * FI_PLANE + 4 * q) = make_float4(0.f, 0.f, dg, 0.f);
    *reinterpret_cast<float4*>(W + 3 * FI_PLANE + 4 * q) = make_float4(0.f, 0.f, 0.f, dg);
  }
  if (tid == 0) *cnt = 0;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    // ======================================================================= chain warp
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
      float* Dv = D + p * 64;            // Dinv of this step, row-major 8 x 8
      float* Pp = Pn + p * 8 * NB;
      if (lane == 0) FS_DBG(0);
      if (lane == 0) {
        float d[8][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(fi_row(T, 2 * p, 2 * p, i));
          const float4 b = *reinterpret_cast<const float4*>(fi_row(T, 2 * p + 1, 2 * p, i));
          const float4 c = *reinterpret_cast<const float4*>(fi_row(T, 2 * p + 1, 2 * p + 1, i));
          d[i][0] = a.x; d[i][1] = a.y; d[i][2] = a.z; d[i][3] = a.w;
          d[i][4] = 0.f; d[i][5] = 0.f; d[i][6] = 0.f; d[i][7] = 0.f;
          d[i + 4][0] = b.x; d[i + 4][1] = b.y; d[i + 4][2] = b.z; d[i + 4][3] = b.w;
          d[i + 4][4] = c.x; d[i + 4][5] = c.y; d[i + 4][6] = c.z; d[i + 4][7] = c.w;
        }
        float rl[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float dv = d[jj][jj];
          if (!(dv > 0.0f)) bad = true;
          float r = rsqrtf(dv);
          r = fmaf(r, fmaf(-0.5f * dv * r, r, 0.5f), r);
          rl[jj] = r;
#pragma unroll
          for (int i = jj + 1; i < 8; ++i) d[i][jj] *= r;
#pragma unroll
          for (int c = jj + 1; c < 8; ++c)
#pragma unroll
            for (int i = c; i < 8; ++i) d[i][c] = fmaf(-d[i][jj], d[c][jj], d[i][c]);
        }
        float di[8][8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
          for (int i = 0; i < 8; ++i) di[i][jj] = 0.0f;
          di[jj][jj] = rl[jj];
#pragma unroll
          for (int i = jj + 1; i < 8; ++i) {
            float sacc = 0.f;
#pragma unroll
            for (int m = jj; m < i; ++m) sacc = fmaf(d[i][m], di[m][jj], sacc);
            di[i][jj] = -sacc * rl[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          *reinterpret_cast<float4*>(Dv + 8 * i) = make_float4(di[i][0], di[i][1], di[i][2], di[i][3]);
          *reinterpret_cast<float4*>(Dv + 8 * i + 4) = make_float4(di[i][4], di[i][5], di[i][6], di[i][7]);
        }
      }
      __syncwarp();
      nbar_arrive((p & 1) ? FI_BAR_P1B : FI_BAR_P1, CT);
      if (lane == 0) FS_DBG(1);
      if (p < 15) {
        if (p >= 1) nbar_sync(FI_BAR_P3A, CT);   // T(p+1,p) and T(p+1,p+1) carry the updates of steps < p
        if (lane == 0) FS_DBG(2);
        const int r = lane >> 2, cg = lane & 3;          // row r of the 8, columns 2 cg and 2 cg + 1
        const int gi = 8 * (p + 1) + r;
        const float4 x0 = *reinterpret_cast<const float4*>(fi_row(T, gi >> 2, 2 * p, gi & 3));
        const float4 x1 = *reinterpret_cast<const float4*>(fi_row(T, gi >> 2, 2 * p + 1, gi & 3));
        const float xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = 2 * cg + q;
          const float4 d0 = *reinterpret_cast<const float4*>(Dv + 8 * c), d1 = *reinterpret_cast<const float4*>(Dv + 8 * c + 4);
          float acc = xr[0] * d0.x;                       // Dinv[c][m] == 0 for m > c
          acc = fmaf(xr[1], d0.y, acc); acc = fmaf(xr[2], d0.z, acc); acc = fmaf(xr[3], d0.w, acc);
          acc = fmaf(xr[4], d1.x, acc); acc = fmaf(xr[5], d1.y, acc); acc = fmaf(xr[6], d1.z, acc); acc = fmaf(xr[7], d1.w, acc);
          Pp[c * NB + gi] = acc;
        }
        fs_signal(cnt, lane);            // Dinv(p) + the chain's 8 rows of panel p -> W workers (includes a __syncwarp)
        {
          float li[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) li[m] = Pp[m * NB + gi];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = 2 * cg + q;
            if (c <= r) {
              const int gc = 8 * (p + 1) + c;
              float acc = 0.f;
#pragma unroll
              for (int m = 0; m < 8; ++m) acc = fmaf(li[m], Pp[m * NB + gc], acc);
              float* e = fi_row(T, gi >> 2, gc >> 2, gi & 3) + (gc & 3);
              *e -= acc;
            }
          }
        }
        __syncwarp();
        nbar_arrive(FI_BAR_LA, CT);
        if (lane == 0) FS_DBG(3);
      } else {
        fs_signal(cnt, lane);
      }
    }
  } else if (warp <= nt) {
    // ======================================================================= T workers
    const int tw = tid - 32;   // 0 .. TTH - 1
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
      const float* Dv = D + p * 64;
      float* Pp = Pn + p * 8 * NB;
      nbar_sync((p & 1) ? FI_BAR_P1B : FI_BAR_P1, CT);
      if (tw == 0) FS_DBG(4);
      if (8 * (p + 2) < NB) {  // ---- panel solve for the rows below the chain warp's 8
        float dv[8][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 d0 = *reinterpret_cast<const float4*>(Dv + 8 * c), d1 = *reinterpret_cast<const float4*>(Dv + 8 * c + 4);
          dv[c][0] = d0.x; dv[c][1] = d0.y; dv[c][2] = d0.z; dv[c][3] = d0.w;
          dv[c][4] = d1.x; dv[c][5] = d1.y; dv[c][6] = d1.z; dv[c][7] = d1.w;
        }
        for (int i = 8 * (p + 2) + tw; i < NB; i += TTH) {
          const float4 x0 = *reinterpret_cast<const float4*>(fi_row(T, i >> 2, 2 * p, i & 3));
          const float4 x1 = *reinterpret_cast<const float4*>(fi_row(T, i >> 2, 2 * p + 1, i & 3));
          const float xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m <= c; ++m) acc = fmaf(xr[m], dv[c][m], acc);
            Pp[c * NB + i] = acc;
          }
        }
      }
      fs_signal(cnt, lane);              // this warp's rows of panel p -> W workers
      if (tw == 0) FS_DBG(5);
      if (p < 15) {
        nbar_sync(FI_BAR_LA, CT);        // every panel entry of this step (T workers' and the chain warp's) is written
        // ---- trailing micro-tiles of the factor: lower triangle from micro-row 2(p+1), minus the next diagonal tile
        // (items 0..2: the chain warp updated it itself)
        const int base = 2 * (p + 1);
        const int g = 32 - base;
        const int total = ((g * (g + 1)) >> 1) - 3;
        if (tw == 0) FS_DBG(6);
        bool first = true;
        for (int it = tw; it < total || first; it += TTH) {
          if (it < total) {
            int u, v;
            tri_index(it + 3, u, v);
            fi_tile_update(T, fi_tidx(base + u, base + v), Pp, Pp, base + u, base + v);
          }
          if (first) {
            first = false;
            if (p < 14) nbar_arrive(FI_BAR_P3A, CT);   // items 0..6 (the chain warp's next inputs) are in the first round
            if (tw == 0) FS_DBG(7);
          }
        }
      }
      if (tw == 0) FS_DBG(8);
      nbar_sync(FS_BAR_TEND, TTH);
    }
  } else {
    // ======================================================================= W workers
    const int ww = tid - 32 - TTH;   // 0 .. WTH - 1
    const int need_per_step = 1 + nt;
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
      const float* Dv = D + p * 64;
      const float* Pp = Pn + p * 8 * NB;
      float* Wq = Wp + (p & 1) * 8 * NB;
      if (lane == 0) {
        const volatile int* vc = cnt;
        const int need = need_per_step * (p + 1);
        while (*vc < need) __nanosleep(32);
        __threadfence_block();
      }
      __syncwarp();
      if (ww == 0) FS_DBG(9);
      {  // ---- row panel p of the inverse: W_p = Dinv W_p (column `col` of the 8 rows), final
        float dv[8][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 d0 = *reinterpret_cast<const float4*>(Dv + 8 * c), d1 = *reinterpret_cast<const float4*>(Dv + 8 * c + 4);
          dv[c][0] = d0.x; dv[c][1] = d0.y; dv[c][2] = d0.z; dv[c][3] = d0.w;
          dv[c][4] = d1.x; dv[c][5] = d1.y; dv[c][6] = d1.z; dv[c][7] = d1.w;
        }
        for (int col = ww; col < 8 * p + 8; col += WTH) {
          float w[8];
          const int tk = col >> 2, cc = col & 3;
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const int rr = 8 * p + m;
            w[m] = (tk <= (rr >> 2)) ? fi_row(W, rr >> 2, tk, rr & 3)[cc] : 0.0f;
          }
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q <= m; ++q) acc = fmaf(dv[m][q], w[q], acc);
            const int rr = 8 * p + m;
            if (tk <= (rr >> 2)) fi_row(W, rr >> 2, tk, rr & 3)[cc] = acc;
            Wq[m * NB + col] = acc;
          }
        }
      }
      nbar_sync(FS_BAR_WSYNC, WTH);
      if (p < 15) {
        // ---- trailing micro-tiles of the inverse: micro-rows >= 2(p+1), micro-columns <= 2p+1
        const int base = 2 * (p + 1);
        const int g = 32 - base;
        const int total = g * base;
        const float rbase = 1.0f / (float)base;
        for (int it = ww; it < total; it += WTH) {
          const int u = (int)(((float)it + 0.5f) * rbase), v = it - u * base;   // exact: (it + 0.5) / base is >= 1/60 away from an integer
          fi_tile_update(W, fi_tidx(base + u, v), Pp, Wq, base + u, v);
        }
        nbar_sync(FS_BAR_WSYNC, WTH);
      }
      if (ww == 0) FS_DBG(10);
    }
  }
  __syncthreads();
#undef FS_DBG
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core variant of the tile tasks: the same task graph, but every 128 x 128 x 128 product runs as 3xTF32
// tcgen05.mma (M = N = 128, K = 8) with the accumulator in TMEM.  The [128][32] operand chunks that cp.async
// lays down (16-byte pieces XOR-swizzled by row & 7, 1024-byte 8-row groups) ARE the K-major SWIZZLE_128B
// canonical layout, so the UMMA descriptors point straight at them; the threads only split each chunk in place
// into hi = tf32(x) and a second lo = x - hi plane.  Per chunk: 12 MMAs (lo*hi, hi*lo, hi*hi) issued by one
// thread, completion tracked by tcgen05.commit -> mbarrier per stage; the TMEM accumulator is drained into fp32
// registers every 256 reduction columns (its accumulation truncates: error grows with the chain length).
// A thread holds row 32 (warp & 3) + lane, columns 64 (warp >> 2) .. + 63 of the tile (the tcgen05.ld layout).
constexpr int TT_STAGES = 3;
constexpr int TT_STAGE_BYTES = 4 * TT_PLANE;         // P hi, Q hi, P lo, Q lo
constexpr int TT_XPLANES = 10 * 4096;                // Linv chunk ch keeps rows >= 32 ch only: 16 + 12 + 8 + 4 KB
constexpr int TT_DATA_BYTES = 8 * TT_PLANE + 2 * TT_XPLANES;       // TRSM phase: C hi / lo (128 KB) + Linv hi / lo (80 KB)
constexpr int TT_SMEM = TT_DATA_BYTES + 1024 + 256;  // + alignment slack + barriers
static_assert(TT_DATA_BYTES >= TT_STAGES * TT_STAGE_BYTES, "pipeline stages must fit");
static_assert(TT_DATA_BYTES >= 2 * NB * 132 * 4, "output staging must fit");
enum { TT_BAR_WORK = 6, TT_BAR_FULL0 = 7 };          // named barriers 7, 8, 9: stage s handed to the MMA warp
constexpr int TT_DRAIN = 8;                          // chunks per TMEM accumulation chain (256 columns)
static_assert(TT_STAGES * TT_STAGE_BYTES >= FI_SMEM_FLOATS * 4, "potf2 scratch must fit in the pipeline buffers");
static_assert(TT_STAGES * TT_STAGE_BYTES >= FS_SMEM_FLOATS * 4, "split potf2 scratch must fit in the pipeline buffers");

__global__ void __launch_bounds__(256, 1)
chol_tiles_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
                     const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapYlo,
                     const __grid_constant__ TileArgs g) {
  extern __shared__ uint8_t tsm_raw[];
  uint8_t* sm = smem_align1024(tsm_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TT_DATA_BYTES);
  uint64_t* mma_done = bars;          // [TT_STAGES]
  uint64_t* acc_done = bars + TT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TT_STAGES + 1);
  uint64_t* full = bars + TT_STAGES + 2;   // [TT_STAGES] TMA form: the four planes of a chunk have landed
  __shared__ int s_task[4];
  __shared__ int s_cnt[2];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int R = 32 * (warp & 3) + lane;       // my row of the tile
  const int C0 = 64 * (warp >> 2);            // my first column
  const int nblk = g.nblk;

  if (tid == 0) {
    for (int i = 0; i < TT_STAGES; ++i) { mbar_init(&mma_done[i], 1); mbar_init(&full[i], 1); }
    mbar_init(acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_mine = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)C0;
  uint32_t par_done = 0;   // bit i: parity to wait for on mma_done[i]
  uint32_t par_full = 0;   // bit i: parity to wait for on full[i] (MMA thread of the TMA form)
  uint32_t par_acc = 0;

  for (;;) {
    // ---------------- fetch the next task
    if (tid == 0) {
      const int t = (int)atomicAdd(g.counter, 1u);
      int type = -1, ti = 0, tc = 0;
      if (t < g.ntasks) {
        // Task order: columns in GROUPS of g.gcols; inside a group (columns cg .. ce - 1) first the small triangle of tiles on
        // and right below the diagonal (the panel chain), then row by row the L tiles of ALL the group's columns, then the
        // Y tiles the same way.  Tiles (i, c) and (i, c + 1) stream the same row i of the factor: handed out back to back
        // they run at the same time on two SMs and the second finds the row in L2 (K = 8192: the factor no longer fits L2
        // and every column sweep re-streamed it from HBM).  Every operand is still the result of a lower-numbered task.
        int rem = t;
        const int G = g.gcols;
        for (int cg = g.c0; cg < g.c1; cg += G) {
          const int ce = min(cg + G, g.c1), w = ce - cg;
          const int n1 = (w * (w + 1)) >> 1;                 // (1) triangle, column-major
          if (rem < n1) {
            int j = 0;
            while (rem >= w - j) { rem -= w - j; ++j; }
            type = 0; tc = cg + j; ti = cg + j + rem; break;
          }
          rem -= n1;
          const int n2 = (nblk - ce) * w;                    // (2) rows below the group, row-major
          if (rem < n2) { type = 0; ti = ce + rem / w; tc = cg + rem % w; break; }
          rem -= n2;
          if (g.with_inv) {
            const int n3 = (w * (w - 1)) >> 1;               // (3) Y tiles whose row lies inside the group, column-major
            if (rem < n3) {
              int j = 1;
              while (rem >= j) { rem -= j; ++j; }
              type = 1; tc = cg + j; ti = cg + rem; break;
            }
            rem -= n3;
            const int n4 = (cg - g.c0) * w;                  // (4) Y tiles above the group, row-major
            if (rem < n4) { type = 1; ti = g.c0 + rem / w; tc = cg + rem % w; break; }
            rem -= n4;
          }
        }
      }
      s_task[0] = type; s_task[1] = ti; s_task[2] = tc; s_task[3] = t;
    }
    __syncthreads();
    const int type = s_task[0], ti = s_task[1], c = s_task[2];
    unsigned long long* tr = (g.trace != nullptr && tid == 0 && type >= 0) ? g.trace + (int64_t)s_task[3] * 8 : nullptr;
    __syncthreads();
    if (type < 0) break;
    if (tr) { tr[0] = ((unsigned long long)type << 32) | ((unsigned long long)ti << 16) | (unsigned long long)c; tr[1] = gtime_ns(); }

    const bool isL = type == 0;
    const bool diag = isL && ti == c;
    const int j0 = isL ? g.c0 : ti;
    const int nsteps = c - j0;
    const float* Pbase = (isL ? g.A : g.Y) + (int64_t)ti * NB * g.ld;
    const float* Pbase_lo = (isL ? g.Alo : g.Ylo) + (int64_t)ti * NB * g.ld;
    const uint32_t* Pflag = isL ? g.flagL + (int64_t)ti * nblk : g.flagY + (int64_t)ti * nblk;
    const float* Qbase = g.A + (int64_t)c * NB * g.ld;
    const float* Qbase_lo = g.Alo + (int64_t)c * NB * g.ld;
    const uint32_t* Qflag = g.flagL + (int64_t)c * nblk;
    const int prow_valid = min(NB, g.k - ti * NB);
    const int qrow_valid = min(NB, g.k - c * NB);

    // ---------------- C = A(i, c) (L tasks) or 0 (Y tasks); identity padding of a short last diagonal block
    float creg[64];
    if (isL) {
      const float* At = g.A + ((int64_t)ti * NB + R) * g.ld + (int64_t)c * NB + C0;
      const bool full = R < prow_valid && C0 + 64 <= qrow_valid;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full) {
          v = __ldcg(reinterpret_cast<const float4*>(At + 4 * j));
        } else {
          float e[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = C0 + 4 * j + q;
            e[q] = (R < prow_valid && col < qrow_valid) ? __ldcg(At + 4 * j + q) : ((diag && R == col) ? 1.0f : 0.0f);
          }
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
        creg[4 * j] = v.x; creg[4 * j + 1] = v.y; creg[4 * j + 2] = v.z; creg[4 * j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 64; ++j) creg[j] = 0.0f;
    }

    // ---------------- pipelined accumulation  C -= sum_j P_j Q_j^T   (chunk = 4 * step + quarter)
    // Warps 0..6 are LOADERS (cp.async of the four planes of a chunk, up to two chunks ahead), warp 7 is the MMA warp: it
    // takes a stage over through a named barrier, issues the 12 MMAs (the issue blocks for about their execution time)
    // and commits to the stage's mbarrier, which the loaders wait on before they overwrite that stage.
    const int total = 4 * nsteps;
    auto issue = [&](int gch) {
      const int step = gch >> 2, kofs = (j0 + step) * NB + (gch & 3) * 32;
      uint8_t* st = sm + (gch % TT_STAGES) * TT_STAGE_BYTES;
      tt_load_plane<224>(st, Pbase, g.ld, kofs, prow_valid, tid);
      tt_load_plane<224>(st + 2 * TT_PLANE, Pbase_lo, g.ld, kofs, prow_valid, tid);
      if (!diag) {
        tt_load_plane<224>(st + TT_PLANE, Qbase, g.ld, kofs, qrow_valid, tid);
        tt_load_plane<224>(st + 3 * TT_PLANE, Qbase_lo, g.ld, kofs, qrow_valid, tid);
      }
      cp_async_commit();   // one group per chunk
    };
    auto flags_ready = [&](int gch) -> bool {  // thread 0: may chunk gch be loaded (non-blocking)?
      if ((gch & 3) != 0) return true;
      const int j = j0 + (gch >> 2);
      const uint32_t fp = ld_acquire_u32(Pflag + j);                 // both loads in flight together
      const uint32_t fq = diag ? 1u : ld_acquire_u32(Qflag + j);
      return fp != 0u && fq != 0u;
    };
    auto poll_block = [&](int gch) {            // thread 0: wait until chunk gch may be loaded
      if ((gch & 3) == 0) {
        const int j = j0 + (gch >> 2);
        wait_flag(Pflag + j);
        if (!diag) wait_flag(Qflag + j);
      }
    };
    auto drain = [&]() {  // creg -= TMEM accumulator (all MMAs issued so far are complete once acc_done flips)
      mbar_wait(acc_done, par_acc);
      par_acc ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;");
      float v[32];
      tmem_ld32(tmem_mine, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) creg[j] -= v[j];
      tmem_ld32(tmem_mine + 32u, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) creg[32 + j] -= v[j];
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();   // all eight warps: nobody starts the next accumulation chain while TMEM is still being read
    };
    if (total > 0 && g.use_tma) {
      // ---- TMA form (default): thread 0 is the producer -- one cp.async.bulk.tensor per plane of a chunk (the SWIZZLE_128B box
      // IS the layout tt_load_plane builds), armed on the stage's `full` mbarrier -- thread 224 issues the MMAs, and the other
      // 254 threads only meet them at the drains.  Same look-ahead rules as the cp.async form below (two chunks ahead, flags
      // checked without blocking unless the pipeline is dry), without its 224-thread issue loops and named barriers per chunk.
      const CUtensorMap* mPh = isL ? &mapA : &mapY;
      const CUtensorMap* mPl = isL ? &mapAlo : &mapYlo;
      auto issue_tma = [&](int gch) {
        const int s = gch % TT_STAGES;
        const int kofs = (j0 + (gch >> 2)) * NB + (gch & 3) * 32;
        uint8_t* st = sm + s * TT_STAGE_BYTES;
        if ((gch & 3) == 0) asm volatile("fence.proxy.async.global;" ::: "memory");   // tiles acquired through flags -> async proxy
        mbar_expect_tx(&full[s], (uint32_t)((diag ? 2 : 4) * TT_PLANE));
        tma_load_2d(mPh, &full[s], st, kofs, ti * NB);
        tma_load_2d(mPl, &full[s], st + 2 * TT_PLANE, kofs, ti * NB);
        if (!diag) {
          tma_load_2d(&mapA, &full[s], st + TT_PLANE, kofs, c * NB);
          tma_load_2d(&mapAlo, &full[s], st + 3 * TT_PLANE, kofs, c * NB);
        }
      };
      int issued = 0, chain = 0;
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the previous task used this memory through the generic proxy
        poll_block(0);
        issue_tma(0);
        issued = 1;
        if (total > 1) { issue_tma(1); issued = 2; }
      }
      for (int it = 0; it < total; ++it) {
        const int s = it % TT_STAGES;
        const bool do_drain = it == total - 1 || chain + 1 == TT_DRAIN;
        if (tid == 0) {
          if (issued <= it) {                                 // pipeline dry: chunk `it` opens a step whose tiles were not ready
            poll_block(it);
            issue_tma(it);
            issued = it + 1;
          }
          if (it >= 1) {                                      // the stage of chunk it+2 is the stage of chunk it-1
            const int sp = (it - 1) % TT_STAGES;
            mbar_wait(&mma_done[sp], (par_done >> sp) & 1u);
            par_done ^= 1u << sp;
          }
          while (issued < total && issued <= it + 2) {
            if (!flags_ready(issued)) break;
            issue_tma(issued);
            ++issued;
          }
        } else if (tid == 224) {
          mbar_wait(&full[s], (par_full >> s) & 1u);
          par_full ^= 1u << s;
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t ph = smem_u32(sm + s * TT_STAGE_BYTES), pl = ph + 2 * TT_PLANE;
          const uint32_t qh = diag ? ph : ph + TT_PLANE, ql = diag ? pl : pl + TT_PLANE;
          tt_mma_chunk(tmem_base, ph, pl, qh, ql, chain == 0);
          umma_commit(&mma_done[s]);
          if (do_drain) umma_commit(acc_done);
        }
        if (do_drain) { drain(); chain = 0; } else { ++chain; }
      }
      if (tid == 0) {  // MMAs of the last chunk: complete (acc_done waited), consume its stage barrier phase
        const int sp = (total - 1) % TT_STAGES;
        mbar_wait(&mma_done[sp], (par_done >> sp) & 1u);
        par_done ^= 1u << sp;
      }
      __syncthreads();
    } else if (total > 0) {
      if (warp < 7) {
        // ---- loaders.  A step's tiles may not be flagged yet: thread 0 checks WITHOUT blocking while chunks are still
        // in flight and blocks only when the pipeline has run dry.
        if (tid == 0) poll_block(0);
        nbar_sync(TT_BAR_WORK, 224);
        issue(0);
        issue(1);
        int issued = 2;
        int chain = 0;
        for (int it = 0; it < total; ++it) {
          const int s = it % TT_STAGES;
          if (issued <= it) {                               // pipeline dry: chunk `it` opens a step whose tiles were not ready
            if (tid == 0) poll_block(it);
            nbar_sync(TT_BAR_WORK, 224);
            issue(it);
            issued = it + 1;
          }
          const int pend = issued - it - 1;                 // younger groups that may stay in flight
          if (pend >= 2) cp_async_wait<2>();
          else if (pend == 1) cp_async_wait<1>();
          else cp_async_wait<0>();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my landed pieces -> visible to the tensor core
          nbar_arrive(TT_BAR_FULL0 + s, 256);
          if (tid == 0) {                                   // how many of the chunks up to it+2 can be loaded now?
            int cnt = 0;
            for (int gch = issued; gch <= it + 2 && gch < total; ++gch) {
              if (!flags_ready(gch)) break;
              ++cnt;
            }
            s_cnt[it & 1] = cnt;
          }
          nbar_sync(TT_BAR_WORK, 224);
          const int can_issue = s_cnt[it & 1];
          const bool do_drain = it == total - 1 || chain + 1 == TT_DRAIN;
          if (it >= 1) {                                    // the stage of chunk it+2 is the stage of chunk it-1
            const int sp = (it - 1) % TT_STAGES;
            mbar_wait(&mma_done[sp], (par_done >> sp) & 1u);
            par_done ^= 1u << sp;
          }
          for (int q = 0; q < can_issue; ++q) issue(issued++);
          if (do_drain) { drain(); chain = 0; } else { ++chain; }
        }
        {  // MMAs of the last chunk: complete (acc_done waited), consume its stage barrier phase
          const int sp = (total - 1) % TT_STAGES;
          mbar_wait(&mma_done[sp], (par_done >> sp) & 1u);
          par_done ^= 1u << sp;
        }
        cp_async_wait<0>();
      } else {
        // ---- MMA warp
        int chain = 0;
        for (int it = 0; it < total; ++it) {
          const int s = it % TT_STAGES;
          nbar_sync(TT_BAR_FULL0 + s, 256);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const bool do_drain = it == total - 1 || chain + 1 == TT_DRAIN;
          if (lane == 0) {
            const uint32_t ph = smem_u32(sm + s * TT_STAGE_BYTES), pl = ph + 2 * TT_PLANE;
            const uint32_t qh = diag ? ph : ph + TT_PLANE, ql = diag ? pl : pl + TT_PLANE;
            tt_mma_chunk(tmem_base, ph, pl, qh, ql, chain == 0);
            umma_commit(&mma_done[s]);
            if (do_drain) umma_commit(acc_done);
          }
          __syncwarp();
          par_done ^= 1u << s;                              // bookkeeping: one phase of mma_done[s] per commit
          if (do_drain) { drain(); chain = 0; } else { ++chain; }
        }
      }
      __syncthreads();
    }
    if (tr) tr[2] = gtime_ns();

    if (diag) {
      // ---------------- potf2 of the diagonal tile fused with the inverse of its factor (fp32 SIMT, shared memory)
      float* fsm = reinterpret_cast<float*>(sm);
      float* Tt = fsm;
      float* Ww = Tt + FI_T_FLOATS;
      float* Pn = Ww + FI_T_FLOATS;
      float* Wp = Pn + (g.fi_nt > 0 ? FS_PN_FLOATS : 2 * 8 * NB);
      float* Dd = Wp + 2 * 8 * NB;
      {
        const int tir = R >> 2;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int q0 = C0 + 4 * j, tk = q0 >> 2;
          if (tk <= tir) {
            float4 v;
            v.x = (q0 + 0 <= R) ? creg[4 * j + 0] : 0.0f;
            v.y = (q0 + 1 <= R) ? creg[4 * j + 1] : 0.0f;
            v.z = (q0 + 2 <= R) ? creg[4 * j + 2] : 0.0f;
            v.w = (q0 + 3 <= R) ? creg[4 * j + 3] : 0.0f;
            *reinterpret_cast<float4*>(fi_row(Tt, tir, tk, R & 3)) = v;
          }
        }
      }
      __syncthreads();
      bool bad = false;
      if (g.fi_nt > 0)
        factor_invert_split(Tt, Ww, Pn, Wp, Dd, reinterpret_cast<int*>(Dd + 16 * 64), bad, tid, g.fi_nt,
                            (g.trace != nullptr && c == 1) ? reinterpret_cast<long long*>(g.trace + (int64_t)g.ntasks * 8) : nullptr);
      else
        factor_invert_la(Tt, Ww, Pn, Wp, Dd, bad, tid,
                         (g.trace != nullptr && c == 1) ? reinterpret_cast<long long*>(g.trace + (int64_t)g.ntasks * 8) : nullptr);
      if (bad && g.status) atomicOr(g.status, LCB_ST_NOT_SPD);
      if (tr) tr[4] = gtime_ns();
      {  // Linv(c) -> dinv[c] hi / lo (row-major), a warp per row: coalesced 512-byte stores
        float* dv = g.dinv + (int64_t)c * NB * NB;
        float* dvl = g.dinv_lo + (int64_t)c * NB * NB;
        for (int idx = tid; idx < NB * 32; idx += 256) {
          const int r = idx >> 5, tk = idx & 31, tir = r >> 2;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (tk <= tir) v = *reinterpret_cast<const float4*>(fi_row(Ww, tir, tk, r & 3));
          float4 h, l;
          tt_split4(v, h, l);
          *reinterpret_cast<float4*>(dv + r * NB + 4 * tk) = h;
          *reinterpret_cast<float4*>(dvl + r * NB + 4 * tk) = l;
        }
      }
      __syncthreads();
      if (tid == 0) { __threadfence(); st_release_u32(g.flagL + (int64_t)c * nblk + c, 1u); }
      if (tr) tr[5] = gtime_ns();
      if (g.with_inv) {  // Y(c, c) = Linv(c)^T
        float* Yt = g.Y + (int64_t)c * NB * g.ld + (int64_t)c * NB;
        float* Ytl = g.Ylo + (int64_t)c * NB * g.ld + (int64_t)c * NB;
        const int r = tid & 127, ch = tid >> 7, tir = r >> 2;
        if (r < qrow_valid) {
#pragma unroll 4
          for (int q = 0; q < 64; ++q) {
            const int cc = 64 * ch + q;
            if (cc < qrow_valid) {
              const float v = ((cc >> 2) <= tir) ? fi_row(Ww, tir, cc >> 2, r & 3)[cc & 3] : 0.0f;
              uint32_t hb;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
              const float h = __uint_as_float(hb);
              Yt[(int64_t)cc * g.ld + r] = h;
              Ytl[(int64_t)cc * g.ld + r] = __fsub_rn(v, h);
            }
          }
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release_u32(g.flagY + (int64_t)c * nblk + c, 1u); }
      }
      __syncthreads();
      continue;
    }

    // ---------------- out = C Linv(c)^T : C hi / lo planes (A operand, 128 KB) + the lower triangle of Linv hi / lo (80 KB:
    // chunk ch of the reduction only meets output columns >= 32 ch, so only those rows of Linv are staged and the MMAs of
    // chunk ch run with N = 128 - 32 ch)
    {
      uint8_t* Ch = sm;                         // 4 chunks hi (64 KB)
      uint8_t* Cl = sm + 4 * TT_PLANE;          // 4 chunks lo (64 KB)
      uint8_t* Xh = sm + 8 * TT_PLANE;          // Linv hi: chunk ch at byte 128 * (128 ch - 16 ch (ch - 1)), rows >= 32 ch
      uint8_t* Xl = Xh + TT_XPLANES;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = C0 + 4 * j;
        const int ch = col >> 5, kq = (col & 31) >> 2;
        const uint32_t off = (uint32_t)(ch * TT_PLANE + R * 128 + ((kq ^ (R & 7)) << 4));
        float4 h, l;
        tt_split4(make_float4(creg[4 * j], creg[4 * j + 1], creg[4 * j + 2], creg[4 * j + 3]), h, l);
        *reinterpret_cast<float4*>(Ch + off) = h;
        *reinterpret_cast<float4*>(Cl + off) = l;
      }
      if (tid == 0) wait_flag(g.flagL + (int64_t)c * nblk + c);
      if (tr) tr[3] = gtime_ns();
      __syncthreads();
      const float* dv = g.dinv + (int64_t)c * NB * NB;
      const float* dvl = g.dinv_lo + (int64_t)c * NB * NB;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int xo = 128 * (128 * ch - 16 * ch * (ch - 1));   // 0, 16 K, 28 K, 36 K
        tt_load_plane<256>(Xh + xo, dv, NB, ch * 32, NB, tid, 32 * ch);
        tt_load_plane<256>(Xl + xo, dvl, NB, ch * 32, NB, tid, 32 * ch);
      }
      cp_async_commit();
      cp_async_wait<0>();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int xo = 128 * (128 * ch - 16 * ch * (ch - 1));
          const uint32_t ah = smem_u32(Ch + ch * TT_PLANE), al = smem_u32(Cl + ch * TT_PLANE);
          const uint32_t bh = smem_u32(Xh + xo), bl = smem_u32(Xl + xo);
          const uint32_t n = 128u - 32u * ch, td = tmem_base + 32u * ch;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ko = k * 32;
            umma_tf32_n(td, tt_desc(al + ko), tt_desc(bh + ko), n, (ch == 0 && k == 0) ? 0u : 1u);
            umma_tf32_n(td, tt_desc(ah + ko), tt_desc(bl + ko), n, 1u);
            umma_tf32_n(td, tt_desc(ah + ko), tt_desc(bh + ko), n, 1u);
          }
        }
        umma_commit(acc_done);
      }
      mbar_wait(acc_done, par_acc);
      par_acc ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;");
      // result -> shared memory (row stride 132 floats: conflict-free for one row per lane) -> coalesced hi / lo stores
      float* Sh = reinterpret_cast<float*>(sm);
      float* Sl = Sh + NB * 132;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tmem_ld32(tmem_mine + (uint32_t)(32 * hh), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 h, l;
          tt_split4(make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]), h, l);
          *reinterpret_cast<float4*>(Sh + R * 132 + C0 + 32 * hh + 4 * j) = h;
          *reinterpret_cast<float4*>(Sl + R * 132 + C0 + 32 * hh + 4 * j) = l;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      const int64_t ooff = (int64_t)ti * NB * g.ld + (int64_t)c * NB;
      float* Ot = (isL ? g.A : g.Y) + ooff;
      float* Otl = (isL ? g.Alo : g.Ylo) + ooff;
      for (int idx = tid; idx < NB * 32; idx += 256) {
        const int r = idx >> 5, q4 = idx & 31;
        if (r < prow_valid) {
          const float4 h = *reinterpret_cast<const float4*>(Sh + r * 132 + 4 * q4);
          const float4 l = *reinterpret_cast<const float4*>(Sl + r * 132 + 4 * q4);
          if (4 * q4 + 3 < qrow_valid) {
            *reinterpret_cast<float4*>(Ot + (int64_t)r * g.ld + 4 * q4) = h;
            *reinterpret_cast<float4*>(Otl + (int64_t)r * g.ld + 4 * q4) = l;
          } else {
            const float hv[4] = {h.x, h.y, h.z, h.w}, lv[4] = {l.x, l.y, l.z, l.w};
            for (int e = 0; e < 4; ++e)
              if (4 * q4 + e < qrow_valid) {
                Ot[(int64_t)r * g.ld + 4 * q4 + e] = hv[e];
                Otl[(int64_t)r * g.ld + 4 * q4 + e] = lv[e];
              }
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      st_release_u32((isL ? g.flagL : g.flagY) + (int64_t)ti * nblk + c, 1u);
    }
    if (tr) tr[5] = gtime_ns();
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
  }
}

// U[a][b] = (b >= a) ? Y[k-1-b][k-1-a] : 0     (Y = L^-T of the index-reversed matrix; U = J L^-1 J)
__global__ void reverse_out_t_kernel(float* __restrict__ U, const float* __restrict__ Y, const float* __restrict__ Ylo,
                                     int64_t k) {
  __shared__ float tile[32][33];
  const int64_t a0 = (int64_t)blockIdx.y * 32, b0 = (int64_t)blockIdx.x * 32;
  // U rows a0.., cols b0..  <-  Y rows k-1-b, cols k-1-a
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t b = b0 + r, a = a0 + threadIdx.x;       // read Y[k-1-b][k-1-a]: contiguous in a (descending)
    float v = 0.0f;
    if (a < k && b < k && b >= a) v = Y[(k - 1 - b) * k + (k - 1 - a)] + Ylo[(k - 1 - b) * k + (k - 1 - a)];
    tile[r][threadIdx.x] = v;                             // tile[b - b0][a - a0]
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t a = a0 + r, b = b0 + threadIdx.x;
    if (a < k && b < k) U[a * k + b] = tile[threadIdx.x][r];
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
    int prio_least = 0, prio_greatest = 0;
    ok = ok && cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&c.chain, cudaStreamNonBlocking, prio_greatest) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c.evIn, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c.evOut, cudaEventDisableTiming) == cudaSuccess;
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

// [k, k] fp32 matrix, boxes of 32 columns x 128 rows, SWIZZLE_128B: one operand chunk of the tile tasks
static int make_tile_map(CUtensorMap* map, const float* base, int64_t k) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc != LCB_OK) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)k};
  cuuint64_t strides[1] = {(cuuint64_t)k * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)NB};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (tile tasks, k = %lld) failed with CUresult %d", (long long)k, (int)r);
    return LCB_ERR_CUDA;
  }
  return LCB_OK;
}

constexpr int TILES_FULL_MAX = 16384;  // largest K factored by one chol_tiles_tc_kernel launch (above: the panel-launch chain)
static int tiles_mode() {             // LCB_CHOL_TILES=0: panel-launch chain (A/B runs)
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("LCB_CHOL_TILES");
    m = (e && e[0] == '0') ? 0 : 1;
  }
  return m;
}
constexpr int SW = 512;        // super-panel width of the tensor-core path (Kd of the big SYRK)
constexpr int TRI_TG_MIN = 1024;  // trtri levels with node size >= this run on the tensor cores
constexpr int TG_CHAIN = 256;     // accumulation chain length (columns) of the tensor-core contractions

// tile-task path: A, Y, lo(A), lo(Y) [k, k] each; dinv hi / lo; flags + task counter; debug trace; diag sum
static size_t tiles_flags_offset(int64_t k) { return (size_t)(4 * k * k) + (size_t)(2 * ceil_div(k, NB) * NB * NB); }
static size_t tiles_trace_offset(int64_t k) {
  const int64_t nblk = ceil_div(k, NB);
  return tiles_flags_offset(k) + (size_t)((2 * nblk * nblk + 16 + 3) / 4 * 4);
}
constexpr size_t TILES_TRACE_EXTRA = 16 * 12 * 2 + 40 * 8 * 2;   // floats: the clock probes behind the task records
static size_t tiles_ws_floats(int64_t k) {
  const int64_t nblk = ceil_div(k, NB);
  return tiles_trace_offset(k) + (size_t)(nblk * nblk * 16) + TILES_TRACE_EXTRA + 64;
}

static size_t chol_ws_floats(int64_t k) {
  const int64_t nblk = ceil_div(k, NB);
  const int64_t kp = ceil_div(k, 4) * 4;
  const size_t t_exact = (size_t)(k * k / 2 + k * NB);
  const size_t t_tg = (size_t)(9 * (kp / 2 + NB) * (kp / 2 + NB));  // planes of one trtri node
  const size_t chain = (size_t)(k * k)   // work matrix
         + std::max(t_exact, t_tg)       // T of the trtri recursion / operand planes
         + (size_t)(nblk * NB * NB)      // inverted diagonal blocks
         + (size_t)(4 * kp * NB)         // TRSM panel (exact path) / two sets of panel hi + lo planes
         + (size_t)(4 * kp * SW)         // two sets of hi / lo planes of a super-panel strip
         + 64;
  return std::max(chain, tiles_ws_floats(k));
}

}  // namespace lcb

using namespace lcb;

extern "C" size_t lcb_chol_ws_bytes(int64_t k) { return chol_ws_floats(k) * sizeof(float); }

// debug aid (LCB_CHOL_TRACE=1): byte offset of the task trace inside the workspace (8 x u64 per task, nblk^2 tasks)
extern "C" size_t lcb_chol_trace_offset(int64_t k) { return tiles_trace_offset(k) * sizeof(float); }

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
  const bool tiles = tg && k <= TILES_FULL_MAX && tiles_mode() != 0;
  if (tiles) dsum = static_cast<float*>(ws) + tiles_ws_floats(k) - 64;
  if (tg && !tiles) LCB_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM));

  diag_sum_kernel<<<1, 1024, 0, st>>>(H, k, dsum);
  LCB_LAUNCH_CHECK();
  dim3 g2((unsigned)ceil_div(k, 256), (unsigned)k);
  gather_reverse_damp_kernel<<<g2, 256, 0, st>>>(A, H, perm, k, damp, dsum);
  LCB_LAUNCH_CHECK();

  if (tiles) {
    // ---- whole factorisation + triangular inverse as ONE persistent tile-task launch
    float* w0 = static_cast<float*>(ws);
    uint32_t* flags = reinterpret_cast<uint32_t*>(w0 + tiles_flags_offset(k));
    const size_t nfl = (size_t)(2 * nblk * nblk + 16);
    LCB_CUDA(cudaMemsetAsync(flags, 0, nfl * sizeof(uint32_t), st));
    TileArgs ta{};
    ta.A = A; ta.Y = w0 + k * k; ta.Alo = w0 + 2 * k * k; ta.Ylo = w0 + 3 * k * k;
    ta.dinv = w0 + 4 * k * k; ta.dinv_lo = ta.dinv + nblk * NB * NB;
    ta.flagL = flags; ta.flagY = flags + nblk * nblk; ta.counter = flags + 2 * nblk * nblk;
    ta.status = status; ta.ld = k; ta.k = (int)k; ta.nblk = (int)nblk; ta.c0 = 0; ta.c1 = (int)nblk; ta.with_inv = 1;
    ta.ntasks = (int)(nblk * nblk);
    {  // LCB_CHOL_FI=0: joint-worker potf2 (round-2 first form); 1..6: T-worker warps of the split form (default 4)
      static int fi = -1;
      if (fi < 0) {
        const char* e = getenv("LCB_CHOL_FI");
        fi = (e && e[0] >= '0' && e[0] <= '6') ? (e[0] - '0') : 4;
      }
      ta.fi_nt = fi;
    }
    {  // LCB_CHOL_GROUP=1..4: columns per task group.  Default 2 where a group's tasks still fit the grid at once (a larger
       // group would stall the panel chain at every group boundary until enough of its tasks have retired)
      static int gc = -1;
      if (gc < 0) {
        const char* e = getenv("LCB_CHOL_GROUP");
        gc = (e && e[0] >= '1' && e[0] <= '4') ? (e[0] - '0') : 0;
      }
      ta.gcols = gc > 0 ? gc : ((int)nblk * 2 <= sm_count() ? 2 : 1);
    }
    if (getenv("LCB_CHOL_TRACE")) {  // debug: task trace at byte offset lcb_chol_trace_offset(k) of the workspace
      ta.trace = reinterpret_cast<unsigned long long*>(w0 + tiles_trace_offset(k));
      LCB_CUDA(cudaMemsetAsync(ta.trace, 0, ((size_t)ta.ntasks * 16 + TILES_TRACE_EXTRA) * sizeof(float), st));
    }
    {  // LCB_CHOL_TMA=0: cp.async operand staging (A/B runs)
      static int tm = -1;
      if (tm < 0) {
        const char* e = getenv("LCB_CHOL_TMA");
        tm = (e && e[0] == '0') ? 0 : 1;
      }
      ta.use_tma = tm;
    }
    CUtensorMap mA, mAlo, mY, mYlo;
    {
      int mrc;
      if ((mrc = make_tile_map(&mA, ta.A, k)) != LCB_OK || (mrc = make_tile_map(&mAlo, ta.Alo, k)) != LCB_OK ||
          (mrc = make_tile_map(&mY, ta.Y, k)) != LCB_OK || (mrc = make_tile_map(&mYlo, ta.Ylo, k)) != LCB_OK)
        return mrc;
    }
    const unsigned grid = (unsigned)std::min<int64_t>(ta.ntasks, sm_count());
    LCB_CUDA(cudaFuncSetAttribute(chol_tiles_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM));
    chol_tiles_tc_kernel<<<grid, 256, TT_SMEM, st>>>(mA, mAlo, mY, mYlo, ta);
    LCB_LAUNCH_CHECK();
    dim3 gr((unsigned)ceil_div(k, 32), (unsigned)ceil_div(k, 32));
    reverse_out_t_kernel<<<gr, dim3(32, 8), 0, st>>>(U, ta.Y, ta.Ylo, k);
    LCB_LAUNCH_CHECK();
    return LCB_OK;
  }

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

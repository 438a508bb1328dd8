// (b) Column-block quantize-and-error-propagate loops of GPTQ / GPTAQ / SparseGPT.
//
// ref: gptq/core.py:226-265, gptaq/core.py:274-319, sparsegpt/core.py:192-218.
// One kernel launch per 128-column block does everything the reference does in its inner Python
// loop (128 iterations x ~25 tiny kernels): quantise / mask the column (or the group), form the
// error, and propagate it to the later columns of the block.  Rows are independent given U, so
// a CTA owns 32 rows and never synchronises with other CTAs; 8 threads share a row (16 columns
// each, held in registers), the 128x128 diagonal block of U (and P for GPTAQ) sits in shared
// memory.  The trailing lazy-batch update W[:, i2:] -= Err @ U[i1:i2, i2:] is an fp32 GEMM.
#include <atomic>
#include <cstdlib>

#include "linalg.cuh"
#include "qmath.cuh"

namespace lcb {

namespace {

constexpr int BLK = 128;           // column block (the reference's block_size)
constexpr int CPT = 16;            // columns per thread
constexpr int TPR = BLK / CPT;     // threads per row = 8
constexpr int ROWS_PER_CTA = 32;   // 256 threads
constexpr int ULD = BLK + 4;

constexpr int SUPER = 1024;        // second-level lazy batch of the tensor-core path (columns)

enum { MODE_QUANT = 0, MODE_SPARSE = 1 };

__device__ __forceinline__ float tf32_hi(float v) {
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
  return __uint_as_float(hb);
}

struct BlockArgs {
  float* W;           // [n, k] working weight (block columns are read; SPARSE: written back)
  float* Q;           // [n, k] output (QUANT)
  const float* U;     // [k, k]
  const float* P;     // [k, k] or null
  const float* scales;  // [n, G]
  const float* zeros;
  const uint8_t* keep;   // [n, k] or null
  const uint8_t* prune;  // [n, BLK] (SPARSE) mask for this block
  float* Err;         // [n, ldE] error columns of this block at column offset eoff (hi plane when ErrLo != null)
  float* ErrLo;       // tf32 split: lo plane (Err = hi + lo), or null for one exact fp32 plane
  float* W1out;       // (GPTAQ) block after in-block updates, same geometry as Err, or null
  float* W1Lo;
  int64_t ldE, eoff;
  int64_t n, k;
  int64_t i1;         // first column of the block
  int count;          // columns in the block (<= BLK)
  int64_t group;      // > 0 grouped, <= 0 per column
  int64_t G;          // parameter columns per row
  QCfg c;
};

template <int MODE, bool HAS_P>
__global__ void __launch_bounds__(256) block_step_kernel(BlockArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Us = smem;                               // [BLK][ULD]
  float* Ps = HAS_P ? smem + BLK * ULD : nullptr;  // [BLK][ULD]
  const int t = threadIdx.x;
  const bool grouped = (MODE == MODE_QUANT) && a.group > 0;
  const int gs = grouped ? (int)a.group : 1;
  // With one group spanning the whole block and no P nothing downstream reads the in-block
  // propagation, so only diag(U) is needed.
  const bool need_prop = HAS_P || !grouped || gs < a.count;
  if (need_prop) {
    // stage the diagonal blocks (zero filled beyond `count`)
    for (int e = t; e < BLK * BLK; e += 256) {
      const int r = e / BLK, c = e % BLK;
      const bool in = r < a.count && c < a.count;
      Us[r * ULD + c] = in ? a.U[(a.i1 + r) * a.k + a.i1 + c] : 0.0f;
      if (HAS_P) Ps[r * ULD + c] = in ? a.P[(a.i1 + r) * a.k + a.i1 + c] : 0.0f;
    }
  } else if (t < BLK) {
    Us[t * ULD + t] = t < a.count ? a.U[(a.i1 + t) * a.k + a.i1 + t] : 0.0f;
  }
  __syncthreads();

  const int t8 = t % TPR;
  const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + t / TPR;
  const bool row_ok = row < a.n;
  const int64_t rr = row_ok ? row : 0;
  const int c0 = t8 * CPT;

  float w[CPT], snap[CPT], qv[CPT], ev[CPT];
  const float* wp = a.W + rr * a.k + a.i1 + c0;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    w[j] = (row_ok && c0 + j < a.count) ? wp[j] : 0.0f;
    snap[j] = w[j]; qv[j] = 0.0f; ev[j] = 0.0f;
  }
  uint32_t keepbits = 0xffffu, prunebits = 0u;
  if (MODE == MODE_QUANT && a.keep && row_ok) {
    keepbits = 0;
#pragma unroll
    for (int j = 0; j < CPT; ++j)
      if (c0 + j < a.count && a.keep[rr * a.k + a.i1 + c0 + j]) keepbits |= 1u << j;
  }
  if (MODE == MODE_SPARSE && row_ok) {
#pragma unroll
    for (int j = 0; j < CPT; ++j)
      if (c0 + j < a.count && a.prune[rr * BLK + c0 + j]) prunebits |= 1u << j;
  }
  float s_cur = 1.0f, z_cur = 0.0f;
  if (MODE == MODE_QUANT && !grouped) {  // per-row parameters
    s_cur = a.scales[rr * a.G];
    z_cur = a.zeros[rr * a.G];
  }
  const int lane = t & 31;
  const int lane_row_base = lane - t8;  // lane of t8 == 0 for this row

#pragma unroll 1
  for (int ib = 0; ib < TPR; ++ib) {
    if (ib * CPT >= a.count) break;
    if (grouped && ((ib * CPT) % gs) == 0) {
      // group entry: columns are quantised from their value at this point (ref :253-262)
#pragma unroll
      for (int j = 0; j < CPT; ++j) snap[j] = w[j];
      const int64_t g = (a.i1 + ib * CPT) / gs;
      s_cur = a.scales[rr * a.G + g];
      z_cur = a.zeros[rr * a.G + g];
    }
#pragma unroll
    for (int ii = 0; ii < CPT; ++ii) {
      const int i = ib * CPT + ii;
      // value of column i as the quantiser sees it, broadcast from its owner thread
      const float mine = grouped ? snap[ii] : w[ii];
      const float wi = __shfl_sync(0xffffffffu, mine, lane_row_base + ib);
      const uint32_t kb = __shfl_sync(0xffffffffu, MODE == MODE_QUANT ? keepbits : prunebits, lane_row_base + ib);
      float q;
      if (MODE == MODE_QUANT) {
        float code;
        q = fake_quant<LCB_F32>(a.c, wi, s_cur, z_cur, code);
        q = ((kb >> ii) & 1u) ? q : 0.0f;  // q *= MASK (ref :244,258)
      } else {
        q = ((kb >> ii) & 1u) ? 0.0f : wi;  // q[MASK1[:, i]] = 0 (ref sparsegpt :209-210)
      }
      const float d = Us[i * ULD + i];
      const float e = __fdiv_rn(__fsub_rn(wi, q), d);
      if (t8 == ib) { qv[ii] = q; ev[ii] = e; }
      // propagate to columns j >= i of the block: w_j -= e * U[i][j] (- w_i * P[i][j]).
      if (need_prop && i < a.count && t8 >= ib) {
        const float* urow = Us + i * ULD + c0;
        const float* prow = HAS_P ? Ps + i * ULD + c0 : nullptr;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          if (t8 > ib || j >= ii) {
            float upd = __fmul_rn(e, urow[j]);
            if (HAS_P) upd = __fsub_rn(upd, __fmul_rn(wi, prow[j]));
            w[j] = __fsub_rn(w[j], upd);
          }
        }
      }
    }
  }

  if (!row_ok) return;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    if (c0 + j < a.count) {
      if (MODE == MODE_QUANT) a.Q[rr * a.k + a.i1 + c0 + j] = qv[j];
      else a.W[rr * a.k + a.i1 + c0 + j] = qv[j];
    }
    const int64_t eo = rr * a.ldE + a.eoff + c0 + j;
    const float e = (c0 + j < a.count) ? ev[j] : 0.0f;
    if (a.ErrLo) {
      const float h = tf32_hi(e);
      a.Err[eo] = h;
      a.ErrLo[eo] = __fsub_rn(e, h);
    } else {
      a.Err[eo] = e;
    }
    if (a.W1out) {
      const float w1 = (c0 + j < a.count) ? w[j] : 0.0f;
      if (a.W1Lo) {
        const float h = tf32_hi(w1);
        a.W1out[eo] = h;
        a.W1Lo[eo] = __fsub_rn(w1, h);
      } else {
        a.W1out[eo] = w1;
      }
    }
  }
}

// Grouped quantiser whose group spans the whole block (int4-g[128], the headline configuration), no P:
// nothing inside the block depends on the in-block propagation (ref: gptq/core.py:249-262 with
// group_size == block_size), so the block step is elementwise: q = QDQ(w) * MASK, e = (w - q) / diag(U).
__global__ void __launch_bounds__(256) block_quant_kernel(BlockArgs a) {
  // programmatic dependent launch (LCB_UPDATE_PDL; no-ops otherwise): let the GEMM that follows start its prologue, then wait
  // until the kernel before this one (the GEMM that updated this block's columns) has completed and flushed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 of the [n, BLK] block
  const int64_t row = idx / (BLK / 4);
  const int c0 = (int)(idx % (BLK / 4)) * 4;
  if (row >= a.n) return;
  const int64_t g = a.i1 / a.group;
  const float s = a.scales[row * a.G + g], z = a.zeros[row * a.G + g];
  float w[4] = {0.f, 0.f, 0.f, 0.f}, q[4], e[4];
  const int64_t wo = row * a.k + a.i1 + c0;
  if (c0 + 3 < a.count) {
    const float4 v = *reinterpret_cast<const float4*>(a.W + wo);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (c0 + j < a.count) w[j] = a.W[wo + j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool in = c0 + j < a.count;
    float code;
    float qq = fake_quant<LCB_F32>(a.c, w[j], s, z, code);
    if (a.keep && in && !a.keep[wo + j]) qq = 0.0f;  // q *= MASK (ref :258)
    const float d = in ? a.U[(a.i1 + c0 + j) * (a.k + 1)] : 1.0f;
    q[j] = qq;
    e[j] = in ? __fdiv_rn(__fsub_rn(w[j], qq), d) : 0.0f;
  }
  if (c0 + 3 < a.count) {
    *reinterpret_cast<float4*>(a.Q + wo) = make_float4(q[0], q[1], q[2], q[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (c0 + j < a.count) a.Q[wo + j] = q[j];
  }
  const int64_t eo = row * a.ldE + a.eoff + c0;
  if (a.ErrLo) {
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { h[j] = tf32_hi(e[j]); l[j] = __fsub_rn(e[j], h[j]); }
    *reinterpret_cast<float4*>(a.Err + eo) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(a.ErrLo + eo) = make_float4(l[0], l[1], l[2], l[3]);
  } else {
    *reinterpret_cast<float4*>(a.Err + eo) = make_float4(e[0], e[1], e[2], e[3]);
  }
}

// SparseGPT saliency of one block: tmp = W1^2 / diag(U1)^2  (ref: sparsegpt/core.py:201)
__global__ void sparse_metric_kernel(const float* W, const float* U, float* metric, int64_t n, int64_t k, int64_t i1,
                                     int count) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * count) return;
  const int64_t r = idx / count;
  const int c = (int)(idx % count);
  const float w = W[r * k + i1 + c];
  const float d = U[(i1 + c) * k + i1 + c];
  metric[idx] = __fdiv_rn(__fmul_rn(w, w), __fmul_rn(d, d));
}

__global__ void sparse_mask_kernel(const float* metric, const float* thresh, uint8_t* prune, int64_t n, int count) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * count) return;
  const int64_t r = idx / count;
  const int c = (int)(idx % count);
  prune[r * BLK + c] = metric[idx] <= thresh[0];
}

}  // namespace

// exact k-th smallest (0-based index `kth`) of n non-negative-or-any fp32 values (masks.cu)
int select_kth_f32(const float* vals, int64_t n, int64_t kth, float* out, void* ws, size_t ws_bytes, cudaStream_t st,
                   lcb_reduce_u32_fn reduce, void* reduce_user);
size_t select_ws_bytes();

}  // namespace lcb

using namespace lcb;

// Trailing updates on the tensor-core path use a two-level lazy batch: after each 128-column block
// only the rest of the current SUPER-column super-block is updated (Kd = 128); the columns beyond it
// receive one GEMM with Kd = SUPER per super-block, which cuts the read-modify-write traffic on W by
// SUPER / 128 relative to the reference's schedule (same sums, different fp32 association).
static bool use_tg(const float* W, const float* U, const float* P, int64_t k) {
  return gemm_mode() == 1 && tg_ok(W, k) && tg_ok(U, k) && (P == nullptr || tg_ok(P, k));
}
static bool chain_prio_mode() {
  static std::atomic<int> g{-1};
  int m = g.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = std::getenv("LCB_CHAIN_PRIO");
    m = e ? (std::atoi(e) != 0 ? 1 : 0) : 0;  // measured on B200: no gain (block loops 0.57-2.1 ms either way) -> off
    g.store(m, std::memory_order_relaxed);
  }
  return m != 0;
}
static int b_streams_mode() {   // LCB_UPDATE_BSTREAMS=1: every B(q) on one side stream (A/B runs); default 2: alternating
  static std::atomic<int> g{-1};
  int m = g.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = std::getenv("LCB_UPDATE_BSTREAMS");
    m = (e && std::atoi(e) == 1) ? 1 : 2;
    g.store(m, std::memory_order_relaxed);
  }
  return m;
}
static bool pdl_mode() {   // LCB_UPDATE_PDL=1: chain kernels of the GPTQ block loop with programmatic dependent launch
  static std::atomic<int> g{-1};
  int m = g.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = std::getenv("LCB_UPDATE_PDL");
    m = (e && std::atoi(e) != 0) ? 1 : 0;
    g.store(m, std::memory_order_relaxed);
  }
  return m != 0;
}
static int64_t super_width() {   // LCB_UPDATE_SUPER = 256 / 512 / 1024 (A/B runs); default SUPER
  static std::atomic<int> g{-1};
  int m = g.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = std::getenv("LCB_UPDATE_SUPER");
    const int v = e ? std::atoi(e) : SUPER;
    m = (v == 256 || v == 512 || v == 1024) ? v : SUPER;
    g.store(m, std::memory_order_relaxed);
  }
  return m;
}
static int64_t super_cols(int64_t k) { return std::min<int64_t>(super_width(), ceil_div(k, BLK) * BLK); }

static size_t gptq_ws_floats(int64_t n, int64_t k, int block) {
  const size_t exact = (size_t)(2 * n * block);
  const size_t tg = (size_t)(8 * n * super_cols(k)) + (size_t)(4 * k * k);
  return std::max(exact, tg) + 64;
}

extern "C" size_t lcb_gptq_ws_bytes(int64_t n, int64_t k, int block) {
  return gptq_ws_floats(n, k, block) * sizeof(float);
}

namespace lcb {
// Shared block loop of GPTQ / GPTAQ (MODE_QUANT) and SparseGPT (MODE_SPARSE).  `pre_block(i1, count)`
// runs before each block step (SparseGPT: saliency + threshold + mask).
template <typename PreBlock>
static int run_block_loop(BlockArgs a, int mode, float* W, const float* U, const float* P, int64_t n, int64_t k,
                          float* wsf, cudaStream_t st, PreBlock pre_block) {
  const bool tg = use_tg(W, U, P, k);
  const int smem = (P ? 2 : 1) * BLK * ULD * (int)sizeof(float);
  const unsigned grid = (unsigned)ceil_div(n, ROWS_PER_CTA);
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(a.Q)) & 15) == 0 && k % 4 == 0;
  auto step = [&]() -> int {
    if (mode == MODE_QUANT && !P && a.group >= a.count && a.group > 0 && aligned16) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)ceil_div(n * (BLK / 4), 256), 1, 1);
      cfg.blockDim = dim3(256, 1, 1);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = (tg && pdl_mode()) ? 1 : 0;
      count_launch();
      LCB_CUDA(cudaLaunchKernelEx(&cfg, block_quant_kernel, a));
      return LCB_OK;
    }
    if (mode == MODE_SPARSE) block_step_kernel<MODE_SPARSE, false><<<grid, 256, smem, st>>>(a);
    else if (P) block_step_kernel<MODE_QUANT, true><<<grid, 256, smem, st>>>(a);
    else block_step_kernel<MODE_QUANT, false><<<grid, 256, smem, st>>>(a);
    LCB_LAUNCH_CHECK();
    return LCB_OK;
  };
  int rc;
  if (!tg) {
    a.Err = wsf; a.ErrLo = nullptr; a.ldE = BLK; a.eoff = 0;
    a.W1out = P ? wsf + n * BLK : nullptr; a.W1Lo = nullptr;
    for (int64_t i1 = 0; i1 < k; i1 += BLK) {
      const int64_t i2 = std::min<int64_t>(i1 + BLK, k);
      a.i1 = i1; a.count = (int)(i2 - i1);
      if ((rc = pre_block(i1, a.count)) != LCB_OK) return rc;
      if ((rc = step()) != LCB_OK) return rc;
      if (i2 < k) {
        // W[:, i2:] -= Err1 @ U[i1:i2, i2:]  (- W1 @ P[i1:i2, i2:])      ref gptq :265, gptaq :319
        rc = sgemm(gemm_args(a.Err, BLK, U + i1 * k + i2, k, W + i2, k, (int)n, (int)(k - i2), a.count, -1.0f, 1.0f, 0), st);
        if (rc != LCB_OK) return rc;
        if (P) {
          rc = sgemm(gemm_args(a.W1out, BLK, P + i1 * k + i2, k, W + i2, k, (int)n, (int)(k - i2), a.count, 1.0f, 1.0f, 0), st);
          if (rc != LCB_OK) return rc;
        }
      }
    }
    return LCB_OK;
  }
  // ---- tensor-core path with look-ahead (same scheme as the Cholesky panels, chol.cu): the block steps are a
  // latency-bound chain; after each step only the NEXT block's columns are updated on the caller's stream, the
  // rest of the super-block (and, per super-block, everything beyond the next one) on side streams underneath
  // the following steps.  All updates are L2 reduce-adds (they commute); events order them against the readers;
  // the error planes are double-buffered per super-block.
  const int64_t S = super_cols(k);
  float* planes = wsf;                       // [parity][ErrH | ErrL | W1H | W1L][n, S]
  float* UTh = wsf + 8 * n * S;
  float* UTl = UTh + k * k;
  float* PTh = UTl + k * k;
  float* PTl = PTh + k * k;
  if ((rc = split_tf32(U, k, (int)k, (int)k, UTh, UTl, k, /*transpose=*/1, st)) != LCB_OK) return rc;
  if (P && (rc = split_tf32(P, k, (int)k, (int)k, PTh, PTl, k, 1, st)) != LCB_OK) return rc;
  SideStreams* ss = side_streams();
  if (ss == nullptr) return LCB_ERR_CUDA;
  // The chain (quantiser + A(q) / S_A(J)) can move to the library's greatest-priority stream, between two events on the caller's
  // stream when LCB_CHAIN_PRIO=1 (an A/B switch: measured, no gain, so the default stays on the caller's stream).
  cudaStream_t const caller = st;
  if (mode == MODE_QUANT && chain_prio_mode()) {  // SparseGPT's pre_block (select + the caller's all-reduce) stays on `caller`
    LCB_CUDA(cudaEventRecord(ss->evIn, caller));
    LCB_CUDA(cudaStreamWaitEvent(ss->chain, ss->evIn, 0));
    st = ss->chain;
  }
  auto fail = [&](int code) {
    cudaStreamSynchronize(ss->s[0]);
    cudaStreamSynchronize(ss->s[1]);
    cudaStreamSynchronize(ss->s[2]);
    if (st != caller) cudaStreamSynchronize(st);
    return code;
  };
  // C[:, c0 : c0 + ncols] -= Err[:, e0 : e0 + kd] @ U[i1 + .., c0 ..]  (+ W1 @ P) on stream `s`
  const bool pdl = pdl_mode() && !P;
  auto update = [&](float* EH, float* EL, float* WH, float* WL, int64_t e0, int64_t urow, int64_t c0, int64_t ncols,
                    int kd, cudaStream_t s) -> int {
    int r = tgemm_nt(EH + e0, EL + e0, S, UTh + c0 * k + urow, UTl + c0 * k + urow, k, W + c0, k, (int)n, (int)ncols, kd,
                     -1.0f, (pdl && s == st) ? TG_PDL : 0, s);
    if (r != LCB_OK || !P) return r;
    return tgemm_nt(WH + e0, WL + e0, S, PTh + c0 * k + urow, PTl + c0 * k + urow, k, W + c0, k, (int)n, (int)ncols, kd,
                    1.0f, 0, s);
  };
  bool evB_live[2] = {false, false}, evS_live[2] = {false, false};
  int64_t q = 0, J = 0;
  for (int64_t s0 = 0; s0 < k; s0 += S, ++J) {
    const int64_t s1 = std::min<int64_t>(s0 + S, k);
    float* EH = planes + (J & 1) * 4 * n * S;
    float* EL = EH + n * S;
    float* WH = EL + n * S;
    float* WL = WH + n * S;
    // this super-block's columns were updated by S_A(J-1) (this stream) and S_B(J-2) (side stream 1), which also
    // read the planes of parity J & 1; the B updates of earlier super-blocks read the other parity
    if (evS_live[J & 1]) { LCB_CUDA(cudaStreamWaitEvent(st, ss->evS[J & 1], 0)); evS_live[J & 1] = false; }
    for (int e = 0; e < 2; ++e)
      if (evB_live[e]) { LCB_CUDA(cudaStreamWaitEvent(st, ss->evB[e], 0)); evB_live[e] = false; }
    a.Err = EH; a.ErrLo = EL; a.ldE = S;
    a.W1out = P ? WH : nullptr; a.W1Lo = P ? WL : nullptr;
    for (int64_t i1 = s0; i1 < s1; i1 += BLK, ++q) {
      const int64_t i2 = std::min<int64_t>(i1 + BLK, k);
      a.i1 = i1; a.count = (int)(i2 - i1); a.eoff = i1 - s0;
      // block q's columns: A(q-1) on this stream, B(q-2) on side stream 0
      if (evB_live[q & 1]) { LCB_CUDA(cudaStreamWaitEvent(st, ss->evB[q & 1], 0)); evB_live[q & 1] = false; }
      if ((rc = pre_block(i1, a.count)) != LCB_OK) return fail(rc);
      if ((rc = step()) != LCB_OK) return fail(rc);
      if (i2 < s1) {  // rest of the super-block, Kd = 128 (short last block: its zero padded columns add 0)
        const int kd = (int)std::min<int64_t>(BLK, S - a.eoff);
        const int64_t nA = std::min<int64_t>(BLK, s1 - i2), nB = s1 - i2 - nA;
        // PDL form: A(q) goes first so that it directly follows the quantiser in the stream (an event record in between would
        // break the programmatic edge); B(q) then also waits for A(q), which its two steps of slack absorb
        if (pdl && (rc = update(EH, EL, WH, WL, a.eoff, i1, i2, nA, kd, st)) != LCB_OK) return fail(rc);
        if (nB > 0) {  // B(q): beyond the next block, on a side stream
          // B(q) and B(q + 1) do not depend on each other (reduce-adds into W commute; each only needs its own quantiser), but on
          // ONE side stream they ran back to back, and a B launch (~25 us) is longer than a chain step (quantiser + A(q)): the
          // loop was paced by the side stream.  Alternating two streams lets consecutive B launches overlap.
          cudaStream_t sb = ss->s[(b_streams_mode() == 2 && (q & 1)) ? 2 : 0];
          LCB_CUDA(cudaEventRecord(ss->evP, st));
          LCB_CUDA(cudaStreamWaitEvent(sb, ss->evP, 0));
          if ((rc = update(EH, EL, WH, WL, a.eoff, i1, i2 + nA, nB, kd, sb)) != LCB_OK) return fail(rc);
          LCB_CUDA(cudaEventRecord(ss->evB[q & 1], sb));
          evB_live[q & 1] = true;
        }
        if (!pdl && (rc = update(EH, EL, WH, WL, a.eoff, i1, i2, nA, kd, st)) != LCB_OK) return fail(rc);  // A(q)
      }
    }
    if (s1 < k) {  // everything beyond the super-block, Kd = S
      const int kd = (int)(s1 - s0);
      const int64_t nA = std::min<int64_t>(S, k - s1), nB = k - s1 - nA;
      if (pdl && (rc = update(EH, EL, WH, WL, 0, s0, s1, nA, kd, st)) != LCB_OK) return fail(rc);  // S_A(J) first (see A(q))
      if (nB > 0) {  // S_B(J): beyond the next super-block, side stream 1
        LCB_CUDA(cudaEventRecord(ss->evP, st));
        LCB_CUDA(cudaStreamWaitEvent(ss->s[1], ss->evP, 0));
        if ((rc = update(EH, EL, WH, WL, 0, s0, s1 + nA, nB, kd, ss->s[1])) != LCB_OK) return fail(rc);
        LCB_CUDA(cudaEventRecord(ss->evS[J & 1], ss->s[1]));
        evS_live[J & 1] = true;
      }
      if (!pdl && (rc = update(EH, EL, WH, WL, 0, s0, s1, nA, kd, st)) != LCB_OK) return fail(rc);  // S_A(J)
    }
  }
  for (int e = 0; e < 2; ++e) {
    if (evB_live[e]) LCB_CUDA(cudaStreamWaitEvent(st, ss->evB[e], 0));
    if (evS_live[e]) LCB_CUDA(cudaStreamWaitEvent(st, ss->evS[e], 0));
  }
  if (st != caller) {
    LCB_CUDA(cudaEventRecord(ss->evOut, st));
    LCB_CUDA(cudaStreamWaitEvent(caller, ss->evOut, 0));
  }
  return LCB_OK;
}
}  // namespace lcb

extern "C" int lcb_gptq_update(const lcb_quant_cfg* cfg, float* W, float* Q, const float* U, const float* P,
                               const float* scales, const float* zeros, const uint8_t* keep, int64_t n, int64_t k,
                               int64_t group, int block, void* ws, size_t ws_bytes, void* stream) {
  LCB_REQUIRE(cfg && W && Q && U && scales && zeros, "lcb_gptq_update: NULL pointer");
  LCB_REQUIRE(block == BLK, "lcb_gptq_update: block must be 128 (the reference's block_size)");
  LCB_REQUIRE(n > 0 && k > 0, "lcb_gptq_update: bad shape");
  if (group > 0) {
    LCB_REQUIRE(BLK % group == 0 && group % CPT == 0 && k % group == 0,
                "lcb_gptq_update: group must divide 128, be a multiple of 16 and divide k (got %lld)", (long long)group);
  }
  if (ws == nullptr || ws_bytes < lcb_gptq_ws_bytes(n, k, block) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("lcb_gptq_update: 16-byte aligned workspace of %zu bytes needed", lcb_gptq_ws_bytes(n, k, block));
    return LCB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BlockArgs a{};
  a.W = W; a.Q = Q; a.U = U; a.P = P; a.scales = scales; a.zeros = zeros; a.keep = keep;
  a.n = n; a.k = k; a.group = group; a.G = group > 0 ? k / group : 1;
  a.c.qtype = cfg->qtype; a.c.zero_point = cfg->zero_point ? 1 : 0;
  a.c.scale_emax = (float)((1 << ((cfg->scale_ebits > 0 ? cfg->scale_ebits : 8) - 1)) - 1);
  a.c.f = make_fmt(cfg->elem);
  const int smem = (P ? 2 : 1) * BLK * ULD * (int)sizeof(float);
  LCB_CUDA(cudaFuncSetAttribute(block_step_kernel<MODE_QUANT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  LCB_CUDA(cudaFuncSetAttribute(block_step_kernel<MODE_QUANT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  return run_block_loop(a, MODE_QUANT, W, U, P, n, k, static_cast<float*>(ws), st, [](int64_t, int) { return LCB_OK; });
}

// ---------------------------------------------------------------------------------------------------
// Entry / exit of update_weight around the block loop, each one pass over the weight:
//   gather : W = weight.float(); MASK = W != 0; W[:, dead] = 0; act-order column permutation
//            (ref: gptq/core.py:164-201) -> permuted fp32 work matrix + keep mask
//   scatter: inverse permutation + cast back to the weight dtype (ref: gptq/core.py:267-278)
namespace lcb {
template <typename T>
__global__ void __launch_bounds__(256) gptq_gather_kernel(const T* __restrict__ W, const int64_t* __restrict__ col_perm,
                                                          const uint8_t* __restrict__ dead, float* __restrict__ Wp,
                                                          uint8_t* __restrict__ keep, int64_t n, int64_t k) {
  const int64_t total = n * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k, j = i - r * k;
    const int64_t src = col_perm ? col_perm[j] : j;
    const float v = to_f<T>(W[r * k + src]);
    keep[i] = v != 0.0f;  // MASK is taken before the dead columns are zeroed (ref :166-177)
    Wp[i] = (dead && dead[src]) ? 0.0f : v;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) gptq_scatter_kernel(const float* __restrict__ Q, const int64_t* __restrict__ col_perm,
                                                           T* __restrict__ out, int64_t n, int64_t k) {
  const int64_t total = n * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k, j = i - r * k;
    const int64_t dst = col_perm ? col_perm[j] : j;
    out[r * k + dst] = from_f<T>(Q[i]);
  }
}
}  // namespace lcb

extern "C" int lcb_gptq_gather(const void* W, int dtype, const int64_t* col_perm, const uint8_t* dead, float* Wp,
                               uint8_t* keep, int64_t n, int64_t k, void* stream) {
  LCB_REQUIRE(W && Wp && keep && n > 0 && k > 0, "lcb_gptq_gather: bad arguments");
  LCB_REQUIRE(dtype == LCB_F32 || dtype == LCB_BF16, "lcb_gptq_gather: dtype must be LCB_F32 or LCB_BF16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t g = ceil_div(n * k, 256 * 8);
  if (g > (int64_t)sm_count() * 16) g = (int64_t)sm_count() * 16;
  if (dtype == LCB_BF16)
    gptq_gather_kernel<__nv_bfloat16><<<(unsigned)g, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(W), col_perm, dead, Wp,
                                                                     keep, n, k);
  else
    gptq_gather_kernel<float><<<(unsigned)g, 256, 0, st>>>(static_cast<const float*>(W), col_perm, dead, Wp, keep, n, k);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_gptq_scatter(const float* Q, const int64_t* col_perm, void* out, int dtype, int64_t n, int64_t k,
                                void* stream) {
  LCB_REQUIRE(Q && out && n > 0 && k > 0, "lcb_gptq_scatter: bad arguments");
  LCB_REQUIRE(dtype == LCB_F32 || dtype == LCB_BF16, "lcb_gptq_scatter: dtype must be LCB_F32 or LCB_BF16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t g = ceil_div(n * k, 256 * 8);
  if (g > (int64_t)sm_count() * 16) g = (int64_t)sm_count() * 16;
  if (dtype == LCB_BF16)
    gptq_scatter_kernel<__nv_bfloat16><<<(unsigned)g, 256, 0, st>>>(Q, col_perm, static_cast<__nv_bfloat16*>(out), n, k);
  else
    gptq_scatter_kernel<float><<<(unsigned)g, 256, 0, st>>>(Q, col_perm, static_cast<float*>(out), n, k);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

// P = alpha * triu(dXXT @ U^T, 1) @ U   (ref: gptaq/core.py:272)
namespace lcb {
__global__ void triu1_scale_kernel(float* A, int64_t k, float alpha) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j < k) A[i * k + j] = (j > i) ? alpha * A[i * k + j] : 0.0f;
}
}  // namespace lcb

extern "C" size_t lcb_gptaq_p_ws_bytes(int64_t k) { return (size_t)(7 * k * k + 64) * sizeof(float); }

extern "C" int lcb_gptaq_p(float* P, float* dxxt, const float* U, int64_t k, float alpha, void* ws, size_t ws_bytes,
                           void* stream) {
  LCB_REQUIRE(P && dxxt && U && k > 0, "lcb_gptaq_p: bad arguments");
  const bool tg = gemm_mode() == 1 && tg_ok(P, k) && tg_ok(ws, 4);
  const size_t need = tg ? lcb_gptaq_p_ws_bytes(k) : (size_t)(k * k) * sizeof(float);
  if (ws == nullptr || ws_bytes < need) {
    set_error("lcb_gptaq_p: workspace of %zu bytes needed", need);
    return LCB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* T = static_cast<float*>(ws);
  dim3 g2((unsigned)ceil_div(k, 256), (unsigned)k);
  int rc;
  if (!tg) {
    rc = sgemm(gemm_args(dxxt, k, U, k, T, k, (int)k, (int)k, (int)k, 1.0f, 0.0f, /*transB=*/1), st);
    if (rc != LCB_OK) return rc;
    triu1_scale_kernel<<<g2, 256, 0, st>>>(T, k, alpha);
    LCB_LAUNCH_CHECK();
    return sgemm(gemm_args(T, k, U, k, P, k, (int)k, (int)k, (int)k, 1.0f, 0.0f, 0), st);
  }
  float* Ah = T + k * k;   // planes of dXXT, later of triu(T, 1)
  float* Al = Ah + k * k;
  float* Uh = Al + k * k;  // planes of U
  float* Ul = Uh + k * k;
  float* UTh = Ul + k * k;  // planes of U^T
  float* UTl = UTh + k * k;
  if ((rc = split_tf32(dxxt, k, (int)k, (int)k, Ah, Al, k, 0, st)) != LCB_OK) return rc;
  if ((rc = split_tf32(U, k, (int)k, (int)k, Uh, Ul, k, 0, st)) != LCB_OK) return rc;
  if ((rc = split_tf32(U, k, (int)k, (int)k, UTh, UTl, k, 1, st)) != LCB_OK) return rc;
  // T = dXXT @ U^T
  rc = tgemm_nt(Ah, Al, k, Uh, Ul, k, T, k, (int)k, (int)k, (int)k, 1.0f, TG_STORE, st);
  if (rc != LCB_OK) return rc;
  triu1_scale_kernel<<<g2, 256, 0, st>>>(T, k, alpha);
  LCB_LAUNCH_CHECK();
  if ((rc = split_tf32(T, k, (int)k, (int)k, Ah, Al, k, 0, st)) != LCB_OK) return rc;
  // P = triu(T, 1) @ U = T' @ (U^T)^T; T' is strictly upper triangular: the k-loop starts at the row tile
  return tgemm_nt(Ah, Al, k, UTh, UTl, k, P, k, (int)k, (int)k, (int)k, 1.0f, TG_STORE | TG_A_UPPER, st);
}

static size_t sparse_tail_bytes(int64_t n) { return (size_t)(n * BLK) * (sizeof(float) + 1) + 512 + select_ws_bytes(); }

extern "C" size_t lcb_sparsegpt_ws_bytes(int64_t n, int64_t k, int block) {
  return gptq_ws_floats(n, k, block) * sizeof(float) + sparse_tail_bytes(n);
}

static int sparsegpt_update_impl(float* W, const float* U, double sparsity, int64_t n, int64_t n_total, int64_t k, int block,
                                 void* ws, size_t ws_bytes, lcb_reduce_u32_fn reduce, void* reduce_user, void* stream) {
  LCB_REQUIRE(W && U && n > 0 && k > 0 && n_total >= n, "lcb_sparsegpt_update: bad arguments");
  LCB_REQUIRE(block == BLK, "lcb_sparsegpt_update: block must be 128");
  if (ws == nullptr || ws_bytes < lcb_sparsegpt_ws_bytes(n, k, block) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("lcb_sparsegpt_update: 16-byte aligned workspace of %zu bytes needed", lcb_sparsegpt_ws_bytes(n, k, block));
    return LCB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* wsf = static_cast<float*>(ws);
  float* metric = wsf + gptq_ws_floats(n, k, block);
  float* thresh = metric + n * BLK;                                 // 64 floats reserved
  uint8_t* prune = reinterpret_cast<uint8_t*>(thresh + 64);          // [n, BLK]
  void* sel_ws = prune + ((n * BLK + 255) / 256) * 256;
  BlockArgs a{};
  a.W = W; a.U = U; a.prune = prune; a.n = n; a.k = k; a.group = 0; a.G = 1;
  a.c.f = make_fmt(LCB_E_INT8);
  const int smem = BLK * ULD * (int)sizeof(float);
  LCB_CUDA(cudaFuncSetAttribute(block_step_kernel<MODE_SPARSE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  auto pre = [&](int64_t i1, int count) -> int {
    // mask of the block from its state at block entry (ref: sparsegpt/core.py:201-203)
    const int64_t numel = n * count;
    sparse_metric_kernel<<<(unsigned)ceil_div(numel, 256), 256, 0, st>>>(W, U, metric, n, k, i1, count);
    LCB_LAUNCH_CHECK();
    const int64_t numel_total = n_total * count;  // the block's rows over all ranks
    int64_t kth = (int64_t)((double)numel_total * sparsity);  // int(tmp.numel() * sparsity_ratio)
    if (kth >= numel_total) kth = numel_total - 1;
    int rc = select_kth_f32(metric, numel, kth, thresh, sel_ws, select_ws_bytes(), st, reduce, reduce_user);
    if (rc != LCB_OK) return rc;
    sparse_mask_kernel<<<(unsigned)ceil_div(numel, 256), 256, 0, st>>>(metric, thresh, prune, n, count);
    LCB_LAUNCH_CHECK();
    return LCB_OK;
  };
  return run_block_loop(a, MODE_SPARSE, W, U, nullptr, n, k, wsf, st, pre);
}

extern "C" int lcb_sparsegpt_update(float* W, const float* U, double sparsity, int64_t n, int64_t k, int block,
                                    void* ws, size_t ws_bytes, void* stream) {
  return sparsegpt_update_impl(W, U, sparsity, n, n, k, block, ws, ws_bytes, nullptr, nullptr, stream);
}

extern "C" int lcb_sparsegpt_update_sharded(float* W, const float* U, double sparsity, int64_t n_local, int64_t n_total,
                                            int64_t k, int block, void* ws, size_t ws_bytes, lcb_reduce_u32_fn reduce,
                                            void* reduce_user, void* stream) {
  LCB_REQUIRE(reduce != nullptr, "lcb_sparsegpt_update_sharded: a reduce callback is required");
  return sparsegpt_update_impl(W, U, sparsity, n_local, n_total, k, block, ws, ws_bytes, reduce, reduce_user, stream);
}

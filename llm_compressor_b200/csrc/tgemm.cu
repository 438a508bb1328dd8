// fp32-accurate tensor-core GEMM for the solver stages (lazy-batch update, Cholesky SYRK):
//
//     C[M,N] += alpha * A[M,Kd] * B[N,Kd]^T                (row-major, "NT", K-major operands)
//
// on tcgen05 with the 3xTF32 split:  x = hi + lo with hi = tf32(x) (round to nearest) and
// lo = x - hi;  A*B^T ~= Ahi*Bhi^T + Ahi*Blo^T + Alo*Bhi^T accumulated in fp32 in TMEM (the lo*lo
// term is below 2^-22 relative).  SURVEY N3 measured 8.6e-6 flipped int4 codes for this scheme
// against exact fp32 on the GPTQ trailing update, inside the north_star tolerance.
// The hi / lo planes are produced by the operand's producer (block_step_kernel for Err, a
// transpose-split pass for U, a split pass for the Cholesky panel), so the GEMM itself streams four
// K-major fp32 tiles per stage through TMA (128B swizzle), issues 12 tcgen05.mma.kind::tf32
// (M=128, N=256, K=8) per 32-wide k block, and reduce-adds the accumulator into C with TMA
// (cp.reduce.async.bulk.tensor .add.f32 at L2) -- C is never read by the SMs.
#include <algorithm>
#include <atomic>

#include "linalg.cuh"
#include "umma.cuh"

namespace lcb {

namespace {

constexpr int BM = 128, BK = 32;            // BK fp32 = one 128 B swizzle row
constexpr int UK = 8;                       // K per tf32 MMA
constexpr int A_BYTES = BM * BK * 4;        // 16384
constexpr int OUT_BUF_BYTES = 32 * 32 * 4;
constexpr int OUT_BYTES = 4 * OUT_BUF_BYTES;  // one staging box per epilogue warp
constexpr int NUM_THREADS = 192;
// Tile width BN = 256 (two TMEM accumulators of 256 columns, two 96 KB stages) for the big contractions; BN = 128 (three
// 64 KB stages) for outputs of at most 128 columns -- the next-block update of the GPTQ chain, whose 256-wide tile
// loaded and multiplied a half that was thrown away.
template <int BN>
struct Tile {
  static constexpr int STAGES = BN == 256 ? 2 : 3;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // hi + lo planes
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + 256 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major operand, 128B swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10)/[10,13), K-major both, N>>3 at [17,23), M>>4 at [24,29)
template <int BN>
constexpr uint32_t make_idesc_tf32() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct TgArgs {
  int M, N, Kd;
  float alpha;
  int tiles_m, tiles_n;
  int flags;  // TG_* (linalg.cuh)
  int kcb;    // k blocks per accumulation chain (work unit = tile x chain)
  int nchunks;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tgemm_nt_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                const __grid_constant__ CUtensorMap map_c, const __grid_constant__ TgArgs a) {
  constexpr int STAGES = Tile<BN>::STAGES, B_BYTES = Tile<BN>::B_BYTES, STAGE_BYTES = Tile<BN>::STAGE_BYTES;
  constexpr int TMEM_COLS = Tile<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* smem_out = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + OUT_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.tiles_m * a.tiles_n;
  const int kblocks = (a.Kd + BK - 1) / BK;
  const bool lower_only = (a.flags & TG_LOWER_OUT) != 0;
  // Work unit u = (tile, accumulation chain): the tile's k range (cut when A is triangular: the zero
  // part is never loaded) is processed in chains of at most kcb k-blocks, each reduce-added into C on
  // its own, which bounds the length of the truncating fp32 accumulation inside the tensor core.
  const int nunits = num_tiles * a.nchunks;
  auto unit = [&](int u, int& m0, int& n0, int& kb0, int& kb1) -> bool {
    const int tile = u % num_tiles, ch = u / num_tiles;
    m0 = (tile / a.tiles_n) * BM;
    n0 = (tile % a.tiles_n) * BN;
    if (lower_only && n0 > m0 + BM - 1) return false;
    int lo = 0, hi = kblocks;
    if (a.flags & TG_A_LOWER) hi = min(kblocks, (m0 + BM + BK - 1) / BK);
    if (a.flags & TG_A_UPPER) lo = min(m0 / BK, kblocks - 1);
    kb0 = max(lo, ch * a.kcb);
    kb1 = min(hi, (ch + 1) * a.kcb);
    return kb1 > kb0;
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ah)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_al)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch (TG_PDL; no-ops otherwise): everything above ran under the tail of the preceding kernel;
  // let the next kernel of the stream start its own prologue, then wait until the preceding grid has completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        int m0, n0, kb0, kb1;
        if (!unit(u, m0, n0, kb0, kb1)) continue;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* s = smem + stage * STAGE_BYTES;
          tma_load_2d(&map_ah, &full[stage], s, kb * BK, m0);
          tma_load_2d(&map_al, &full[stage], s + A_BYTES, kb * BK, m0);
          tma_load_2d(&map_bh, &full[stage], s + 2 * A_BYTES, kb * BK, n0);
          tma_load_2d(&map_bl, &full[stage], s + 2 * A_BYTES + B_BYTES, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32<BN>();
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        int m0, n0, kb0, kb1;
        if (!unit(u, m0, n0, kb0, kb1)) continue;
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sah = s, sal = s + A_BYTES, sbh = s + 2 * A_BYTES, sbl = s + 2 * A_BYTES + B_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint32_t ko = k * UK * 4;  // 32 B inside the 128 B swizzle row
            const uint64_t dah = make_desc_k(sah + ko), dal = make_desc_k(sal + ko);
            const uint64_t dbh = make_desc_k(sbh + ko), dbl = make_desc_k(sbl + ko);
            umma_tf32(tmem_d, dal, dbh, idesc, (kb > kb0 || k > 0) ? 1u : 0u);  // small terms first
            umma_tf32(tmem_d, dah, dbl, idesc, 1u);
            umma_tf32(tmem_d, dah, dbh, idesc, 1u);
          }
          umma_commit(&empty[stage]);
          if (kb == kb1 - 1) umma_commit(&tfull[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ++iter;
      }
    }
  } else {
    const int q = warp & 3;
    uint8_t* buf = smem_out + (warp - 2) * OUT_BUF_BYTES;
    int iter = 0;
    bool pending = false;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      int m0, n0, kb0, kb1;
      if (!unit(u, m0, n0, kb0, kb1)) continue;
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;");
      const int row0 = m0 + q * 32;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        const bool live = row0 < a.M && col0 < a.N && !(lower_only && col0 > row0 + 31);
        if (!live) continue;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), v);
        if (pending) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        uint8_t* rowp = buf + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 o = make_float4(a.alpha * v[4 * j], a.alpha * v[4 * j + 1], a.alpha * v[4 * j + 2],
                                       a.alpha * v[4 * j + 3]);
          *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (a.flags & TG_STORE) tma_store_2d(&map_c, buf, col0, row0);
          else tma_reduce_add_2d(&map_c, buf, col0, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        pending = true;
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      ++iter;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// hi = tf32(x) (round to nearest, ties away), lo = x - hi; optional transpose: out[c][r] = split(in[r][c])
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ in, int64_t ld_in, int rows, int cols,
                                                         float* __restrict__ hi, float* __restrict__ lo, int64_t ld_out,
                                                         int transpose) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[(int64_t)r * ld_in + c] : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    float v;
    int64_t o;
    bool ok;
    if (transpose) {
      const int r = c0 + i, c = r0 + tx;  // output row = input column
      v = tile[tx][i];
      ok = r < cols && c < rows;
      o = (int64_t)r * ld_out + c;
    } else {
      const int r = r0 + i, c = c0 + tx;
      v = tile[i][tx];
      ok = r < rows && c < cols;
      o = (int64_t)r * ld_out + c;
    }
    if (ok) {
      uint32_t hb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
      const float h = __uint_as_float(hb);
      hi[o] = h;
      lo[o] = __fsub_rn(v, h);
    }
  }
}

int make_map_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc != LCB_OK) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 %lld x %lld ld %lld) failed with CUresult %d", (long long)rows,
              (long long)cols, (long long)ld, (int)r);
    return LCB_ERR_CUDA;
  }
  return LCB_OK;
}

}  // namespace

// C[M,N] += alpha * (Ah + Al)[M,Kd] * (Bh + Bl)[N,Kd]^T   (3xTF32).  All bases 16 B aligned, lds % 4 == 0.
int tgemm_nt(const float* Ah, const float* Al, int64_t lda, const float* Bh, const float* Bl, int64_t ldb, float* C,
             int64_t ldc, int M, int N, int Kd, float alpha, int flags, cudaStream_t st, int kchain) {
  if (M <= 0 || N <= 0 || Kd <= 0) return LCB_OK;
  const int kblocks = (int)ceil_div(Kd, BK);
  const int kcb = kchain > 0 ? (int)std::max<int64_t>(1, kchain / BK) : kblocks;
  const int nchunks = (int)ceil_div(kblocks, kcb);
  if (nchunks > 1 && (flags & TG_STORE)) {
    // several chains per tile are combined by the L2 reduce-add: start from zero
    LCB_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
    flags &= ~TG_STORE;
  }
  CUtensorMap mah, mal, mbh, mbl, mc;
  int rc;
  if ((rc = make_map_f32(&mah, Ah, M, Kd, lda, BK, BM)) != LCB_OK) return rc;
  if ((rc = make_map_f32(&mal, Al, M, Kd, lda, BK, BM)) != LCB_OK) return rc;
  const int BN = N <= 128 ? 128 : 256;
  if ((rc = make_map_f32(&mbh, Bh, N, Kd, ldb, BK, BN)) != LCB_OK) return rc;
  if ((rc = make_map_f32(&mbl, Bl, N, Kd, ldb, BK, BN)) != LCB_OK) return rc;
  if ((rc = make_map_f32(&mc, C, M, N, ldc, 32, 32)) != LCB_OK) return rc;
  TgArgs a{};
  a.M = M; a.N = N; a.Kd = Kd; a.alpha = alpha; a.flags = flags;
  a.tiles_m = (int)ceil_div(M, BM); a.tiles_n = (int)ceil_div(N, BN);
  a.kcb = kcb; a.nchunks = nchunks;
  const int tiles = a.tiles_m * a.tiles_n * nchunks;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (flags & TG_PDL) ? 1 : 0;
  count_launch();
  if (BN == 128) {
    LCB_CUDA(cudaFuncSetAttribute(tgemm_nt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<128>::SMEM_BYTES));
    cfg.dynamicSmemBytes = Tile<128>::SMEM_BYTES;
    LCB_CUDA(cudaLaunchKernelEx(&cfg, tgemm_nt_kernel<128>, mah, mal, mbh, mbl, mc, a));
  } else {
    LCB_CUDA(cudaFuncSetAttribute(tgemm_nt_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<256>::SMEM_BYTES));
    cfg.dynamicSmemBytes = Tile<256>::SMEM_BYTES;
    LCB_CUDA(cudaLaunchKernelEx(&cfg, tgemm_nt_kernel<256>, mah, mal, mbh, mbl, mc, a));
  }
  return LCB_OK;
}

int split_tf32(const float* in, int64_t ld_in, int rows, int cols, float* hi, float* lo, int64_t ld_out, int transpose,
               cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return LCB_OK;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  split_tf32_kernel<<<grid, 256, 0, st>>>(in, ld_in, rows, cols, hi, lo, ld_out, transpose);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

}  // namespace lcb

using namespace lcb;

namespace lcb {
static std::atomic<int> g_gemm_mode{1};
int gemm_mode() { return g_gemm_mode.load(std::memory_order_relaxed); }
}  // namespace lcb

extern "C" int lcb_set_gemm_mode(int mode) {
  return g_gemm_mode.exchange(mode ? 1 : 0, std::memory_order_relaxed);
}

extern "C" size_t lcb_tgemm_ws_bytes(int64_t m, int64_t n, int64_t kd) {
  return (size_t)(2 * (m + n) * kd + 64) * sizeof(float);
}

extern "C" int lcb_tgemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t m,
                            int64_t n, int64_t kd, float alpha, int accumulate, int kchain, void* ws, size_t ws_bytes,
                            void* stream) {
  LCB_REQUIRE(A && B && C && m > 0 && n > 0 && kd > 0, "lcb_tgemm_nt: bad arguments");
  LCB_REQUIRE(kd % 4 == 0 && tg_ok(C, ldc), "lcb_tgemm_nt: kd and ldc must be multiples of 4, C 16-byte aligned");
  if (ws == nullptr || ws_bytes < lcb_tgemm_ws_bytes(m, n, kd) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("lcb_tgemm_nt: 16-byte aligned workspace of %zu bytes needed", lcb_tgemm_ws_bytes(m, n, kd));
    return LCB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ah = static_cast<float*>(ws);
  float* al = ah + m * kd;
  float* bh = al + m * kd;
  float* bl = bh + n * kd;
  int rc = split_tf32(A, lda, (int)m, (int)kd, ah, al, kd, 0, st);
  if (rc != LCB_OK) return rc;
  rc = split_tf32(B, ldb, (int)n, (int)kd, bh, bl, kd, 0, st);
  if (rc != LCB_OK) return rc;
  return tgemm_nt(ah, al, kd, bh, bl, kd, C, ldc, (int)m, (int)n, (int)kd, alpha, accumulate ? 0 : TG_STORE, st, kchain);
}

// (a) Calibration statistics on the 5th-generation tensor cores.
//
//   H    = beta * H    + alpha * X^T X             ref: gptq/core.py:103-119, sparsegpt/core.py:85-101
//   dXXT = beta * dXXT + alpha * (X_fp - X)^T X    ref: gptaq/core.py:116-141
//   s    = beta * s    + alpha * sum_t X[t,:]^2    ref: wanda/core.py:92-105
//
// X is the token-major bf16 activation [T, K] exactly as the forward hook receives it, so both MMA
// operands are "MN-major" views of the same tensor: A = X[:, m-tile]^T, B = X[:, n-tile]^T with
// the token axis as the reduction dimension.  TMA (128B swizzle, 64-channel x 64-token boxes)
// stages the operands in shared memory, one elected thread issues tcgen05.mma (M=128, N=256,
// K=16, bf16 x bf16 -> fp32) into TMEM, and four epilogue warps read the accumulator back with
// tcgen05.ld, stage it in swizzled shared memory and hand it to the TMA engine as a
// cp.reduce.async.bulk.tensor (fp32 add performed at L2): the SMs never read H.
// bf16 products are exact in fp32, so the result differs from the reference's fp32 SGEMM only by
// summation order (the tensor core accumulates a 2048-token chain with truncation: measured
// bias ~ -4e-6 relative).
//
// Scheduling: persistent stream-K.  The (tile, 64-token block) work units are split evenly over
// one CTA per SM; because every contribution is an L2 reduce-add, a tile whose token range is cut
// between two CTAs needs no fix-up pass.  With `upper_only` only the tiles that intersect the
// upper triangle are computed (X^T X is symmetric; lcb_hessian_finalize mirrors and scales once).
// Two 256-column TMEM accumulators let the epilogue of one segment overlap the MMAs of the next;
// a 4-stage smem ring decouples TMA from MMA.
#include <cuda.h>

#include "umma.cuh"

#include <atomic>
#include <cstdlib>

namespace lcb {

// Which tcgen05 kernel accumulates X^T X: one CTA per 128 x 256 tile (cta_group::1) or CTA pairs on 256 x 256
// tiles (cta_group::2).  Measured on B200 (2048 tokens, symmetric half): k = 8192 pair 0.117 ms vs 0.129 ms,
// k = 3072 pair 0.031 ms vs 0.029 ms (fewer, larger work units and two cluster barriers per launch), so the
// pair kernel is used from k = 4096 up.  LCB_HESSIAN_PAIR = 0 / 1 forces one of them (A/B measurements).
static std::atomic<int> g_hessian_pair{-2};
static bool hessian_pair_mode(int64_t k) {
  int m = g_hessian_pair.load(std::memory_order_relaxed);
  if (m == -2) {
    const char* e = std::getenv("LCB_HESSIAN_PAIR");
    m = e ? (std::atoi(e) != 0 ? 1 : 0) : -1;
    g_hessian_pair.store(m, std::memory_order_relaxed);
  }
  return m < 0 ? k >= 4096 : m != 0;
}

namespace {

constexpr int BM = 128;         // UMMA M (rows of the H tile)
constexpr int BN = 256;         // UMMA N (columns of the H tile)
constexpr int BKT = 64;         // tokens per pipeline stage
constexpr int UMMA_K = 16;      // tokens per tcgen05.mma (bf16)
constexpr int STAGES = 4;
constexpr int BOX_C = 64;       // channels per TMA box (64 * 2 B = one 128 B swizzle row)
constexpr int BOX_BYTES = BOX_C * BKT * 2;               // 8192
constexpr int A_STAGE_BYTES = (BM / BOX_C) * BOX_BYTES;  // 16384
constexpr int B_STAGE_BYTES = (BN / BOX_C) * BOX_BYTES;  // 32768
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int OUT_BOX = 32;                              // 32 x 32 fp32 store box (128 B rows)
constexpr int OUT_BUF_BYTES = OUT_BOX * OUT_BOX * 4;     // 4096
constexpr int OUT_BYTES = 4 * 2 * OUT_BUF_BYTES;         // 4 epilogue warps x double buffer
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + 256 + 1024;  // + barriers + alignment slack
constexpr int NUM_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int TMEM_COLS = 512;    // two accumulators of BN fp32 columns
constexpr int MAX_TILES_M = 128;  // k <= 16384

// Shared-memory matrix descriptor, MN-major operand, 128B swizzle (cute::UMMA::SmemDescriptor):
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (stride between 64-channel
//   chunks), [32,46) stride byte offset >> 4 (stride between 8-token groups), [46,48) version = 1,
//   [61,64) layout type (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((BOX_BYTES >> 4) & 0x3fff) << 16;  // LBO: next 64-channel box
  d |= (uint64_t)((1024 >> 4) & 0x3fff) << 32;       // SBO: next 8 tokens (8 rows of 128 B)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1)
// at [7,10)/[10,13), a_major / b_major = MN (1) at 15 / 16, N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

constexpr int MAX_SAMPLES = 8;  // hook inputs accumulated by one launch (lcb_hessian_accum_multi)
struct alignas(64) XMaps {
  CUtensorMap m[MAX_SAMPLES];
};

struct HessArgs {
  int64_t k;       // channels
  int64_t tokens;  // per sample
  int kb_per_sample;  // 64-token blocks per sample; kblocks = samples * kb_per_sample
  float alpha;
  int tiles_m, tiles_n;
  int num_tiles;   // computed tiles
  int kblocks;     // 64-token blocks per tile
  int upper_only;
  int sched;       // 1: whole tiles in lock step + stream-K remainder, 0: stream-K over all units
  int16_t row_start[MAX_TILES_M + 1];  // prefix sum of computed tiles per tile row
};

// Work schedule of one CTA (or CTA pair): a sequence of segments (tile, [kb0, kb1)).
//   1. whole tiles  cta, cta + G, cta + 2 G, ...  (G = CTAs / pairs in the grid), each over ALL token blocks: every CTA
//      sweeps the token blocks 0 .. kblocks-1 in step with the others, so at any moment the whole grid reads the same
//      few token blocks of X -- X streams through L2 once even when several deferred hook inputs (4 x 33 MB at K = 8192)
//      no longer fit it, and the tiles in flight are neighbours in the tile list (shared row slices);
//   2. the num_tiles % G left-over tiles, cut evenly over the grid stream-K style (partial tiles need no fix-up: every
//      contribution is an L2 reduce-add).  `sched` = 0 keeps the plain stream-K split of all units (round-1 v1 schedule).
struct WorkIter {
  int whole, cta, G, kblocks, base_tile, i;
  int64_t u, r1;
  __device__ __forceinline__ WorkIter(const HessArgs& a, int cta_, int G_) {
    cta = cta_; G = G_; kblocks = a.kblocks; i = 0;
    whole = a.sched ? a.num_tiles / G : 0;
    base_tile = whole * G;
    const int64_t rem = (int64_t)(a.num_tiles - base_tile) * kblocks;
    u = rem * cta / G;
    r1 = rem * (cta + 1) / G;
  }
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (i < whole) {
      tile = cta + i * G; kb0 = 0; kb1 = kblocks; ++i;
      return true;
    }
    if (u < r1) {
      const int t = (int)(u / kblocks);
      kb0 = (int)(u - (int64_t)t * kblocks);
      kb1 = (int)min((int64_t)kblocks, kb0 + (r1 - u));
      tile = base_tile + t;
      u += kb1 - kb0;
      return true;
    }
    return false;
  }
};

// linear computed-tile index -> (m0, n0) for bm x BN tiles
__device__ __forceinline__ void tile_coords(const HessArgs& a, int tile, int& m0, int& n0, int bm = BM) {
  int lo = 0, hi = a.tiles_m;  // largest mt with row_start[mt] <= tile
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.row_start[mid] <= tile) lo = mid; else hi = mid;
  }
  const int first_nt = a.upper_only ? (lo * bm) / BN : 0;
  m0 = lo * bm;
  n0 = (first_nt + (tile - a.row_start[lo])) * BN;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
hessian_umma_kernel(const __grid_constant__ XMaps maps_a, const __grid_constant__ XMaps maps_b,
                    const __grid_constant__ CUtensorMap map_h, const __grid_constant__ HessArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_out = smem + STAGES * STAGE_BYTES;  // 1024-aligned: [warp 0..3][buf 0..1][4096]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + OUT_BYTES);
  uint64_t* full = bars;                     // [STAGES]   TMA -> MMA
  uint64_t* empty = bars + STAGES;           // [STAGES]   MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;       // [2]        MMA -> epilogue
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sched_cta = (int)blockIdx.x, sched_n = (int)gridDim.x;  // see WorkIter

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps_a.m[0])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps_b.m[0])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_h)) : "memory");
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int tile, kb0, kb1;
      for (WorkIter it(a, sched_cta, sched_n); it.next(tile, kb0, kb1);) {
        int m0, n0;
        tile_coords(a, tile, m0, n0);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int smp = kb / a.kb_per_sample, kbl = kb - smp * a.kb_per_sample;  // hook input, token block inside it
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
#pragma unroll
          for (int h = 0; h < BM / BOX_C; ++h) tma_load_2d(&maps_a.m[smp], &full[stage], sa + h * BOX_BYTES, m0 + h * BOX_C, kbl * BKT);
#pragma unroll
          for (int h = 0; h < BN / BOX_C; ++h) tma_load_2d(&maps_b.m[smp], &full[stage], sb + h * BOX_BYTES, n0 + h * BOX_C, kbl * BKT);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc();
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      int tile, kb0, kb1;
      for (WorkIter it(a, sched_cta, sched_n); it.next(tile, kb0, kb1); ++iter) {
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BKT / UMMA_K; ++k) {
            // 16 tokens = 16 swizzled rows of 128 B = 2048 B further into each 64-channel box
            const uint64_t da = make_desc(sa + k * UMMA_K * 128);
            const uint64_t db = make_desc(sb + k * UMMA_K * 128);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (kb == kb1 - 1) umma_commit(&tfull[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> swizzled smem -> TMA reduce-add into H
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    uint8_t* obuf = smem_out + (warp - 2) * 2 * OUT_BUF_BYTES;
    int iter = 0, nstore = 0;
    int tile, kb0, kb1;
    for (WorkIter it(a, sched_cta, sched_n); it.next(tile, kb0, kb1); ++iter) {
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      int m0, n0;
      tile_coords(a, tile, m0, n0);
      mbar_wait(&tfull[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;");
      const int row0 = m0 + q * 32;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        // skip boxes outside the matrix and (symmetric mode) boxes entirely below the diagonal
        const bool live = row0 < a.k && col0 < a.k && !(a.upper_only && col0 + 31 < row0);
        if (!live) continue;  // warp-uniform
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), v);
        uint8_t* buf = obuf + (nstore & 1) * OUT_BUF_BYTES;
        if (nstore >= 2) {  // the store issued two boxes ago has finished reading this buffer
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
        // row `lane` of the box, 128 B, 16-byte chunks XOR-swizzled with (row & 7) (SWIZZLE_128B)
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
          tma_reduce_add_2d(&map_h, buf, col0, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++nstore;
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
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

// ---------------------------------------------------------------------------------------------------
// CTA-pair version (cta_group::2): two CTAs of a cluster compute one 256 x 256 tile of H.  Each CTA stages
// its own 128 rows of A and HALF of B (128 channels) -- 32 KB per 64-token stage instead of 48 KB -- and the
// leader issues tcgen05.mma.cta_group::2 (M = 256) that reads both CTAs' shared memory; each CTA ends up
// with its 128 x 256 half of the accumulator in its own TMEM and runs its own epilogue.  ncu on the
// single-CTA kernel: tensor pipe 48 % active, bounded by operand traffic into / out of shared memory
// (12 KB read per 128-cycle MMA + 48 KB of TMA writes per 512 cycles against 128 B/clk); the pair halves
// the B traffic per SM, which is what cuBLAS' 2-SM kernels do to reach the measured 1.4-1.6 PFLOP/s.
constexpr int PM = 128;                                   // rows of the tile per CTA
constexpr int P_STAGES = 6;
constexpr int PA_BYTES = (PM / BOX_C) * BOX_BYTES;        // 16384
constexpr int PB_BYTES = (BN / 2 / BOX_C) * BOX_BYTES;    // 16384: this CTA's half of B
constexpr int P_STAGE_BYTES = PA_BYTES + PB_BYTES;
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + OUT_BYTES + 256 + 1024;

constexpr uint32_t make_idesc_pair() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)((2 * PM) >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
hessian_umma_pair_kernel(const __grid_constant__ XMaps maps_a, const __grid_constant__ XMaps maps_b,
                         const __grid_constant__ CUtensorMap map_h, const __grid_constant__ HessArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + P_STAGES * PA_BYTES;
  uint8_t* smem_out = smem + P_STAGES * P_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + OUT_BYTES);
  uint64_t* full = bars;                         // [P_STAGES]  TMA (both CTAs) -> MMA; the leader's copy is used
  uint64_t* empty = bars + P_STAGES;             // [P_STAGES]  MMA -> TMA, signalled in both CTAs
  uint64_t* tfull = bars + 2 * P_STAGES;         // [2]         MMA -> epilogue, signalled in both CTAs
  uint64_t* tempty = bars + 2 * P_STAGES + 2;    // [2]         epilogues of both CTAs -> MMA; the leader's copy is used
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int sched_cta = (int)cluster, sched_n = (int)nclusters;  // see WorkIter

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps_a.m[0])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps_b.m[0])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_h)) : "memory");
    for (int i = 0; i < P_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): my 128 rows of A, my 128 channels of B
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int tile, kb0, kb1;
      for (WorkIter it(a, sched_cta, sched_n); it.next(tile, kb0, kb1);) {
        int m0, n0;
        tile_coords(a, tile, m0, n0, 2 * PM);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int smp = kb / a.kb_per_sample, kbl = kb - smp * a.kb_per_sample;
          mbar_wait(&empty[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * P_STAGE_BYTES);  // both CTAs' boxes land on the leader's barrier
          const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
          uint8_t* sa = smem_a + stage * PA_BYTES;
          uint8_t* sb = smem_b + stage * PB_BYTES;
#pragma unroll
          for (int h = 0; h < PM / BOX_C; ++h)
            tma_load_2d_pair(&maps_a.m[smp], lead_full, sa + h * BOX_BYTES, m0 + (int)rank * PM + h * BOX_C, kbl * BKT);
#pragma unroll
          for (int h = 0; h < BN / 2 / BOX_C; ++h)
            tma_load_2d_pair(&maps_b.m[smp], lead_full, sb + h * BOX_BYTES, n0 + (int)rank * (BN / 2) + h * BOX_C, kbl * BKT);
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: leader CTA only
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_pair();
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      int tile, kb0, kb1;
      for (WorkIter it(a, sched_cta, sched_n); it.next(tile, kb0, kb1); ++iter) {
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t sa = smem_u32(smem_a + stage * PA_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * PB_BYTES);
#pragma unroll
          for (int k = 0; k < BKT / UMMA_K; ++k) {
            const uint64_t da = make_desc(sa + k * UMMA_K * 128);
            const uint64_t db = make_desc(sb + k * UMMA_K * 128);
            umma_bf16_pair(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty[stage], 3);  // frees the slot in both CTAs
          if (kb == kb1 - 1) umma_commit_pair(&tfull[as], 3);
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): my 128 x 256 half of the tile
    const int q = warp & 3;
    uint8_t* obuf = smem_out + (warp - 2) * 2 * OUT_BUF_BYTES;
    int iter = 0, nstore = 0;
    int tile, kb0, kb1;
    for (WorkIter it(a, sched_cta, sched_n); it.next(tile, kb0, kb1); ++iter) {
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      int m0, n0;
      tile_coords(a, tile, m0, n0, 2 * PM);
      mbar_wait(&tfull[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;");
      const int row0 = m0 + (int)rank * PM + q * 32;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        const bool live = row0 < a.k && col0 < a.k && !(a.upper_only && col0 + 31 < row0);
        if (!live) continue;  // warp-uniform
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), v);
        uint8_t* buf = obuf + (nstore & 1) * OUT_BUF_BYTES;
        if (nstore >= 2) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
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
          tma_reduce_add_2d(&map_h, buf, col0, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++nstore;
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));  // the leader's MMA warp waits for 8 warps
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  // nobody leaves while the partner may still signal its barriers or read its shared memory
  asm volatile("tcgen05.fence::before_thread_sync;");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// dX split for GPTAQ: d = X_fp - X (exact in fp32 for bf16 inputs of similar magnitude),
// hi = bf16(d), lo = bf16(d - hi)  ->  d^T X = hi^T X + lo^T X with ~16 mantissa bits of d.
__global__ void dx_split_kernel(const __nv_bfloat16* __restrict__ xfp, const __nv_bfloat16* __restrict__ x,
                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = __fsub_rn(__bfloat162float(xfp[i]), __bfloat162float(x[i]));
    const __nv_bfloat16 h = __float2bfloat16_rn(d);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(__fsub_rn(d, __bfloat162float(h)));
  }
}

__global__ void scale_kernel(float* __restrict__ p, int64_t n, float s) {
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = p4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    p4[i] = v;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] *= s;
}

// H[i][j] = H[j][i] = scale * S[min(i,j)][max(i,j)]   (S: upper-triangle raw sums, in place)
__global__ void finalize_sym_kernel(float* __restrict__ H, int64_t k, float scale) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;  // one CTA per upper 32x32 block
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = (int64_t)bi * 32 + r, j = (int64_t)bj * 32 + tx;
    float v = 0.f;
    if (i < k && j < k) {
      v = scale * H[i * k + j];
      if (bi != bj || tx >= r) H[i * k + j] = v;
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // mirrored element (j, i): row bj*32 + r, column bi*32 + tx  <- tile[tx][r]
    const int64_t jj = (int64_t)bj * 32 + r, ii = (int64_t)bi * 32 + tx;
    if (jj < k && ii < k && (bi != bj ? true : tx < r)) H[jj * k + ii] = tile[tx][r];
  }
}

// ---- packed upper triangle (token-sharded ranks all-reduce HALF the bytes of the raw sums) ----
// The upper 32 x 32 blocks of S (block row bi, block column bj >= bi; nb = ceil(k / 32) per side) are laid end to end,
// row of blocks after row of blocks: block (bi, bj) starts at float 1024 * (bi * nb - bi * (bi - 1) / 2 + bj - bi) and is
// stored row-major (32 floats = one 128-byte line per row; elements beyond k are zero).  One CTA per upper block.
__device__ __forceinline__ int64_t packed_block_offset(int bi, int bj, int nb) {
  return 1024 * ((int64_t)bi * nb - (int64_t)bi * (bi - 1) / 2 + (bj - bi));
}

__global__ void __launch_bounds__(256) pack_upper_kernel(const float* __restrict__ S, int64_t k, float* __restrict__ packed) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  float* dst = packed + packed_block_offset(bi, bj, (int)gridDim.x);
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = (int64_t)bi * 32 + r, j = (int64_t)bj * 32 + tx;
    dst[r * 32 + tx] = (i < k && j < k) ? S[i * k + j] : 0.f;
  }
}

// H[i][j] = H[j][i] = scale * packed(min(i,j), max(i,j)): finalize_sym_kernel reading the packed sums
__global__ void __launch_bounds__(256) finalize_packed_kernel(const float* __restrict__ packed, float* __restrict__ H,
                                                              int64_t k, float scale) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = packed + packed_block_offset(bi, bj, (int)gridDim.x);
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = (int64_t)bi * 32 + r, j = (int64_t)bj * 32 + tx;
    const float v = scale * src[r * 32 + tx];
    if (i < k && j < k && (bi != bj || tx >= r)) H[i * k + j] = v;
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t jj = (int64_t)bj * 32 + r, ii = (int64_t)bi * 32 + tx;
    if (jj < k && ii < k && (bi != bj ? true : tx < r)) H[jj * k + ii] = tile[tx][r];
  }
}

// s[k] = beta * s[k] + alpha * sum_t x[t][k]^2      (32 x 8 threads: 64 channels per CTA)
__global__ void __launch_bounds__(256) rownorm_kernel(float* __restrict__ s, const __nv_bfloat16* __restrict__ x,
                                                      int64_t tokens, int64_t k, float alpha, float beta) {
  __shared__ float red[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 64 + tx * 2;
  float a0 = 0.f, a1 = 0.f;
  if (c < k) {  // k is even
    for (int64_t t = ty; t < tokens; t += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + t * k + c);
      const float f0 = __bfloat162float(v.x), f1 = __bfloat162float(v.y);
      a0 = fmaf(f0, f0, a0); a1 = fmaf(f1, f1, a1);
    }
  }
  red[ty][tx * 2] = a0; red[ty][tx * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int64_t cc = (int64_t)blockIdx.x * 64 + threadIdx.x;
    if (cc < k) {
      float acc = 0.f;
      for (int i = 0; i < 8; ++i) acc += red[i][threadIdx.x];
      s[cc] = (beta != 0.0f ? beta * s[cc] : 0.0f) + alpha * acc;
    }
  }
}

// [tokens, k] bf16 row-major -> 2D map, dim0 = channels (contiguous), dim1 = tokens, 64 x 64 boxes
// LCB_HESSIAN_SCHED=0 selects the plain stream-K schedule (A/B measurements); default: lock-step whole tiles
int hessian_sched_mode() {
  static const int mode = [] {
    const char* e = std::getenv("LCB_HESSIAN_SCHED");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();
  return mode;
}

int make_x_map(CUtensorMap* map, const void* x, int64_t tokens, int64_t k) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc != LCB_OK) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)tokens};
  cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  cuuint32_t box[2] = {BOX_C, BKT};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (x=%p tokens=%lld k=%lld)", (int)r, x, (long long)tokens,
              (long long)k);
    return LCB_ERR_CUDA;
  }
  return LCB_OK;
}

// [k, k] fp32 row-major -> 2D map with 32 x 32 boxes (128 B rows, 128B swizzle) for the reduce-add
int make_h_map(CUtensorMap* map, float* h, int64_t k) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc != LCB_OK) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)k};
  cuuint64_t strides[1] = {(cuuint64_t)k * 4};
  cuuint32_t box[2] = {OUT_BOX, OUT_BOX};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (H) failed with CUresult %d (h=%p k=%lld)", (int)r, (void*)h, (long long)k);
    return LCB_ERR_CUDA;
  }
  return LCB_OK;
}

// out += alpha * sum_s A_s^T B_s over the tokens of `count` hook inputs (each [tokens, k]); upper_only: only tiles
// touching the upper triangle.  One launch, one fp32 accumulation chain per tile over all count * tokens tokens.
int launch_xtx_multi(float* out, const void* const* a_src, const void* const* b_src, int count, int64_t tokens, int64_t k,
                     float alpha, int upper_only, cudaStream_t st) {
  XMaps maps_a, maps_b;
  CUtensorMap map_h;
  int rc;
  for (int i = 0; i < MAX_SAMPLES; ++i) {
    if (i >= count) {  // unused slots repeat sample 0 (never dereferenced)
      maps_a.m[i] = maps_a.m[0];
      maps_b.m[i] = maps_b.m[0];
      continue;
    }
    if ((rc = make_x_map(&maps_a.m[i], a_src[i], tokens, k)) != LCB_OK) return rc;
    if (b_src[i] == a_src[i]) maps_b.m[i] = maps_a.m[i];
    else if ((rc = make_x_map(&maps_b.m[i], b_src[i], tokens, k)) != LCB_OK) return rc;
  }
  rc = make_h_map(&map_h, out, k);
  if (rc != LCB_OK) return rc;
  HessArgs a{};
  a.k = k; a.tokens = tokens; a.alpha = alpha; a.upper_only = upper_only;
  a.sched = hessian_sched_mode();
  const bool pair = hessian_pair_mode(k);
  const int bm = pair ? 2 * PM : BM;
  a.tiles_m = (int)ceil_div(k, bm); a.tiles_n = (int)ceil_div(k, BN);
  if (a.tiles_m > MAX_TILES_M) {
    set_error("lcb_hessian_accum: k = %lld is larger than the supported 16384", (long long)k);
    return LCB_ERR_UNSUPPORTED;
  }
  int acc = 0;
  for (int mt = 0; mt < a.tiles_m; ++mt) {
    a.row_start[mt] = (int16_t)acc;
    const int first_nt = upper_only ? (mt * bm) / BN : 0;
    acc += a.tiles_n > first_nt ? a.tiles_n - first_nt : 0;
  }
  a.row_start[a.tiles_m] = (int16_t)acc;
  a.num_tiles = acc;
  a.kb_per_sample = (int)ceil_div(tokens, BKT);
  a.kblocks = count * a.kb_per_sample;
  const int64_t units = (int64_t)a.num_tiles * a.kblocks;
  if (pair) {
    LCB_CUDA(cudaFuncSetAttribute(hessian_umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
    // one CTA pair per TPC; never less than ~4 token blocks per pair
    int clusters = sm_count() / 2;
    const int64_t cap = units / 4 > 0 ? units / 4 : 1;
    if (clusters > cap) clusters = (int)cap;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * clusters), 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = P_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    LCB_CUDA(cudaLaunchKernelEx(&cfg, hessian_umma_pair_kernel, maps_a, maps_b, map_h, a));
    return LCB_OK;
  }
  LCB_CUDA(cudaFuncSetAttribute(hessian_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  // one CTA per SM; never less than ~8 token blocks per CTA so the epilogue stays amortised
  int grid = sm_count();
  const int64_t cap = units / 8 > 0 ? units / 8 : 1;
  if (grid > cap) grid = (int)cap;
  hessian_umma_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps_a, maps_b, map_h, a);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

int launch_xtx(float* out, const void* a_src, const void* b_src, int64_t tokens, int64_t k, float alpha, int upper_only,
               cudaStream_t st) {
  return launch_xtx_multi(out, &a_src, &b_src, 1, tokens, k, alpha, upper_only, st);
}

int scale_inplace(float* p, int64_t n, float s, cudaStream_t st) {
  int64_t g = ceil_div(n, 256 * 16);
  if (g > (int64_t)sm_count() * 8) g = (int64_t)sm_count() * 8;
  scale_kernel<<<(unsigned)(g < 1 ? 1 : g), 256, 0, st>>>(p, n, s);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

}  // namespace

}  // namespace lcb

using namespace lcb;

extern "C" size_t lcb_hessian_ws_bytes(int64_t tokens, int64_t k) { return (size_t)(tokens * k) * 2 * 2 + 512; }

extern "C" int lcb_hessian_accum_multi(float* H, const void* const* xs, int count, int64_t tokens, int64_t k, float alpha,
                                       int upper_only, void* stream) {
  LCB_REQUIRE(H != nullptr && xs != nullptr && count >= 1 && count <= MAX_SAMPLES, "lcb_hessian_accum_multi: bad arguments");
  LCB_REQUIRE(tokens > 0 && k > 0 && k % 8 == 0, "lcb_hessian_accum_multi: need tokens > 0 and k a positive multiple of 8");
  LCB_REQUIRE((reinterpret_cast<uintptr_t>(H) & 15) == 0, "lcb_hessian_accum_multi: H must be 16-byte aligned");
  for (int i = 0; i < count; ++i)
    LCB_REQUIRE(xs[i] != nullptr && (reinterpret_cast<uintptr_t>(xs[i]) & 15) == 0,
                "lcb_hessian_accum_multi: every x must be a 16-byte aligned device pointer");
  return launch_xtx_multi(H, xs, xs, count, tokens, k, alpha, upper_only, static_cast<cudaStream_t>(stream));
}

extern "C" int lcb_hessian_accum(float* H, float* dxxt, const void* x, const void* x_fp, int64_t tokens, int64_t k,
                                 float alpha, float beta, int upper_only, void* ws, size_t ws_bytes, void* stream) {
  LCB_REQUIRE(H != nullptr && x != nullptr, "lcb_hessian_accum: NULL pointer");
  LCB_REQUIRE(tokens > 0 && k > 0 && k % 8 == 0, "lcb_hessian_accum: need tokens > 0 and k a positive multiple of 8");
  LCB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0,
              "lcb_hessian_accum: x and H must be 16-byte aligned");
  LCB_REQUIRE((dxxt == nullptr) == (x_fp == nullptr), "lcb_hessian_accum: dxxt and x_fp go together");
  LCB_REQUIRE(!upper_only || beta == 1.0f, "lcb_hessian_accum: upper_only accumulates raw sums (beta must be 1)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (beta != 1.0f) {  // running-mean form: scale first, every contribution below is an L2 reduce-add
    rc = scale_inplace(H, k * k, beta, st);
    if (rc != LCB_OK) return rc;
  }
  rc = launch_xtx(H, x, x, tokens, k, alpha, upper_only, st);
  if (rc != LCB_OK) return rc;
  if (dxxt != nullptr) {
    if (ws == nullptr || ws_bytes < lcb_hessian_ws_bytes(tokens, k)) {
      set_error("lcb_hessian_accum: workspace of %zu bytes needed for the dXXT term", lcb_hessian_ws_bytes(tokens, k));
      return LCB_ERR_WORKSPACE;
    }
    LCB_REQUIRE((reinterpret_cast<uintptr_t>(x_fp) & 15) == 0 && (reinterpret_cast<uintptr_t>(dxxt) & 15) == 0,
                "lcb_hessian_accum: x_fp and dxxt must be 16-byte aligned");
    uintptr_t wsa = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(wsa);
    __nv_bfloat16* lo = hi + ((tokens * k + 127) / 128) * 128;
    const int64_t n = tokens * k;
    int64_t g = ceil_div(n, 256 * 4);
    if (g > (int64_t)sm_count() * 8) g = (int64_t)sm_count() * 8;
    dx_split_kernel<<<(unsigned)g, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x_fp),
                                                  static_cast<const __nv_bfloat16*>(x), hi, lo, n);
    LCB_LAUNCH_CHECK();
    if (beta != 1.0f) {
      rc = scale_inplace(dxxt, k * k, beta, st);
      if (rc != LCB_OK) return rc;
    }
    rc = launch_xtx(dxxt, hi, x, tokens, k, alpha, 0, st);
    if (rc != LCB_OK) return rc;
    rc = launch_xtx(dxxt, lo, x, tokens, k, alpha, 0, st);
    if (rc != LCB_OK) return rc;
  }
  return LCB_OK;
}

extern "C" int lcb_hessian_finalize(float* H, int64_t k, float scale, int symmetric_from_upper, void* stream) {
  LCB_REQUIRE(H != nullptr && k > 0, "lcb_hessian_finalize: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!symmetric_from_upper) return scale_inplace(H, k * k, scale, st);
  dim3 grid((unsigned)ceil_div(k, 32), (unsigned)ceil_div(k, 32));
  finalize_sym_kernel<<<grid, 256, 0, st>>>(H, k, scale);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" size_t lcb_hessian_packed_floats(int64_t k) {
  if (k <= 0) return 0;
  const size_t nb = (size_t)ceil_div(k, 32);
  return nb * (nb + 1) / 2 * 1024;
}

extern "C" int lcb_hessian_pack_upper(const float* S, int64_t k, float* packed, void* stream) {
  LCB_REQUIRE(S != nullptr && packed != nullptr && k > 0 && k <= 16384 * 4, "lcb_hessian_pack_upper: bad arguments");
  dim3 grid((unsigned)ceil_div(k, 32), (unsigned)ceil_div(k, 32));
  pack_upper_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, k, packed);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_hessian_finalize_packed(const float* packed, float* H, int64_t k, float scale, void* stream) {
  LCB_REQUIRE(H != nullptr && packed != nullptr && k > 0 && k <= 16384 * 4, "lcb_hessian_finalize_packed: bad arguments");
  dim3 grid((unsigned)ceil_div(k, 32), (unsigned)ceil_div(k, 32));
  finalize_packed_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, H, k, scale);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_rownorm_accum(float* s, const void* x, int64_t tokens, int64_t k, float alpha, float beta,
                                 void* stream) {
  LCB_REQUIRE(s != nullptr && x != nullptr && tokens > 0 && k > 0 && k % 2 == 0, "lcb_rownorm_accum: bad arguments");
  rownorm_kernel<<<(unsigned)ceil_div(k, 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      s, static_cast<const __nv_bfloat16*>(x), tokens, k, alpha, beta);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

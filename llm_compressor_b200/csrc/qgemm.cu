// (f3) Fused activation fake-quant + GEMM for the calibration forwards:
//
//     y[M, N] = QDQ(x)[M, K] @ W[N, K]^T (+ bias)        ref: modules/qlinear.py:86-88
//
// The reference quantise-dequantises the activation into a new tensor (write + read of M*K bf16 per Linear and forward,
// ~3.9 TB over a Llama-3.2-3B GPTQ run with --act-in) and then calls F.linear.  Here the QDQ happens in the
// operand prologue of a tcgen05 GEMM: TMA lays the raw bf16 activation tile into shared memory (K-major, 128-byte
// swizzle), eight TRANSFORM warps quantise-dequantise it in place with the row's scale / zero point -- the very packed
// bf16 arithmetic of the streaming fake-quant kernel (apply16, qdq_stream.cuh), so the operand is bit-identical to
// input_quantizer(x) -- and hand the stage to the MMA warp; the quantised activation never exists in HBM.
// Scales / zero points come from one find-only pass over x (lcb_qdq, mode FIND): per token (one per row) or per group of
// `group` columns along K (group % 64 == 0).
// Warp roles: 0 TMA producer, 1 MMA issuer (+ TMEM owner), 2..9 transform, 10..13 epilogue (tcgen05.ld -> bias -> bf16 ->
// global).  Tiles 128 x 256 x 64, 4 stages, two TMEM accumulators (the epilogue of tile i overlaps the MMAs of tile i+1).
#include "qdq_apply.cuh"
#include "umma.cuh"

namespace lcb {
namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;      // 16 KB
constexpr int B_BYTES = BN * BK * 2;      // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int NUM_THREADS = 14 * 32;
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// c_format F32, a / b BF16, both K-major, N = 256, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct QgArgs {
  const __nv_bfloat16* scales;   // [M, G] or null (no activation quantiser)
  const __nv_bfloat16* zeros;
  const __nv_bfloat16* bias;     // [N] or null
  __nv_bfloat16* y;              // [M, N]
  int M, N, K;
  int group;                     // columns per scale along K (K for per-token)
  int G;
  int kind;                      // FK_INT4 / FK_INT8
  int zero_point;
  int tiles_m, tiles_n;
};

template <bool QUANT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
qgemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ QgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                    // TMA landed
  uint64_t* ready = bars + STAGES;          // transformed (QUANT) -- what the MMA warp waits on
  uint64_t* empty = bars + 2 * STAGES;      // MMAs done with the stage
  uint64_t* tfull = bars + 3 * STAGES;      // [2] accumulator complete
  uint64_t* tempty = bars + 3 * STAGES + 2; // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.tiles_m * a.tiles_n;
  const int kblocks = (a.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&ready[i], 256); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t % a.tiles_m) * BM, n0 = (t / a.tiles_m) * BN;   // consecutive CTAs share a W tile
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* s = smem + stage * STAGE_BYTES;
          tma_load_2d(&map_x, &full[stage], s, kb * BK, m0);
          tma_load_2d(&map_w, &full[stage], s + A_BYTES, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0, iter = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int as = iter & 1;
        mbar_wait(&tempty[as], ((iter >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(QUANT ? &ready[stage] : &full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_d, desc_k128(sa + k * 32), desc_k128(sb + k * 32), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (kb == kblocks - 1) umma_commit(&tfull[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ++iter;
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ transform: QDQ of the activation tile in place
    if constexpr (QUANT) {
      const int tt = threadIdx.x - 64;          // 0 .. 255
      const int r = tt >> 1, h = tt & 1;        // row of the tile, which 32 of its 64 columns
      int stage = 0;
      uint32_t phase = 0;
      const bool zp = a.zero_point != 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t % a.tiles_m) * BM;
        const int row = min(m0 + r, a.M - 1);
        int gcur = -1;
        float s = 1.0f, z = 0.0f;
        for (int kb = 0; kb < kblocks; ++kb) {
          const int g = (kb * BK) / a.group;
          if (g != gcur) {
            gcur = g;
            s = __bfloat162float(a.scales[(int64_t)row * a.G + g]);
            z = __bfloat162float(a.zeros[(int64_t)row * a.G + g]);
          }
          mbar_wait(&full[stage], phase);
          uint8_t* rowp = smem + stage * STAGE_BYTES + r * 128;
          const bool finite = (__float_as_uint(s) & 0x7f800000u) != 0x7f800000u && (__float_as_uint(z) & 0x7f800000u) != 0x7f800000u;
#pragma unroll
          for (int c = 0; c < 2; ++c) {          // two chunks of 16 elements = pieces 4h+2c, 4h+2c+1
            const int p0 = 4 * h + 2 * c;
            uint4* q0 = reinterpret_cast<uint4*>(rowp + (((p0) ^ (r & 7)) << 4));
            uint4* q1 = reinterpret_cast<uint4*>(rowp + (((p0 + 1) ^ (r & 7)) << 4));
            const uint4 u0 = *q0, u1 = *q1;
            W8 v;
            v.w[0] = u0.x; v.w[1] = u0.y; v.w[2] = u0.z; v.w[3] = u0.w;
            v.w[4] = u1.x; v.w[5] = u1.y; v.w[6] = u1.z; v.w[7] = u1.w;
            if (!finite) {
              QCfg c{};
              c.qtype = LCB_Q_INT; c.zero_point = a.zero_point; c.f = make_fmt(a.kind == FK_INT4 ? LCB_E_INT4 : LCB_E_INT8);
              v = apply16_generic(c, v, s, z);
            } else if (a.kind == FK_INT4) {
              apply16<FK_INT4>(v, s, z, zp);
            } else {
              apply16<FK_INT8>(v, s, z, zp);
            }
            *q0 = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
            *q1 = make_uint4(v.w[4], v.w[5], v.w[6], v.w[7]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&ready[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;    // TMEM lane quarter of this warp (warps 10..13 -> 2, 3, 0, 1)
    int iter = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t % a.tiles_m) * BM, n0 = (t / a.tiles_m) * BN;
      const int as = iter & 1;
      mbar_wait(&tfull[as], (iter >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;");
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= a.N) break;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), v);
        if (row < a.M) {
          __nv_bfloat16* yp = a.y + (int64_t)row * a.N + col0;
          if (col0 + 32 <= a.N && (a.N % 8) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float f0 = v[8 * j + 2 * e], f1 = v[8 * j + 2 * e + 1];
                if (a.bias) {   // F.linear adds the bias in fp32 before the single rounding to bf16 (cuBLAS epilogue)
                  f0 += __bfloat162float(a.bias[col0 + 8 * j + 2 * e]);
                  f1 += __bfloat162float(a.bias[col0 + 8 * j + 2 * e + 1]);
                }
                w[e] = pack_bf2(f0, f1);
              }
              *reinterpret_cast<uint4*>(yp + 8 * j) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < a.N; ++j) {
              float f = v[j];
              if (a.bias) f += __bfloat162float(a.bias[col0 + j]);
              yp[j] = __float2bfloat16_rn(f);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      ++iter;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

int make_map_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_cols, int box_rows) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc != LCB_OK) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 %lld x %lld) failed with CUresult %d", (long long)rows, (long long)cols, (int)r);
    return LCB_ERR_CUDA;
  }
  return LCB_OK;
}

}  // namespace
}  // namespace lcb

using namespace lcb;

extern "C" int lcb_qlinear_fwd(const lcb_quant_cfg* cfg, const void* x, const void* w, const void* bias, void* y, int64_t m,
                               int64_t n, int64_t k, int64_t group, const void* scales, const void* zeros, void* stream) {
  LCB_REQUIRE(x && w && y && m > 0 && n > 0 && k > 0, "lcb_qlinear_fwd: bad arguments");
  LCB_REQUIRE(k % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0,
              "lcb_qlinear_fwd: k must be a multiple of 8 and x / w 16-byte aligned (TMA)");
  const bool quant = cfg != nullptr;
  QgArgs a{};
  if (quant) {
    LCB_REQUIRE(cfg->qtype == LCB_Q_INT && (cfg->elem == LCB_E_INT4 || cfg->elem == LCB_E_INT8),
                "lcb_qlinear_fwd: fused activation quantiser must be INT4 / INT8 (others: lcb_qdq, then cfg = NULL)");
    LCB_REQUIRE(scales && zeros && group > 0 && group % BK == 0 && k % group == 0,
                "lcb_qlinear_fwd: group must be a multiple of 64 that divides k (per token: group = k), with scales / zeros");
    a.kind = cfg->elem == LCB_E_INT4 ? FK_INT4 : FK_INT8;
    a.zero_point = cfg->zero_point ? 1 : 0;
    a.group = (int)group; a.G = (int)(k / group);
    a.scales = static_cast<const __nv_bfloat16*>(scales); a.zeros = static_cast<const __nv_bfloat16*>(zeros);
  } else {
    a.group = (int)k; a.G = 1;
  }
  a.bias = static_cast<const __nv_bfloat16*>(bias);
  a.y = static_cast<__nv_bfloat16*>(y);
  a.M = (int)m; a.N = (int)n; a.K = (int)k;
  a.tiles_m = (int)ceil_div(m, BM); a.tiles_n = (int)ceil_div(n, BN);
  CUtensorMap mx, mw;
  int rc;
  if ((rc = make_map_bf16(&mx, x, m, k, BK, BM)) != LCB_OK) return rc;
  if ((rc = make_map_bf16(&mw, w, n, k, BK, BN)) != LCB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tiles = a.tiles_m * a.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (quant) {
    LCB_CUDA(cudaFuncSetAttribute(qgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    qgemm_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mw, a);
  } else {
    LCB_CUDA(cudaFuncSetAttribute(qgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    qgemm_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mw, a);
  }
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

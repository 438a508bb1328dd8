// (c) Fused quantize-dequantize kernels for sm_100a.  HBM-bound streaming work: 128-bit
// coalesced loads/stores, sub-warp shuffle reductions for the group statistics, one pass over
// HBM wherever the reference's granularity allows it (group / token), two passes where a
// tensor-wide or column-wide statistic must exist first (per-tensor, per-channel, NVFP amax).
//
// Replaces ref: quantizers/{int,fp,mx,nvfp}_quant.py find_params/forward/fake_quantize and
// quantizers/utils.py _reshape_to_blocks/_undo_reshape_to_blocks/_quantize_elemwise_core.
#include <algorithm>
#include <type_traits>

#include "qdq_fast.cuh"
#include "qmath.cuh"

namespace lcb {

struct QdqArgs {
  const void* x;
  void* out;
  void* scales;
  void* zeros;
  uint8_t* codes;
  const float* nv_amax;  // device float: whole-tensor amax for NVFP (already available)
  uint32_t* status;
  int64_t nrows;  // batch * rows (axis -1) or batch (axis -2)
  int64_t rows;   // axis -2: rows per batch
  int64_t cols;
  int64_t group;
  int64_t G;  // groups per row (axis -1) / per column (axis -2)
  int find, apply;
  int mse;  // find_params with the mse clip search
  QCfg c;
};

__device__ __forceinline__ void flag_nan_scale(const QdqArgs& a, float s) {
  if (a.status != nullptr && s != s) atomicOr(a.status, LCB_ST_NAN_SCALE);
}

}  // namespace lcb
#include "qdq_stream.cuh"
namespace lcb {

template <typename T, int N>
__device__ __forceinline__ void stats_of(const float (&v)[N], float& mx, float& mn, float& amax) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    mx = nan_max(mx, v[i]);
    mn = nan_min(mn, v[i]);
    amax = nan_max(amax, fabsf(v[i]));
  }
}

template <typename T, int N>
__device__ __forceinline__ void apply_vec(const QdqArgs& a, const float (&v)[N], float s, float z, T* outp,
                                          uint8_t* codep) {
  constexpr int DT = DtOf<T>::value;
  float o[N];
  uint8_t cd[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float code;
    o[i] = fake_quant<DT>(a.c, v[i], s, z, code);
    if (codep != nullptr) cd[i] = encode_code(a.c, code);
  }
  store16<T>(outp, o);
  if (codep != nullptr) {
    if constexpr (N == 8) {
      uint2 p;
      p.x = cd[0] | (cd[1] << 8) | (cd[2] << 16) | (cd[3] << 24);
      p.y = cd[4] | (cd[5] << 8) | (cd[6] << 16) | (cd[7] << 24);
      *reinterpret_cast<uint2*>(codep) = p;
    } else {
      *reinterpret_cast<uint32_t*>(codep) = cd[0] | (cd[1] << 8) | (cd[2] << 16) | (cd[3] << 24);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Kernel A: axis -1, group = LPG * VEC elements handled by LPG lanes of one warp (group 16..256
// for bf16, 4..128 for fp32).  One pass: load 16 B per lane, shuffle-reduce, quantise, store.
// NVFP_PASS1: only reduce the block statistic to the tensor-wide amax (atomicMax on ws word).
template <typename T, int LPG, bool NVFP_PASS1>
__global__ void __launch_bounds__(256) qdq_subwarp_kernel(QdqArgs a, uint32_t* amax_key) {
  constexpr int DT = DtOf<T>::value;
  constexpr int VEC = 16 / sizeof(T);
  constexpr int GPW = 32 / LPG;  // groups per warp and step
  constexpr int U = 4;           // steps in flight per warp: four 16-byte loads per lane before the first reduction
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPG, sl = lane % LPG;
  const int64_t total = a.nrows * a.G;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const T* x = static_cast<const T*>(a.x);
  T* out = static_cast<T*>(a.out);
  T* sc = static_cast<T*>(a.scales);
  T* zr = static_cast<T*>(a.zeros);
  float nv_g = 0.0f;
  if (!NVFP_PASS1 && a.c.qtype == LCB_Q_NVFP && a.find) nv_g = *a.nv_amax;
  float local_amax = 0.0f;

  for (int64_t g0 = warp * GPW * U; g0 < total; g0 += nwarps * GPW * U) {
    float v[U][VEC];
    int64_t gid[U], off[U];
    bool active[U], inb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      gid[u] = g0 + u * GPW + sub;
      active[u] = gid[u] < total;
      const int64_t r = active[u] ? gid[u] / a.G : 0;
      const int64_t b = active[u] ? gid[u] - r * a.G : 0;
      const int64_t col = b * a.group + (int64_t)sl * VEC;
      inb[u] = active[u] && col < a.cols;
      off[u] = r * a.cols + col;
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[u][i] = 0.0f;
      if (inb[u]) load16<T>(x + off[u], v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s, z;
      if (a.find) {
        float mx = -INFINITY, mn = INFINITY, amax = 0.0f;
        stats_of<T, VEC>(v[u], mx, mn, amax);  // out-of-range lanes contribute the zero padding
#pragma unroll
        for (int o = 1; o < LPG; o <<= 1) {
          mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          amax = nan_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        }
        if constexpr (NVFP_PASS1) {
          float vb, zb;
          nvfp_block_stat<DT>(mx, mn, amax, a.c.zero_point, vb, zb);
          if (active[u]) local_amax = fmaxf(local_amax, fabsf(vb));
          continue;
        }
        find_params<DT, DT>(a.c, mx, mn, amax, nv_g, s, z);
        if (active[u] && sl == 0) {
          flag_nan_scale(a, s);
          if (sc != nullptr) sc[gid[u]] = from_f<T>(s);
          if (zr != nullptr) zr[gid[u]] = from_f<T>(z);
        }
      } else {
        s = active[u] ? to_f<T>(sc[gid[u]]) : 1.0f;
        z = active[u] ? to_f<T>(zr[gid[u]]) : 0.0f;
      }
      if (a.apply && inb[u]) apply_vec<T, VEC>(a, v[u], s, z, out + off[u], a.codes ? a.codes + off[u] : nullptr);
    }
  }
  if constexpr (NVFP_PASS1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_amax = fmaxf(local_amax, __shfl_xor_sync(0xffffffffu, local_amax, o));
    if (lane == 0) atomicMax(amax_key, __float_as_uint(local_amax));  // non-negative floats order as uints
  }
}

// ------------------------------------------------------------------------------------------
// Kernel B: axis -1, one CTA per group for long groups (per-token rows up to 256*VEC*UNR
// elements).  The whole group stays in registers between the reduction and the quantisation,
// so HBM is still read once.
template <typename T, int UNR, bool NVFP_PASS1>
__global__ void __launch_bounds__(256) qdq_rowcta_kernel(QdqArgs a, uint32_t* amax_key) {
  constexpr int DT = DtOf<T>::value;
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float red[3][8];
  __shared__ float bc[2];
  const T* x = static_cast<const T*>(a.x);
  T* out = static_cast<T*>(a.out);
  T* sc = static_cast<T*>(a.scales);
  T* zr = static_cast<T*>(a.zeros);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t total = a.nrows * a.G;
  float nv_g = 0.0f;
  if (!NVFP_PASS1 && a.c.qtype == LCB_Q_NVFP && a.find) nv_g = *a.nv_amax;

  for (int64_t gid = blockIdx.x; gid < total; gid += gridDim.x) {
    const int64_t r = gid / a.G, b = gid - r * a.G;
    const int64_t c0 = b * a.group;
    const int64_t glen = min(a.group, a.cols - c0);
    float v[UNR][VEC];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t e = ((int64_t)u * 256 + tid) * VEC;
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[u][i] = 0.0f;
      if (e < glen) load16<T>(x + r * a.cols + c0 + e, v[u]);
    }
    float s, z;
    if (a.find) {
      float mx = -INFINITY, mn = INFINITY, amax = 0.0f;
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t e = ((int64_t)u * 256 + tid) * VEC;
        if (e < a.group) stats_of<T, VEC>(v[u], mx, mn, amax);  // e >= glen inside the group: zero padding
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        amax = nan_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      }
      if (lane == 0) { red[0][wid] = mx; red[1][wid] = mn; red[2][wid] = amax; }
      __syncthreads();
      if (wid == 0) {
        mx = lane < 8 ? red[0][lane] : -INFINITY;
        mn = lane < 8 ? red[1][lane] : INFINITY;
        amax = lane < 8 ? red[2][lane] : 0.0f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          amax = nan_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        }
        if (lane == 0) {
          if constexpr (NVFP_PASS1) {
            float vb, zb;
            nvfp_block_stat<DT>(mx, mn, amax, a.c.zero_point, vb, zb);
            atomicMax(amax_key, __float_as_uint(fabsf(vb)));
          } else {
            find_params<DT, DT>(a.c, mx, mn, amax, nv_g, s, z);
            flag_nan_scale(a, s);
            if (sc != nullptr) sc[gid] = from_f<T>(s);
            if (zr != nullptr) zr[gid] = from_f<T>(z);
            bc[0] = s; bc[1] = z;
          }
        }
      }
      __syncthreads();
      if constexpr (NVFP_PASS1) continue;
      s = bc[0]; z = bc[1];
    } else {
      s = to_f<T>(sc[gid]);
      z = to_f<T>(zr[gid]);
    }
    if (a.apply) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t e = ((int64_t)u * 256 + tid) * VEC;
        if (e < glen) {
          const int64_t off = r * a.cols + c0 + e;
          apply_vec<T, VEC>(a, v[u], s, z, out + off, a.codes ? a.codes + off : nullptr);
        }
      }
    }
    __syncthreads();  // bc / red reuse
  }
}


// ------------------------------------------------------------------------------------------
// Fast variants of kernels A and B for bf16 tensors without code output (qdq_fast.cuh): same
// results, ~5x fewer instructions per element, two independent 16 B loads in flight per lane.
template <int KIND, int LPG>
__global__ void __launch_bounds__(256) qdq_subwarp_fast_kernel(QdqArgs a) {
  constexpr int VEC = 8;
  constexpr int GPW = 32 / LPG;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPG, sl = lane % LPG;
  const int64_t total = a.nrows * a.G;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
  __nv_bfloat16* sc = static_cast<__nv_bfloat16*>(a.scales);
  __nv_bfloat16* zr = static_cast<__nv_bfloat16*>(a.zeros);
  const bool zp = a.c.zero_point != 0;
  const bool even = a.cols == a.G * a.group;
  float nv_g = 0.0f;
  if (a.c.qtype == LCB_Q_NVFP && a.find) nv_g = *a.nv_amax;

  for (int64_t g0 = warp * GPW * 2; g0 < total; g0 += nwarps * GPW * 2) {
    int64_t gid[2], off[2];
    bool active[2], inb[2];
    uint4 v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      gid[u] = g0 + u * GPW + sub;
      active[u] = gid[u] < total;
      if (even) {  // cols == G * group: groups tile the tensor, no 64-bit division needed
        inb[u] = active[u];
        off[u] = gid[u] * a.group + (int64_t)sl * VEC;
      } else {
        const int64_t r = active[u] ? gid[u] / a.G : 0;
        const int64_t b = active[u] ? gid[u] - r * a.G : 0;
        const int64_t col = b * a.group + (int64_t)sl * VEC;
        inb[u] = active[u] && col < a.cols;
        off[u] = r * a.cols + col;
      }
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (inb[u]) v[u] = *reinterpret_cast<const uint4*>(x + off[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float s, z;
      if (a.find) {
        Stat2 st = stats8(v[u], zp);
#pragma unroll
        for (int o = 1; o < LPG; o <<= 1) stat_shfl_xor(st, o, zp);
        float mx, mn, amax;
        stat_finish(st, zp, mx, mn, amax);
        if constexpr (KIND == FK_INT4 || KIND == FK_INT8) int_params_fast(mx, mn, amax, zp, a.c.f, s, z);
        else find_params<LCB_BF16, LCB_BF16>(a.c, mx, mn, amax, nv_g, s, z);
        if (active[u] && sl == 0) {
          flag_nan_scale(a, s);
          if (sc != nullptr) sc[gid[u]] = __float2bfloat16_rn(s);
          if (zr != nullptr) zr[gid[u]] = __float2bfloat16_rn(z);
        }
      } else {
        s = active[u] ? __bfloat162float(sc[gid[u]]) : 1.0f;
        z = active[u] ? __bfloat162float(zr[gid[u]]) : 0.0f;
      }
      if (a.apply && inb[u]) {
        const uint4 o4 = apply8<KIND>(v[u], s, z, rcp_fast(s), zp);
        *reinterpret_cast<uint4*>(out + off[u]) = o4;
      }
    }
  }
}


// NVFP pass 1 (whole-tensor amax of the block statistics), bf16, groups tiling the tensor.
template <int LPG>
__global__ void __launch_bounds__(256) nvfp_amax_fast_kernel(QdqArgs a, uint32_t* amax_key) {
  constexpr int VEC = 8;
  constexpr int GPW = 32 / LPG;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPG, sl = lane % LPG;
  const int64_t total = a.nrows * a.G;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  const bool zp = a.c.zero_point != 0;
  float local = 0.0f;
  for (int64_t g0 = warp * GPW * 2; g0 < total; g0 += nwarps * GPW * 2) {
    uint4 v[2];
    bool active[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t gid = g0 + u * GPW + sub;
      active[u] = gid < total;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (active[u]) v[u] = *reinterpret_cast<const uint4*>(x + gid * a.group + (int64_t)sl * VEC);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      Stat2 st = stats8(v[u], zp);
#pragma unroll
      for (int o = 1; o < LPG; o <<= 1) stat_shfl_xor(st, o, zp);
      float mx, mn, amax, vb, zb;
      stat_finish(st, zp, mx, mn, amax);
      nvfp_block_stat<LCB_BF16>(mx, mn, amax, a.c.zero_point, vb, zb);
      if (active[u]) local = fmaxf(local, fabsf(vb));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local = fmaxf(local, __shfl_xor_sync(0xffffffffu, local, o));
  if (lane == 0) atomicMax(amax_key, __float_as_uint(local));
}

template <int KIND, int UNR>
__global__ void __launch_bounds__(256) qdq_rowcta_fast_kernel(QdqArgs a) {
  constexpr int VEC = 8;
  __shared__ uint32_t red[2][8];
  __shared__ float bc[2];
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
  __nv_bfloat16* sc = static_cast<__nv_bfloat16*>(a.scales);
  __nv_bfloat16* zr = static_cast<__nv_bfloat16*>(a.zeros);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t total = a.nrows * a.G;
  const bool zp = a.c.zero_point != 0;
  float nv_g = 0.0f;
  if (a.c.qtype == LCB_Q_NVFP && a.find) nv_g = *a.nv_amax;

  for (int64_t gid = blockIdx.x; gid < total; gid += gridDim.x) {
    const int64_t r = gid / a.G, b = gid - r * a.G;
    const int64_t c0 = b * a.group;
    const int64_t glen = min(a.group, a.cols - c0);
    uint4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t e = ((int64_t)u * 256 + tid) * VEC;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (e < glen) v[u] = *reinterpret_cast<const uint4*>(x + r * a.cols + c0 + e);
    }
    float s, z;
    if (a.find) {
      Stat2 st = stats8(v[0], zp);  // thread 0..: e = tid*8 < group always holds for u == 0 when group >= 2048
#pragma unroll
      for (int u = 1; u < UNR; ++u) {
        const int64_t e = ((int64_t)u * 256 + tid) * VEC;
        if (e < a.group) { Stat2 t = stats8(v[u], zp); stat_combine(st, t, zp); }
      }
      // threads entirely beyond a short group hold zeros = the padding value only if the group is ragged;
      // for a group shorter than 2048 elements that is not ragged they must not contribute
      if ((int64_t)tid * VEC >= a.group) { st.mxmn = 0xff80ff80u; st.amax = 0; }  // (-inf, -inf) / 0
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) stat_shfl_xor(st, o, zp);
      if (lane == 0) { red[0][wid] = st.mxmn; red[1][wid] = st.amax; }
      __syncthreads();
      if (wid == 0) {
        Stat2 t;
        t.mxmn = lane < 8 ? red[0][lane] : 0xff80ff80u;
        t.amax = lane < 8 ? red[1][lane] : 0u;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) stat_shfl_xor(t, o, zp);
        if (lane == 0) {
          float mx, mn, amax;
          stat_finish(t, zp, mx, mn, amax);
          if constexpr (KIND == FK_INT4 || KIND == FK_INT8) int_params_fast(mx, mn, amax, zp, a.c.f, s, z);
          else find_params<LCB_BF16, LCB_BF16>(a.c, mx, mn, amax, nv_g, s, z);
          flag_nan_scale(a, s);
          if (sc != nullptr) sc[gid] = __float2bfloat16_rn(s);
          if (zr != nullptr) zr[gid] = __float2bfloat16_rn(z);
          bc[0] = s; bc[1] = z;
        }
      }
      __syncthreads();
      s = bc[0]; z = bc[1];
    } else {
      s = __bfloat162float(sc[gid]);
      z = __bfloat162float(zr[gid]);
    }
    if (a.apply) {
      const float rr = rcp_fast(s);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t e = ((int64_t)u * 256 + tid) * VEC;
        if (e < glen) *reinterpret_cast<uint4*>(out + r * a.cols + c0 + e) = apply8<KIND>(v[u], s, z, rr, zp);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Kernel C: axis -1, any geometry (unaligned rows, odd group sizes, very long groups).  One warp
// per group, scalar accesses, two passes over the group (second pass hits L1/L2).
template <typename T, bool NVFP_PASS1>
__global__ void __launch_bounds__(256) qdq_generic_kernel(QdqArgs a, uint32_t* amax_key) {
  constexpr int DT = DtOf<T>::value;
  const T* x = static_cast<const T*>(a.x);
  T* out = static_cast<T*>(a.out);
  T* sc = static_cast<T*>(a.scales);
  T* zr = static_cast<T*>(a.zeros);
  const int lane = threadIdx.x & 31;
  const int64_t total = a.nrows * a.G;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float nv_g = 0.0f;
  if (!NVFP_PASS1 && a.c.qtype == LCB_Q_NVFP && a.find) nv_g = *a.nv_amax;
  for (int64_t gid = warp; gid < total; gid += nwarps) {
    const int64_t r = gid / a.G, b = gid - r * a.G;
    const int64_t c0 = b * a.group;
    const int64_t glen = min(a.group, a.cols - c0);
    const T* xp = x + r * a.cols + c0;
    float s, z;
    if (a.find) {
      float mx = -INFINITY, mn = INFINITY, amax = 0.0f;
      for (int64_t i = lane; i < glen; i += 32) {
        float v = to_f<T>(xp[i]);
        mx = nan_max(mx, v); mn = nan_min(mn, v); amax = nan_max(amax, fabsf(v));
      }
      if (glen < a.group) { mx = nan_max(mx, 0.0f); mn = nan_min(mn, 0.0f); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        amax = nan_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      }
      if constexpr (NVFP_PASS1) {
        float vb, zb;
        nvfp_block_stat<DT>(mx, mn, amax, a.c.zero_point, vb, zb);
        if (lane == 0) atomicMax(amax_key, __float_as_uint(fabsf(vb)));
        continue;
      }
      find_params<DT, DT>(a.c, mx, mn, amax, nv_g, s, z);
      if (lane == 0) {
        flag_nan_scale(a, s);
        if (sc != nullptr) sc[gid] = from_f<T>(s);
        if (zr != nullptr) zr[gid] = from_f<T>(z);
      }
    } else {
      s = to_f<T>(sc[gid]);
      z = to_f<T>(zr[gid]);
    }
    if (a.apply) {
      for (int64_t i = lane; i < glen; i += 32) {
        float code;
        float o = fake_quant<DT>(a.c, to_f<T>(xp[i]), s, z, code);
        const int64_t off = r * a.cols + c0 + i;
        out[off] = from_f<T>(o);
        if (a.codes) a.codes[off] = encode_code(a.c, code);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// mse clip search (ref: int_quant.py:115-162, fp_quant.py:127-174, mx_quant.py:114-149): 80 shrink factors
// p = 1 - i/100 of the group's (max, min); candidate parameters from the shrunken range (scale NOT clamped),
// error sum_i |QDQ(x_i) - x_i|^2.4 with every op rounded in the tensor dtype, the sum accumulated in fp32 and
// rounded once; the first candidate with a strictly smaller error wins.  One warp per group; the group is
// re-read from L1 for every candidate (the reference makes 80 passes over HBM with ~25 kernels each).
template <typename T>
__global__ void __launch_bounds__(256) qdq_mse_kernel(QdqArgs a) {
  constexpr int DT = DtOf<T>::value;
  const T* x = static_cast<const T*>(a.x);
  T* sc = static_cast<T*>(a.scales);
  T* zr = static_cast<T*>(a.zeros);
  const int lane = threadIdx.x & 31;
  const int64_t total = a.nrows * a.G;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float nv_g = (a.c.qtype == LCB_Q_NVFP) ? *a.nv_amax : 0.0f;
  for (int64_t gid = warp; gid < total; gid += nwarps) {
    const T* xp = x + gid * a.group;  // groups tile the rows
    float mx = -INFINITY, mn = INFINITY, amax = 0.0f;
    for (int64_t i = lane; i < a.group; i += 32) {
      const float v = to_f<T>(xp[i]);
      mx = nan_max(mx, v); mn = nan_min(mn, v); amax = nan_max(amax, fabsf(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      amax = nan_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    if (!a.c.zero_point) { mx = amax; mn = -amax; }  // max_val = |x|.amax(), min_val = -max_val
    float s_sel, z_sel, best = INFINITY;
    find_params_unclamped<DT>(a.c, mx, mn, amax, nv_g, s_sel, z_sel);
#pragma unroll 1
    for (int it = 0; it < 80; ++it) {
      const float p = (float)(1.0 - (double)it / 100.0);
      const float mx1 = R<DT>(__fmul_rn(mx, p)), mn1 = R<DT>(__fmul_rn(mn, p));
      float s1, z1;
      // NVFP: the tensor-wide amax of the shrunken block maxima is p * g (monotone rounding), nvfp_quant.py:121-131
      find_params_unclamped<DT>(a.c, mx1, mn1, mx1, R<DT>(__fmul_rn(nv_g, p)), s1, z1);
      float acc = 0.0f;
      for (int64_t i = lane; i < a.group; i += 32) {
        const float v = to_f<T>(xp[i]);
        float code;
        const float dq = fake_quant<DT>(a.c, v, s1, z1, code);
        const float d = fabsf(R<DT>(__fsub_rn(dq, v)));
        // pow_(2.4): torch casts the python exponent to the tensor dtype first (bf16(2.4) = 2.40625)
        acc = __fadd_rn(acc, R<DT>(powf(d, R<DT>(2.4f))));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
      const float err = R<DT>(acc);
      if (err < best) { best = err; s_sel = s1; z_sel = z1; }
    }
    if (lane == 0) {
      const float s = clamp_min_nan(s_sel, scale_floor<DT>());
      flag_nan_scale(a, s);
      sc[gid] = from_f<T>(s);
      zr[gid] = from_f<T>(z_sel);
    }
  }
}

// ------------------------------------------------------------------------------------------
// axis -2 (groups run down the rows, one parameter per column): statistics kernel + apply kernel.
// stats: CTA = (32 x 8) threads over a [group rows x 64 columns] tile; thread (tx, ty) owns
// columns 2*tx, 2*tx+1 and rows ty, ty+8, ...  Coalesced 128 B row segments for bf16.
// For NVFP the per-block statistic v is written to `scales` and finalised once the tensor-wide
// amax is known (colgroup_nvfp_finalize_kernel).
template <typename T>
__global__ void __launch_bounds__(256) colgroup_stats_kernel(QdqArgs a, T* sc, T* zr, uint32_t* amax_key) {
  constexpr int DT = DtOf<T>::value;
  __shared__ float red[3][8][64];
  const T* x = static_cast<const T*>(a.x);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t bidx = blockIdx.z, gy = blockIdx.y;
  const int64_t c = (int64_t)blockIdx.x * 64 + tx * 2;
  const int64_t r0 = gy * a.group, r1 = min(r0 + a.group, a.rows);
  float mx[2] = {-INFINITY, -INFINITY}, mn[2] = {INFINITY, INFINITY}, am[2] = {0.0f, 0.0f};
  const T* xb = x + bidx * a.rows * a.cols;
  for (int64_t r = r0 + ty; r < r1; r += 8) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (c + j < a.cols) {
        float v = to_f<T>(xb[r * a.cols + c + j]);
        mx[j] = nan_max(mx[j], v); mn[j] = nan_min(mn[j], v); am[j] = nan_max(am[j], fabsf(v));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    red[0][ty][tx * 2 + j] = mx[j]; red[1][ty][tx * 2 + j] = mn[j]; red[2][ty][tx * 2 + j] = am[j];
  }
  __syncthreads();
  const int t = ty * 32 + tx;
  if (t < 64) {
    const int64_t cc = (int64_t)blockIdx.x * 64 + t;
    if (cc < a.cols) {
      float fmx = -INFINITY, fmn = INFINITY, fam = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        fmx = nan_max(fmx, red[0][k][t]); fmn = nan_min(fmn, red[1][k][t]); fam = nan_max(fam, red[2][k][t]);
      }
      if (r1 - r0 < a.group) { fmx = nan_max(fmx, 0.0f); fmn = nan_min(fmn, 0.0f); }  // zero padded tail
      const int64_t pidx = (bidx * a.G + gy) * a.cols + cc;
      float s, z;
      if (a.c.qtype == LCB_Q_NVFP) {
        nvfp_block_stat<DT>(fmx, fmn, fam, a.c.zero_point, s, z);  // s holds v for now
        atomicMax(amax_key, __float_as_uint(fabsf(s)));
      } else {
        find_params<DT, DT>(a.c, fmx, fmn, fam, 0.0f, s, z);
        flag_nan_scale(a, s);
      }
      sc[pidx] = from_f<T>(s);
      zr[pidx] = from_f<T>(z);
    }
  }
}

template <typename T>
__global__ void colgroup_nvfp_finalize_kernel(QdqArgs a, T* sc, int64_t n, const float* amax) {
  constexpr int DT = DtOf<T>::value;
  const float g = *amax;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = nvfp_scale<DT>(to_f<T>(sc[i]), g, a.c.f);
    flag_nan_scale(a, s);
    sc[i] = from_f<T>(s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colgroup_apply_kernel(QdqArgs a, const T* sc, const T* zr) {
  constexpr int DT = DtOf<T>::value;
  const T* x = static_cast<const T*>(a.x);
  T* out = static_cast<T*>(a.out);
  const int64_t total = a.nrows * a.rows * a.cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i % a.cols;
    const int64_t br = i / a.cols;
    const int64_t r = br % a.rows, b = br / a.rows;
    const int64_t pidx = (b * a.G + r / a.group) * a.cols + c;
    float code;
    float o = fake_quant<DT>(a.c, to_f<T>(x[i]), to_f<T>(sc[pidx]), to_f<T>(zr[pidx]), code);
    out[i] = from_f<T>(o);
    if (a.codes) a.codes[i] = encode_code(a.c, code);
  }
}

// ------------------------------------------------------------------------------------------
// per tensor (group_size 0; INT / FP): global min / max / amax, then one elementwise pass.
// keys[0] = key(max), keys[1] = key(-min), keys[2] = bits(amax)
template <typename T>
__global__ void __launch_bounds__(256) tensor_stats_kernel(const T* x, int64_t n, uint32_t* keys) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float red[3][8];
  float mx = -INFINITY, mn = INFINITY, amax = 0.0f;
  const int64_t nvec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / VEC : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[VEC];
    load16<T>(x + i * VEC, v);
    stats_of<T, VEC>(v, mx, mn, amax);
  }
  for (int64_t i = nvec * VEC + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = to_f<T>(x[i]);
    mx = nan_max(mx, v); mn = nan_min(mn, v); amax = nan_max(amax, fabsf(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    amax = nan_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = mx; red[1][wid] = mn; red[2][wid] = amax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      mx = nan_max(mx, red[0][k]); mn = nan_min(mn, red[1][k]); amax = nan_max(amax, red[2][k]);
    }
    // NaN keys sort above every number in f2key order for positive-sign NaN; good enough to
    // make the result NaN (torch.amax propagates NaN).
    atomicMax(&keys[0], f2key(mx));
    atomicMax(&keys[1], f2key(-mn));
    atomicMax(&keys[2], f2key(amax));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) tensor_apply_kernel(QdqArgs a, const uint32_t* keys, int64_t n) {
  constexpr int DT = DtOf<T>::value;
  constexpr int VEC = 16 / sizeof(T);
  const T* x = static_cast<const T*>(a.x);
  T* out = static_cast<T*>(a.out);
  float s, z;
  if (a.find) {
    const float mx = key2f(keys[0]), mn = -key2f(keys[1]), amax = key2f(keys[2]);
    if (a.c.qtype == LCB_Q_INT) int_params<DT, LCB_F32>(mx, mn, amax, a.c.zero_point, a.c.f, s, z);
    else fp_params<DT, LCB_F32>(mx, mn, amax, a.c.zero_point, a.c.f, s, z);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      flag_nan_scale(a, s);
      if (a.scales) *static_cast<float*>(a.scales) = s;
      if (a.zeros) *static_cast<float*>(a.zeros) = z;
    }
  } else {
    s = *static_cast<const float*>(a.scales);
    z = *static_cast<const float*>(a.zeros);
  }
  if (!a.apply) return;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const int64_t nvec = aligned ? n / VEC : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[VEC];
    load16<T>(x + i * VEC, v);
    apply_vec<T, VEC>(a, v, s, z, out + i * VEC, a.codes ? a.codes + i * VEC : nullptr);
  }
  for (int64_t i = nvec * VEC + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float code;
    float o = fake_quant<DT>(a.c, to_f<T>(x[i]), s, z, code);
    out[i] = from_f<T>(o);
    if (a.codes) a.codes[i] = encode_code(a.c, code);
  }
}

// ------------------------------------------------------------------------------------------
// Bandwidth-shaped forms of the two-pass granularities (north_star: per-tensor and per-channel at the HBM roofline too).
// Same op-by-op arithmetic as the generic kernels (qmath.cuh), different memory shape: FLAT launches -- a CTA takes a
// fixed slab and retires, so CTAs in different phases share an SM and HBM always has requests in flight (the
// grid-stride forms above measure 1.8 TB/s for per-tensor and 0.2 TB/s for axis -2 on a 403 MB tensor) -- and 128-bit
// accesses on every path.
//
// per tensor, pass 1: min / max / amax of a slab of 4 x 256 vectors per CTA -> keys (as tensor_stats_kernel)
template <typename T>
__global__ void __launch_bounds__(256) tensor_stats_flat_kernel(const T* __restrict__ x, int64_t nvec, uint32_t* keys) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float red[3][8];
  float mx = -INFINITY, mn = INFINITY, amax = 0.0f;
  const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  float v[4][VEC];
  bool act[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = base + u * 256;
    act[u] = i < nvec;
    if (act[u]) load16<T>(x + i * VEC, v[u]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (act[u]) stats_of<T, VEC>(v[u], mx, mn, amax);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    amax = nan_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = mx; red[1][wid] = mn; red[2][wid] = amax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      mx = nan_max(mx, red[0][k]); mn = nan_min(mn, red[1][k]); amax = nan_max(amax, red[2][k]);
    }
    atomicMax(&keys[0], f2key(mx));
    atomicMax(&keys[1], f2key(-mn));
    atomicMax(&keys[2], f2key(amax));
  }
}

// per tensor, pass 2: 2 x 256 vectors per CTA
template <typename T>
__global__ void __launch_bounds__(256) tensor_apply_flat_kernel(QdqArgs a, const uint32_t* keys, int64_t nvec) {
  constexpr int DT = DtOf<T>::value;
  constexpr int VEC = 16 / sizeof(T);
  const T* x = static_cast<const T*>(a.x);
  T* out = static_cast<T*>(a.out);
  const int64_t base = (int64_t)blockIdx.x * 512 + threadIdx.x;
  float v[2][VEC];
  bool act[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int64_t i = base + u * 256;
    act[u] = i < nvec;
    if (act[u]) load16<T>(x + i * VEC, v[u]);
  }
  float s, z;
  if (a.find) {
    const float mx = key2f(keys[0]), mn = -key2f(keys[1]), amax = key2f(keys[2]);
    if (a.c.qtype == LCB_Q_INT) int_params<DT, LCB_F32>(mx, mn, amax, a.c.zero_point, a.c.f, s, z);
    else fp_params<DT, LCB_F32>(mx, mn, amax, a.c.zero_point, a.c.f, s, z);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      flag_nan_scale(a, s);
      if (a.scales) *static_cast<float*>(a.scales) = s;
      if (a.zeros) *static_cast<float*>(a.zeros) = z;
    }
  } else {
    s = *static_cast<const float*>(a.scales);
    z = *static_cast<const float*>(a.zeros);
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int64_t i = base + u * 256;
    if (act[u]) apply_vec<T, VEC>(a, v[u], s, z, out + i * VEC, a.codes ? a.codes + i * VEC : nullptr);
  }
}

// fp32 tensors, INT formats, groups along the last axis that tile the rows (the find_params / fake-quant of the permuted
// fp32 weight inside GPTQ, ref: gptq/core.py:179,198,257): the generic sub-warp kernel spends ~65 instructions per
// element (three NaN-propagating statistics, redundant parameters) and is issue bound at 2 TB/s.  Here a lane owns
// 8 floats (32 bytes), LPG lanes a group; the statistics are plain fmax / fmin plus ONE integer maximum of the
// magnitude bits that also detects non-finite inputs (those groups fall back to the NaN-propagating arithmetic);
// apply is the IEEE op sequence of int_fq<fp32>.
template <int LPG>
__global__ void __launch_bounds__(256) qdq_stream_f32_kernel(QdqArgs a) {
  const float* x = static_cast<const float*>(a.x);
  float* out = static_cast<float*>(a.out);
  float* sc = static_cast<float*>(a.scales);
  float* zr = static_cast<float*>(a.zeros);
  const bool zp = a.c.zero_point != 0;
  const int64_t chunks = a.nrows * a.cols / 8;
  const int64_t base = (int64_t)blockIdx.x * 512 + threadIdx.x;
  float v[2][8];
  int64_t c[2];
  bool act[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    c[u] = base + u * 256;
    act[u] = c[u] < chunks;
    if (act[u]) {
      const W8 w = ldg256(x + c[u] * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[u][i] = __uint_as_float(w.w[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[u][i] = 0.0f;
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    float mx = v[u][0], mn = v[u][0];
    uint32_t ab = __float_as_uint(v[u][0]) & 0x7fffffffu;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      if (zp) { mx = fmaxf(mx, v[u][i]); mn = fminf(mn, v[u][i]); }
      ab = max(ab, __float_as_uint(v[u][i]) & 0x7fffffffu);
    }
#pragma unroll
    for (int o = 1; o < LPG; o <<= 1) {
      if (zp) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      ab = max(ab, __shfl_xor_sync(0xffffffffu, ab, o));
    }
    float s, z;
    const bool finite = ab < 0x7f800000u;
    if (finite) {
      int_params<LCB_F32, LCB_F32>(mx, mn, __uint_as_float(ab), a.c.zero_point, a.c.f, s, z);
    } else {  // NaN / inf somewhere in the group: torch.amax / amin propagate NaN
      float gmx = -INFINITY, gmn = INFINITY, gam = 0.0f;
      stats_of<float, 8>(v[u], gmx, gmn, gam);
#pragma unroll
      for (int o = 1; o < LPG; o <<= 1) {
        gmx = nan_max(gmx, __shfl_xor_sync(0xffffffffu, gmx, o));
        gmn = nan_min(gmn, __shfl_xor_sync(0xffffffffu, gmn, o));
        gam = nan_max(gam, __shfl_xor_sync(0xffffffffu, gam, o));
      }
      int_params<LCB_F32, LCB_F32>(gmx, gmn, gam, a.c.zero_point, a.c.f, s, z);
    }
    if (!act[u]) continue;
    if ((c[u] & (LPG - 1)) == 0) {
      flag_nan_scale(a, s);
      const int64_t gid = c[u] / LPG;
      if (sc != nullptr) sc[gid] = s;
      if (zr != nullptr) zr[gid] = z;
    }
    if (!a.apply) continue;
    W8 o;
    const float qmax = a.c.f.qmax;
    // all-equal asymmetric groups give s0 = 0 and a NaN / inf zero point (ref: int_quant.py:93-97): reference arithmetic
    if (finite && s == s && (__float_as_uint(z) & 0x7f800000u) != 0x7f800000u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float q = __fadd_rn(__fdiv_rn(v[u][i], s), z);
        q = fminf(fmaxf(rintf(q), -qmax), qmax);
        o.w[i] = __float_as_uint(__fmul_rn(__fsub_rn(q, z), s));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float code;
        o.w[i] = __float_as_uint(fake_quant<LCB_F32>(a.c, v[u][i], s, z, code));
      }
    }
    stg256(out + c[u] * 8, o);
  }
}

// per tensor, pass 1, bf16: packed NaN-propagating statistics (stats16), only the ones the parameters need --
// (max, -min) for asymmetric, |x| maximum for symmetric; 4 x 256 chunks of 16 elements per CTA
__global__ void __launch_bounds__(256) tensor_stats_stream_kernel(const __nv_bfloat16* __restrict__ x, int64_t chunks, int zp,
                                                                  uint32_t* keys) {
  __shared__ uint32_t red[8];
  const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  W8 v[4];
  bool act[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t c = base + u * 256;
    act[u] = c < chunks;
    v[u] = ldg256(x + (act[u] ? c : 0) * 16);     // idle slots re-read chunk 0: duplicates do not change max / min
  }
  Stat2 st = stats16(v[0], zp != 0);
#pragma unroll
  for (int u = 1; u < 4; ++u) {
    const Stat2 t = stats16(v[u], zp != 0);
    stat_combine(st, t, zp != 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) stat_shfl_xor(st, o, zp != 0);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = zp ? st.mxmn : st.amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    Stat2 t = st;
    for (int k = 1; k < 8; ++k) {
      Stat2 o;
      o.mxmn = red[k]; o.amax = red[k];
      stat_combine(t, o, zp != 0);
    }
    float mx, mn, amax;
    stat_finish(t, zp != 0, mx, mn, amax);
    if (zp) {
      atomicMax(&keys[0], f2key(mx));
      atomicMax(&keys[1], f2key(-mn));
    } else {
      atomicMax(&keys[2], f2key(amax));
    }
  }
}

// per tensor, pass 2, bf16 + INT formats: packed arithmetic (apply16_tensor), 2 x 256 chunks of 16 elements per CTA.
// Non-finite parameters (NaN / inf in the tensor) and zero points that are not bf16-exact take the op-by-op arithmetic.
template <int KIND>
__global__ void __launch_bounds__(256) tensor_apply_stream_kernel(QdqArgs a, const uint32_t* keys, int64_t chunks) {
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
  const int64_t base = (int64_t)blockIdx.x * 512 + threadIdx.x;
  W8 v[2];
  bool act[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int64_t c = base + u * 256;
    act[u] = c < chunks;
    if (act[u]) v[u] = ldg256(x + c * 16);
  }
  float s, z;
  if (a.find) {
    const float mx = key2f(keys[0]), mn = -key2f(keys[1]), amax = key2f(keys[2]);
    int_params<LCB_BF16, LCB_F32>(mx, mn, amax, a.c.zero_point, a.c.f, s, z);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      flag_nan_scale(a, s);
      if (a.scales) *static_cast<float*>(a.scales) = s;
      if (a.zeros) *static_cast<float*>(a.zeros) = z;
    }
  } else {
    s = *static_cast<const float*>(a.scales);
    z = *static_cast<const float*>(a.zeros);
  }
  const bool fast = (__float_as_uint(s) & 0x7f800000u) != 0x7f800000u && (__float_as_uint(z) & 0x7f800000u) != 0x7f800000u &&
                    z == R<LCB_BF16>(z) && fabsf(z) <= 256.0f;
  const float rinv = __frcp_rn(s);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (!act[u]) continue;
    if (fast) apply16_tensor<KIND>(v[u], s, rinv, z);
    else v[u] = apply16_generic(a.c, v[u], s, z);
    stg256(out + (base + u * 256) * 16, v[u]);
  }
}

// axis -2, pass 1: a CTA reduces a [CG_ROWS rows x 32 * VEC columns] slab (a warp reads one 512-byte row segment per
// step, 8 warps = 8 rows per step) and merges its column statistics into keys[(b, g, c)][3] with atomicMax; the slab
// never crosses a group, so per-channel scaling of a tall tensor is spread over rows / CG_ROWS CTAs per column strip
// instead of one (CG_ROWS: 128 for short groups, 1024 for tall ones -- fewer atomics per column).
template <typename T>
__global__ void __launch_bounds__(256) colgroup_stats_flat_kernel(QdqArgs a, uint32_t* keys, int64_t slabs_per_group, int CG_ROWS) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float red[3][8][32 * VEC];
  const T* x = static_cast<const T*>(a.x);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t bidx = blockIdx.z;
  const int64_t g = blockIdx.y / slabs_per_group, slab = blockIdx.y % slabs_per_group;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + lane) * VEC;
  const int64_t r0 = g * a.group + slab * CG_ROWS;
  const int64_t r1 = min(min(r0 + CG_ROWS, (g + 1) * a.group), a.rows);
  float mx[VEC], mn[VEC], am[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { mx[j] = -INFINITY; mn[j] = INFINITY; am[j] = 0.0f; }
  const T* xb = x + bidx * a.rows * a.cols;
  if (c0 < a.cols) {
#pragma unroll 4
    for (int64_t r = r0 + wid; r < r1; r += 8) {
      float v[VEC];
      load16<T>(xb + r * a.cols + c0, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        mx[j] = nan_max(mx[j], v[j]); mn[j] = nan_min(mn[j], v[j]); am[j] = nan_max(am[j], fabsf(v[j]));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    red[0][wid][lane * VEC + j] = mx[j]; red[1][wid][lane * VEC + j] = mn[j]; red[2][wid][lane * VEC + j] = am[j];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 32 * VEC; t += 256) {
    const int64_t cc = (int64_t)blockIdx.x * 32 * VEC + t;
    if (cc >= a.cols) continue;
    float fmx = -INFINITY, fmn = INFINITY, fam = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      fmx = nan_max(fmx, red[0][k][t]); fmn = nan_min(fmn, red[1][k][t]); fam = nan_max(fam, red[2][k][t]);
    }
    // a group cut short by the end of the tensor is zero padded (ref: utils.py:119-132); only its last slab knows
    if (r1 == a.rows && a.rows < (g + 1) * a.group) { fmx = nan_max(fmx, 0.0f); fmn = nan_min(fmn, 0.0f); }
    uint32_t* kp = keys + ((bidx * a.G + g) * a.cols + cc) * 3;
    atomicMax(kp + 0, f2key(fmx));
    atomicMax(kp + 1, f2key(-fmn));
    atomicMax(kp + 2, f2key(fam));
  }
}

// axis -2, pass 1, bf16: the same slab decomposition with packed bf16x2 statistics -- a lane's 8 columns are four
// packed registers and the reduction DOWN the rows is an elementwise __hmax2_nan / __hmin2_nan per register (one
// instruction per two elements), only the statistics the parameters need.
__global__ void __launch_bounds__(256) colgroup_stats_stream_kernel(QdqArgs a, uint32_t* keys, int64_t slabs_per_group,
                                                                    int CG_ROWS) {
  __shared__ uint32_t red[2][8][32 * 4];
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  const bool zp = a.c.zero_point != 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t bidx = blockIdx.z;
  const int64_t g = blockIdx.y / slabs_per_group, slab = blockIdx.y % slabs_per_group;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + lane) * 8;
  const int64_t r0 = g * a.group + slab * CG_ROWS;
  const int64_t r1 = min(min(r0 + (int64_t)CG_ROWS, (g + 1) * a.group), a.rows);
  const uint32_t NINF2 = 0xff80ff80u, PINF2 = 0x7f807f80u;
  uint32_t mx[4], mn[4];   // symmetric: mx holds the |x| maxima
#pragma unroll
  for (int j = 0; j < 4; ++j) { mx[j] = zp ? NINF2 : 0u; mn[j] = PINF2; }
  const __nv_bfloat16* xb = x + bidx * a.rows * a.cols;
  if (c0 < a.cols) {
#pragma unroll 4
    for (int64_t r = r0 + wid; r < r1; r += 8) {
      const uint4 t = *reinterpret_cast<const uint4*>(xb + r * a.cols + c0);
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (zp) {
          mx[j] = bf22u(__hmax2_nan(u2bf2(mx[j]), u2bf2(w[j])));
          mn[j] = bf22u(__hmin2_nan(u2bf2(mn[j]), u2bf2(w[j])));
        } else {
          mx[j] = bf22u(__hmax2_nan(u2bf2(mx[j]), u2bf2(w[j] & 0x7fff7fffu)));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[0][wid][lane * 4 + j] = mx[j]; red[1][wid][lane * 4 + j] = mn[j]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int t = threadIdx.x;                      // packed register t covers columns 2t, 2t+1 of the strip
    uint32_t fx = red[0][0][t], fn = red[1][0][t];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      fx = bf22u(__hmax2_nan(u2bf2(fx), u2bf2(red[0][k][t])));
      if (zp) fn = bf22u(__hmin2_nan(u2bf2(fn), u2bf2(red[1][k][t])));
    }
    const bool pad = r1 == a.rows && a.rows < (g + 1) * a.group;   // zero padded tail group (ref: utils.py:119-132)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t cc = (int64_t)blockIdx.x * 256 + 2 * t + h;
      if (cc >= a.cols) continue;
      float vx = __uint_as_float(h ? (fx & 0xffff0000u) : (fx << 16));
      float vn = __uint_as_float(h ? (fn & 0xffff0000u) : (fn << 16));
      uint32_t* kp = keys + ((bidx * a.G + g) * a.cols + cc) * 3;
      if (zp) {
        if (pad) { vx = nan_max(vx, 0.0f); vn = nan_min(vn, 0.0f); }
        atomicMax(kp + 0, f2key(vx));
        atomicMax(kp + 1, f2key(-vn));
      } else {
        atomicMax(kp + 2, f2key(vx));
      }
    }
  }
}

// axis -2: keys -> parameters (NVFP: the block statistic, finalised by colgroup_nvfp_finalize_kernel)
template <typename T>
__global__ void __launch_bounds__(256) colgroup_params_kernel(QdqArgs a, const uint32_t* keys, T* sc, T* zr, int64_t n,
                                                              uint32_t* amax_key) {
  constexpr int DT = DtOf<T>::value;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float fmx = key2f(keys[3 * i]), fmn = -key2f(keys[3 * i + 1]), fam = key2f(keys[3 * i + 2]);
  // the packed bf16 statistics pass leaves the ones the configuration does not use unset (key 0)
  if (keys[3 * i + 2] == 0u) fam = nan_max(fabsf(fmx), fabsf(fmn));
  if (keys[3 * i] == 0u) { fmx = fam; fmn = -fam; }
  float s, z;
  if (a.c.qtype == LCB_Q_NVFP) {
    nvfp_block_stat<DT>(fmx, fmn, fam, a.c.zero_point, s, z);
    atomicMax(amax_key, __float_as_uint(fabsf(s)));
  } else {
    find_params<DT, DT>(a.c, fmx, fmn, fam, 0.0f, s, z);
    flag_nan_scale(a, s);
  }
  sc[i] = from_f<T>(s);
  zr[i] = from_f<T>(z);
}

// axis -2, pass 2: a thread quantises one 16-byte vector with the 16-byte vectors of its columns' parameters
template <typename T>
__global__ void __launch_bounds__(256) colgroup_apply_flat_kernel(QdqArgs a, const T* __restrict__ sc, const T* __restrict__ zr) {
  constexpr int DT = DtOf<T>::value;
  constexpr int VEC = 16 / sizeof(T);
  const T* x = static_cast<const T*>(a.x);
  T* out = static_cast<T*>(a.out);
  const int64_t vpr = a.cols / VEC;                       // vectors per row
  const int64_t nvec = a.nrows * a.rows * vpr;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int64_t i = (int64_t)blockIdx.x * 512 + u * 256 + threadIdx.x;
    if (i >= nvec) continue;
    const int64_t br = i / vpr, cv = i - br * vpr;
    const int64_t r = br % a.rows, b = br / a.rows;
    const int64_t pidx = (b * a.G + r / a.group) * a.cols + cv * VEC;
    float v[VEC], sv[VEC], zv[VEC], o[VEC];
    load16<T>(x + i * VEC, v);
    load16<T>(sc + pidx, sv);
    load16<T>(zr + pidx, zv);
    uint8_t cd[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float code;
      o[j] = fake_quant<DT>(a.c, v[j], sv[j], zv[j], code);
      cd[j] = encode_code(a.c, code);
    }
    store16<T>(out + i * VEC, o);
    if (a.codes) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) a.codes[i * VEC + j] = cd[j];
    }
  }
}

// axis -2, pass 2, bf16 + INT formats: a lane owns 16 columns (their 16 scales / zero points stay in registers as packed
// bf16x2 pairs, loaded once) and walks down CA_ROWS / 8 rows of a [CA_ROWS x 512] tile; the per-column parameters are
// bf16 (x.dtype), so the packed arithmetic of apply16 applies with (s_c, s_c+1) pairs.  Tiles never straddle a group.
constexpr int CA_ROWS = 64;
template <int KIND>
__global__ void __launch_bounds__(256) colgroup_apply_stream_kernel(QdqArgs a, const __nv_bfloat16* __restrict__ sc,
                                                                    const __nv_bfloat16* __restrict__ zr) {
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + lane) * 16;
  if (c0 >= a.cols) return;
  const int64_t tiles_per_batch = a.rows / CA_ROWS;
  const int64_t b = blockIdx.y / tiles_per_batch, r0 = (blockIdx.y % tiles_per_batch) * CA_ROWS;
  const int64_t pidx = (b * a.G + r0 / a.group) * a.cols + c0;
  const W8 sv = ldg256(sc + pidx), zv = ldg256(zr + pidx);
  float rc[16];
  bool finite = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t sw = sv.w[j], zw = zv.w[j];
    finite = finite && ((sw & 0x7f80u) != 0x7f80u) && ((sw & 0x7f800000u) != 0x7f800000u) && ((zw & 0x7f80u) != 0x7f80u) &&
             ((zw & 0x7f800000u) != 0x7f800000u);
    rc[2 * j] = rcp_fast(__uint_as_float(sw << 16));
    rc[2 * j + 1] = rcp_fast(__uint_as_float(sw & 0xffff0000u));
  }
  const __nv_bfloat16* xb = x + (b * a.rows + r0) * a.cols + c0;
  __nv_bfloat16* ob = out + (b * a.rows + r0) * a.cols + c0;
#pragma unroll 2
  for (int r = wid; r < CA_ROWS; r += 8) {
    W8 v = ldg256(xb + (int64_t)r * a.cols);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t w = v.w[j], sw = sv.w[j], zw = zv.w[j];
      const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
      if (finite) {
        __nv_bfloat162 q = u2bf2(pack_bf2(mul_snap(x0, rc[2 * j]), mul_snap(x1, rc[2 * j + 1])));
        const __nv_bfloat162 z2 = u2bf2(zw), s2 = u2bf2(sw);
        q = __hadd2_rn(q, z2);
        if constexpr (KIND == FK_INT4) {
          q = __hmin2(__hmax2(q, u2bf2(0xc0e0c0e0u)), u2bf2(0x40e040e0u));
          const __nv_bfloat162 magic = u2bf2(0x43404340u);
          q = __hsub2_rn(__hadd2_rn(q, magic), magic);
        } else {
          q = __hmin2(__hmax2(q, u2bf2(0xc2fec2feu)), u2bf2(0x42fe42feu));
          const uint32_t qb = bf22u(q);
          const float MG = 12582912.0f;
          q = u2bf2(pack_bf2(__fsub_rn(__fadd_rn(__uint_as_float(qb << 16), MG), MG),
                             __fsub_rn(__fadd_rn(__uint_as_float(qb & 0xffff0000u), MG), MG)));
        }
        v.w[j] = bf22u(__hmul2_rn(__hsub2_rn(q, z2), s2));
      } else {   // a non-finite parameter among my 16 columns: op-by-op reference arithmetic for all of them
        float code;
        const float o0 = fake_quant<LCB_BF16>(a.c, x0, __uint_as_float(sw << 16), __uint_as_float(zw << 16), code);
        const float o1 = fake_quant<LCB_BF16>(a.c, x1, __uint_as_float(sw & 0xffff0000u), __uint_as_float(zw & 0xffff0000u), code);
        v.w[j] = (__float_as_uint(o0) >> 16) | (__float_as_uint(o1) & 0xffff0000u);
      }
    }
    stg256(ob + (int64_t)r * a.cols, v);
  }
}

// ------------------------------------------------------------------------------------------
// host dispatch
static int grid_for(int64_t work_items, int64_t items_per_cta, int ctas_per_sm) {
  int64_t need = ceil_div(work_items, items_per_cta);
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}


// which fast path (qdq_fast.cuh) covers this configuration; -1 = generic kernels
static int fast_kind(const QdqArgs& a) {
  if (a.codes != nullptr) return -1;
  if (a.c.qtype == LCB_Q_INT) return a.c.f.mbits == 4 ? FK_INT4 : FK_INT8;
  if (a.c.f.ebits == 0) return -1;  // MX with integer elements
  return a.c.f.ebits == 2 ? FK_E2M1 : (a.c.f.ebits == 4 ? FK_E4M3 : FK_E5M2);
}

template <int KIND>
static int launch_subwarp_fast_k(const QdqArgs& a, int lpg, int64_t total, cudaStream_t st) {
  const int gpw = 32 / lpg;
  const int grid = grid_for(total, (int64_t)8 * gpw * 2, 8);
  switch (lpg) {
    case 1: qdq_subwarp_fast_kernel<KIND, 1><<<grid, 256, 0, st>>>(a); break;
    case 2: qdq_subwarp_fast_kernel<KIND, 2><<<grid, 256, 0, st>>>(a); break;
    case 4: qdq_subwarp_fast_kernel<KIND, 4><<<grid, 256, 0, st>>>(a); break;
    case 8: qdq_subwarp_fast_kernel<KIND, 8><<<grid, 256, 0, st>>>(a); break;
    case 16: qdq_subwarp_fast_kernel<KIND, 16><<<grid, 256, 0, st>>>(a); break;
    default: qdq_subwarp_fast_kernel<KIND, 32><<<grid, 256, 0, st>>>(a); break;
  }
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}
static int launch_subwarp_fast(const QdqArgs& a, int kind, int lpg, int64_t total, cudaStream_t st) {
  switch (kind) {
    case FK_INT4: return launch_subwarp_fast_k<FK_INT4>(a, lpg, total, st);
    case FK_INT8: return launch_subwarp_fast_k<FK_INT8>(a, lpg, total, st);
    case FK_E2M1: return launch_subwarp_fast_k<FK_E2M1>(a, lpg, total, st);
    case FK_E4M3: return launch_subwarp_fast_k<FK_E4M3>(a, lpg, total, st);
    default: return launch_subwarp_fast_k<FK_E5M2>(a, lpg, total, st);
  }
}
template <int KIND>
static int launch_rowcta_fast_k(const QdqArgs& a, int64_t total, cudaStream_t st) {
  const int grid = grid_for(total, 1, 8);
  const int64_t per = 256 * 8;
  if (a.group <= per) qdq_rowcta_fast_kernel<KIND, 1><<<grid, 256, 0, st>>>(a);
  else if (a.group <= 2 * per) qdq_rowcta_fast_kernel<KIND, 2><<<grid, 256, 0, st>>>(a);
  else if (a.group <= 4 * per) qdq_rowcta_fast_kernel<KIND, 4><<<grid, 256, 0, st>>>(a);
  else qdq_rowcta_fast_kernel<KIND, 8><<<grid, 256, 0, st>>>(a);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}
static int launch_rowcta_fast(const QdqArgs& a, int kind, int64_t total, cudaStream_t st) {
  switch (kind) {
    case FK_INT4: return launch_rowcta_fast_k<FK_INT4>(a, total, st);
    case FK_INT8: return launch_rowcta_fast_k<FK_INT8>(a, total, st);
    case FK_E2M1: return launch_rowcta_fast_k<FK_E2M1>(a, total, st);
    case FK_E4M3: return launch_rowcta_fast_k<FK_E4M3>(a, total, st);
    default: return launch_rowcta_fast_k<FK_E5M2>(a, total, st);
  }
}

template <int KIND>
static int launch_stream_k(const QdqArgs& a, int lpg, int grid, cudaStream_t st) {
  switch (lpg) {
    case 1: qdq_stream_kernel<KIND, 1><<<grid, 256, 0, st>>>(a); break;
    case 2: qdq_stream_kernel<KIND, 2><<<grid, 256, 0, st>>>(a); break;
    case 4: qdq_stream_kernel<KIND, 4><<<grid, 256, 0, st>>>(a); break;
    case 8: qdq_stream_kernel<KIND, 8><<<grid, 256, 0, st>>>(a); break;
    case 16: qdq_stream_kernel<KIND, 16><<<grid, 256, 0, st>>>(a); break;
    default: qdq_stream_kernel<KIND, 32><<<grid, 256, 0, st>>>(a); break;
  }
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}
// streaming kernels (qdq_stream.cuh): groups of 16 * 2^j <= 512 elements tiling the rows, 32 B aligned
static bool stream_geometry(const QdqArgs& a, bool need_out) {
  const int64_t lpg = a.group / 16;
  return a.group % 16 == 0 && lpg >= 1 && lpg <= 32 && (lpg & (lpg - 1)) == 0 && a.cols == a.G * a.group &&
         (reinterpret_cast<uintptr_t>(a.x) & 31) == 0 && (!need_out || (reinterpret_cast<uintptr_t>(a.out) & 31) == 0);
}
static int launch_stream(const QdqArgs& a, int kind, cudaStream_t st) {
  const int lpg = (int)(a.group / 16);
  const int64_t chunks = a.nrows * a.cols / 16;
  LCB_REQUIRE(ceil_div(chunks, 512) < (int64_t)0x7fffffff, "lcb_qdq: tensor too large for one launch");
  const int grid = (int)ceil_div(chunks, 512);  // flat: 2 x 256 chunks per CTA (see qdq_stream_kernel)
  switch (kind) {
    case FK_INT4: return launch_stream_k<FK_INT4>(a, lpg, grid, st);
    case FK_INT8: return launch_stream_k<FK_INT8>(a, lpg, grid, st);
    case FK_E2M1: return launch_stream_k<FK_E2M1>(a, lpg, grid, st);
    case FK_E4M3: return launch_stream_k<FK_E4M3>(a, lpg, grid, st);
    default: return launch_stream_k<FK_E5M2>(a, lpg, grid, st);
  }
}

// long groups (per-token rows): one CTA per group, group % 16 == 0, 512 < group <= 16384, groups tiling the rows
static bool rowstream_geometry(const QdqArgs& a) {
  return a.group % 16 == 0 && a.group > 512 && a.group <= 16384 && a.cols == a.G * a.group &&
         ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out)) & 31) == 0 &&
         a.nrows * a.G < (int64_t)0x7fffffff;
}
template <int KIND>
static int launch_rowstream_k(const QdqArgs& a, cudaStream_t st) {
  const int cpg = (int)(a.group / 16);
  const int grid = (int)(a.nrows * a.G);
  const int cpt = cpg <= 256 ? 1 : (cpg <= 512 ? 2 : 4);
  int threads = (int)ceil_div(ceil_div(cpg, cpt), 32) * 32;
  if (cpt == 1) qdq_rowstream_kernel<KIND, 1><<<grid, threads, 0, st>>>(a);
  else if (cpt == 2) qdq_rowstream_kernel<KIND, 2><<<grid, threads, 0, st>>>(a);
  else qdq_rowstream_kernel<KIND, 4><<<grid, threads, 0, st>>>(a);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}
static int launch_rowstream(const QdqArgs& a, int kind, cudaStream_t st) {
  switch (kind) {
    case FK_INT4: return launch_rowstream_k<FK_INT4>(a, st);
    case FK_INT8: return launch_rowstream_k<FK_INT8>(a, st);
    case FK_E2M1: return launch_rowstream_k<FK_E2M1>(a, st);
    case FK_E4M3: return launch_rowstream_k<FK_E4M3>(a, st);
    default: return launch_rowstream_k<FK_E5M2>(a, st);
  }
}

template <typename T, bool P1>
static int launch_rowwise(const QdqArgs& a, uint32_t* amax_key, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const int64_t total = a.nrows * a.G;
  const bool ptr_ok = ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out) |
                        (a.codes ? reinterpret_cast<uintptr_t>(a.codes) : 0)) & 15) == 0;
  const bool vec_ok = ptr_ok && (a.cols % VEC == 0) && (a.group % VEC == 0);
  const int64_t lpg = a.group / VEC;
  if constexpr (std::is_same<T, __nv_bfloat16>::value && P1) {
    if (stream_geometry(a, false)) {
      const int grid = (int)ceil_div(a.nrows * a.cols / 16, 1024);
      switch (a.group / 16) {
        case 1: nvfp_amax_stream_kernel<1><<<grid, 256, 0, st>>>(a, amax_key); break;
        case 2: nvfp_amax_stream_kernel<2><<<grid, 256, 0, st>>>(a, amax_key); break;
        case 4: nvfp_amax_stream_kernel<4><<<grid, 256, 0, st>>>(a, amax_key); break;
        case 8: nvfp_amax_stream_kernel<8><<<grid, 256, 0, st>>>(a, amax_key); break;
        case 16: nvfp_amax_stream_kernel<16><<<grid, 256, 0, st>>>(a, amax_key); break;
        default: nvfp_amax_stream_kernel<32><<<grid, 256, 0, st>>>(a, amax_key); break;
      }
      LCB_LAUNCH_CHECK();
      return LCB_OK;
    }
    if (vec_ok && a.cols == a.G * a.group && lpg >= 1 && lpg <= 32 && (lpg & (lpg - 1)) == 0) {
      const int grid = grid_for(total, (int64_t)8 * (32 / (int)lpg) * 2, 8);
      switch (lpg) {
        case 1: nvfp_amax_fast_kernel<1><<<grid, 256, 0, st>>>(a, amax_key); break;
        case 2: nvfp_amax_fast_kernel<2><<<grid, 256, 0, st>>>(a, amax_key); break;
        case 4: nvfp_amax_fast_kernel<4><<<grid, 256, 0, st>>>(a, amax_key); break;
        case 8: nvfp_amax_fast_kernel<8><<<grid, 256, 0, st>>>(a, amax_key); break;
        case 16: nvfp_amax_fast_kernel<16><<<grid, 256, 0, st>>>(a, amax_key); break;
        default: nvfp_amax_fast_kernel<32><<<grid, 256, 0, st>>>(a, amax_key); break;
      }
      LCB_LAUNCH_CHECK();
      return LCB_OK;
    }
  }
  if constexpr (std::is_same<T, __nv_bfloat16>::value && !P1) {
    const int kind = fast_kind(a);
    if (kind >= 0 && a.find && a.apply && stream_geometry(a, true)) return launch_stream(a, kind, st);
    if (kind >= 0 && a.find && a.apply && rowstream_geometry(a)) return launch_rowstream(a, kind, st);
    if (kind >= 0 && vec_ok) {
      if (lpg >= 1 && lpg <= 32 && (lpg & (lpg - 1)) == 0) return launch_subwarp_fast(a, kind, (int)lpg, total, st);
      if (a.group <= (int64_t)256 * VEC * 8) return launch_rowcta_fast(a, kind, total, st);
    }
  }
  if constexpr (std::is_same<T, float>::value && !P1) {
    const int64_t l8 = a.group / 8;
    if (a.c.qtype == LCB_Q_INT && a.find && !a.mse && a.codes == nullptr && a.group % 8 == 0 && a.cols == a.G * a.group &&
        l8 >= 1 && l8 <= 32 && (l8 & (l8 - 1)) == 0 &&
        ((reinterpret_cast<uintptr_t>(a.x) | (a.apply ? reinterpret_cast<uintptr_t>(a.out) : 0)) & 31) == 0) {
      const unsigned grid = (unsigned)ceil_div(a.nrows * a.cols / 8, 512);
      switch (l8) {
        case 1: qdq_stream_f32_kernel<1><<<grid, 256, 0, st>>>(a); break;
        case 2: qdq_stream_f32_kernel<2><<<grid, 256, 0, st>>>(a); break;
        case 4: qdq_stream_f32_kernel<4><<<grid, 256, 0, st>>>(a); break;
        case 8: qdq_stream_f32_kernel<8><<<grid, 256, 0, st>>>(a); break;
        case 16: qdq_stream_f32_kernel<16><<<grid, 256, 0, st>>>(a); break;
        default: qdq_stream_f32_kernel<32><<<grid, 256, 0, st>>>(a); break;
      }
      LCB_LAUNCH_CHECK();
      return LCB_OK;
    }
  }
  if (vec_ok && lpg >= 1 && lpg <= 32 && (lpg & (lpg - 1)) == 0) {
    const int gpw = 32 / (int)lpg;
    // flat: every warp takes one set of groups and the CTA retires (see qdq_stream.cuh on grid-stride vs flat)
    const int grid = (int)std::min<int64_t>(ceil_div(total, (int64_t)8 * gpw * 4), (int64_t)1 << 30);
    switch (lpg) {
      case 1: qdq_subwarp_kernel<T, 1, P1><<<grid, 256, 0, st>>>(a, amax_key); break;
      case 2: qdq_subwarp_kernel<T, 2, P1><<<grid, 256, 0, st>>>(a, amax_key); break;
      case 4: qdq_subwarp_kernel<T, 4, P1><<<grid, 256, 0, st>>>(a, amax_key); break;
      case 8: qdq_subwarp_kernel<T, 8, P1><<<grid, 256, 0, st>>>(a, amax_key); break;
      case 16: qdq_subwarp_kernel<T, 16, P1><<<grid, 256, 0, st>>>(a, amax_key); break;
      default: qdq_subwarp_kernel<T, 32, P1><<<grid, 256, 0, st>>>(a, amax_key); break;
    }
  } else if (vec_ok && a.group <= (int64_t)256 * VEC * 8) {
    const int grid = grid_for(total, 1, 8);
    const int64_t per = (int64_t)256 * VEC;
    if (a.group <= per) qdq_rowcta_kernel<T, 1, P1><<<grid, 256, 0, st>>>(a, amax_key);
    else if (a.group <= 2 * per) qdq_rowcta_kernel<T, 2, P1><<<grid, 256, 0, st>>>(a, amax_key);
    else if (a.group <= 4 * per) qdq_rowcta_kernel<T, 4, P1><<<grid, 256, 0, st>>>(a, amax_key);
    else qdq_rowcta_kernel<T, 8, P1><<<grid, 256, 0, st>>>(a, amax_key);
  } else {
    const int grid = grid_for(total, 8, 8);
    qdq_generic_kernel<T, P1><<<grid, 256, 0, st>>>(a, amax_key);
  }
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

template <typename T>
static int qdq_typed(QdqArgs a, int axis, int64_t batch, void* ws, size_t ws_bytes, cudaStream_t st) {
  uint32_t* keys = static_cast<uint32_t*>(ws);  // 4 words at the start of the workspace
  const bool nvfp = a.c.qtype == LCB_Q_NVFP;
  if (a.group == 0) {  // per tensor
    const int64_t n = a.nrows * a.cols;
    constexpr int VEC = 16 / sizeof(T);
    const bool flat = n % VEC == 0 && n >= 4096 &&
                      ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out) |
                        (a.codes ? reinterpret_cast<uintptr_t>(a.codes) : 0)) & 15) == 0;
    if (a.find) {
      LCB_CUDA(cudaMemsetAsync(keys, 0, 16, st));
      bool packed = false;
      if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        if (flat && n % 16 == 0 && (reinterpret_cast<uintptr_t>(a.x) & 31) == 0) {
          tensor_stats_stream_kernel<<<(unsigned)ceil_div(n / 16, 1024), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a.x), n / 16,
                                                                                  a.c.zero_point, keys);
          packed = true;
        }
      }
      if (packed) {
      } else if (flat) tensor_stats_flat_kernel<T><<<(unsigned)ceil_div(n / VEC, 1024), 256, 0, st>>>(static_cast<const T*>(a.x), n / VEC, keys);
      else tensor_stats_kernel<T><<<grid_for(n, 256 * 16, 4), 256, 0, st>>>(static_cast<const T*>(a.x), n, keys);
      LCB_LAUNCH_CHECK();
    }
    bool done = false;
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      if (a.apply && flat && a.c.qtype == LCB_Q_INT && a.codes == nullptr && n % 16 == 0 &&
          ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out)) & 31) == 0) {
        const unsigned g = (unsigned)ceil_div(n / 16, 512);
        if (a.c.f.mbits == 4) tensor_apply_stream_kernel<FK_INT4><<<g, 256, 0, st>>>(a, keys, n / 16);
        else tensor_apply_stream_kernel<FK_INT8><<<g, 256, 0, st>>>(a, keys, n / 16);
        LCB_LAUNCH_CHECK();
        done = true;
      }
    }
    if (done) {
    } else if (a.apply && flat) {
      tensor_apply_flat_kernel<T><<<(unsigned)ceil_div(n / VEC, 512), 256, 0, st>>>(a, keys, n / VEC);
      LCB_LAUNCH_CHECK();
    } else if (a.apply || a.scales || a.zeros) {
      tensor_apply_kernel<T><<<a.apply ? grid_for(n, 256 * 16, 8) : 1, 256, 0, st>>>(a, keys, n);
      LCB_LAUNCH_CHECK();
    }
    return LCB_OK;
  }
  if (axis == -1) {
    if (nvfp && a.find && a.nv_amax == nullptr) {
      LCB_CUDA(cudaMemsetAsync(keys, 0, 16, st));
      int rc = launch_rowwise<T, true>(a, keys, st);
      if (rc != LCB_OK) return rc;
      a.nv_amax = reinterpret_cast<const float*>(keys);
    }
    if (a.find && a.mse) {
      // clip search writes the parameters, the ordinary kernels apply them
      const int64_t total = a.nrows * a.G;
      qdq_mse_kernel<T><<<grid_for(total, 8, 8), 256, 0, st>>>(a);
      LCB_LAUNCH_CHECK();
      if (!a.apply) return LCB_OK;
      a.find = 0;
    }
    return launch_rowwise<T, false>(a, nullptr, st);
  }
  // axis == -2
  const int64_t nparams = batch * a.G * a.cols;
  T* sc = static_cast<T*>(a.scales);
  T* zr = static_cast<T*>(a.zeros);
  if (a.find) {
    char* p = static_cast<char*>(ws) + 16;
    if (sc == nullptr) { sc = reinterpret_cast<T*>(p); p += (size_t)nparams * sizeof(T); }
    if (zr == nullptr) { zr = reinterpret_cast<T*>(p); }
    if (nvfp && a.nv_amax == nullptr) LCB_CUDA(cudaMemsetAsync(keys, 0, 16, st));
    constexpr int VEC = 16 / sizeof(T);
    const size_t koff = (16 + 2 * (size_t)nparams * sizeof(T) + 15) & ~(size_t)15;
    const int cg_rows = std::min<int64_t>(a.group, a.rows) <= 2048 ? 128 : 1024;
    const int64_t spg = ceil_div(std::min<int64_t>(a.group, a.rows), cg_rows);   // row slabs per group
    const bool flat_ok = a.cols % VEC == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 &&
                         ws_bytes >= koff + (size_t)nparams * 12 && a.G * spg <= 65535 && batch <= 65535;
    if (flat_ok) {
      // statistics as order-preserving keys [nparams][3] behind the parameter scratch, merged over row slabs
      uint32_t* ckeys = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + koff);
      LCB_CUDA(cudaMemsetAsync(ckeys, 0, (size_t)nparams * 12, st));
      dim3 grid((unsigned)ceil_div(a.cols, 32 * VEC), (unsigned)(a.G * spg), (unsigned)batch);
      bool packed = false;
      if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        colgroup_stats_stream_kernel<<<grid, 256, 0, st>>>(a, ckeys, spg, cg_rows);
        packed = true;
      }
      if (!packed) colgroup_stats_flat_kernel<T><<<grid, 256, 0, st>>>(a, ckeys, spg, cg_rows);
      LCB_LAUNCH_CHECK();
      colgroup_params_kernel<T><<<(unsigned)ceil_div(nparams, 256), 256, 0, st>>>(a, ckeys, sc, zr, nparams, keys);
      LCB_LAUNCH_CHECK();
    } else {
      dim3 grid((unsigned)ceil_div(a.cols, 64), (unsigned)a.G, (unsigned)batch), block(32, 8);
      colgroup_stats_kernel<T><<<grid, block, 0, st>>>(a, sc, zr, keys);
      LCB_LAUNCH_CHECK();
    }
    if (nvfp) {
      const float* g = a.nv_amax ? a.nv_amax : reinterpret_cast<const float*>(keys);
      colgroup_nvfp_finalize_kernel<T><<<grid_for(nparams, 256, 4), 256, 0, st>>>(a, sc, nparams, g);
      LCB_LAUNCH_CHECK();
    }
  }
  if (a.apply) {
    constexpr int VEC2 = 16 / sizeof(T);
    const bool vec = a.cols % VEC2 == 0 && ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out) |
                                              reinterpret_cast<uintptr_t>(sc) | reinterpret_cast<uintptr_t>(zr)) & 15) == 0;
    bool done = false;
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      if (a.c.qtype == LCB_Q_INT && a.codes == nullptr && a.cols % 16 == 0 && a.rows % CA_ROWS == 0 && a.group % CA_ROWS == 0 &&
          a.nrows * (a.rows / CA_ROWS) <= 65535 &&
          ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(sc) |
            reinterpret_cast<uintptr_t>(zr)) & 31) == 0) {
        dim3 g((unsigned)ceil_div(a.cols, 512), (unsigned)(a.nrows * (a.rows / CA_ROWS)));
        if (a.c.f.mbits == 4) colgroup_apply_stream_kernel<FK_INT4><<<g, 256, 0, st>>>(a, sc, zr);
        else colgroup_apply_stream_kernel<FK_INT8><<<g, 256, 0, st>>>(a, sc, zr);
        done = true;
      }
    }
    if (done) {
    } else if (vec) {
      const int64_t nvec = a.nrows * a.rows * (a.cols / VEC2);
      colgroup_apply_flat_kernel<T><<<(unsigned)ceil_div(nvec, 512), 256, 0, st>>>(a, sc, zr);
    } else {
      colgroup_apply_kernel<T><<<grid_for(a.nrows * a.rows * a.cols, 256 * 8, 8), 256, 0, st>>>(a, sc, zr);
    }
    LCB_LAUNCH_CHECK();
  }
  return LCB_OK;
}

static int check_cfg(const lcb_quant_cfg* cfg, int64_t group) {
  LCB_REQUIRE(cfg != nullptr, "cfg is NULL");
  LCB_REQUIRE(cfg->qtype >= LCB_Q_INT && cfg->qtype <= LCB_Q_NVFP, "unknown qtype %d", cfg->qtype);
  LCB_REQUIRE(cfg->elem >= LCB_E_INT4 && cfg->elem <= LCB_E_FP8_E5M2, "unknown element format %d", cfg->elem);
  if (cfg->qtype == LCB_Q_INT) LCB_REQUIRE(cfg->elem <= LCB_E_INT8, "INT quantizer needs int4/int8");
  if (cfg->qtype == LCB_Q_FP) LCB_REQUIRE(cfg->elem >= LCB_E_FP4_E2M1, "FP quantizer needs an fp element format");
  if (cfg->qtype == LCB_Q_NVFP) LCB_REQUIRE(cfg->elem == LCB_E_FP4_E2M1, "NVFP quantizer needs fp4_e2m1");
  if (group == 0) LCB_REQUIRE(cfg->qtype <= LCB_Q_FP, "per-tensor quantisation exists for INT / FP only");
  return LCB_OK;
}

static QCfg make_qcfg(const lcb_quant_cfg* cfg) {
  QCfg c{};
  c.qtype = cfg->qtype;
  c.zero_point = cfg->zero_point ? 1 : 0;
  const int eb = cfg->scale_ebits > 0 ? cfg->scale_ebits : 8;
  c.scale_emax = (float)((1 << (eb - 1)) - 1);
  c.f = make_fmt(cfg->elem);
  return c;
}

}  // namespace lcb

using namespace lcb;

extern "C" size_t lcb_qdq_ws_bytes(const lcb_quant_cfg* cfg, int dtype, int64_t batch, int64_t rows, int64_t cols,
                                   int axis, int64_t group) {
  (void)cfg;
  size_t bytes = 16;
  if (group > 0 && axis == -2) {
    const size_t es = dtype == LCB_BF16 ? 2 : 4;
    const size_t np = (size_t)(batch * ceil_div(rows, group) * cols);
    bytes += 2 * np * es + 32 + np * 12;   // parameter scratch + per-column statistic keys
  }
  return bytes;
}

extern "C" int lcb_qdq(const lcb_quant_cfg* cfg, int dtype, int mode, const void* x, void* out, int64_t batch,
                       int64_t rows, int64_t cols, int axis, int64_t group, void* scales, void* zeros, uint8_t* codes,
                       const float* nv_amax, void* ws, size_t ws_bytes, uint32_t* status, void* stream) {
  int rc = check_cfg(cfg, group);
  if (rc != LCB_OK) return rc;
  LCB_REQUIRE(dtype == LCB_F32 || dtype == LCB_BF16, "dtype must be LCB_F32 or LCB_BF16");
  LCB_REQUIRE(batch >= 0 && rows >= 0 && cols >= 0 && group >= 0, "negative geometry");
  LCB_REQUIRE(axis == -1 || axis == -2, "axis must be -1 or -2");
  LCB_REQUIRE((mode & (LCB_QDQ_FIND | LCB_QDQ_APPLY)) != 0, "mode selects nothing");
  if (batch * rows * cols == 0) return LCB_OK;
  LCB_REQUIRE(x != nullptr, "x is NULL");
  const int find = (mode & LCB_QDQ_FIND) ? 1 : 0, apply = (mode & LCB_QDQ_APPLY) ? 1 : 0;
  LCB_REQUIRE(!apply || out != nullptr, "APPLY needs out");
  LCB_REQUIRE(find || (scales != nullptr && zeros != nullptr), "without FIND, scales and zeros are inputs");
  const size_t need = lcb_qdq_ws_bytes(cfg, dtype, batch, rows, cols, axis, group);
  if (ws == nullptr || ws_bytes < need) {
    set_error("lcb_qdq: workspace of %zu bytes needed, %zu given", need, ws_bytes);
    return LCB_ERR_WORKSPACE;
  }
  QdqArgs a{};
  a.x = x; a.out = out; a.scales = scales; a.zeros = zeros; a.codes = apply ? codes : nullptr;
  a.nv_amax = nv_amax; a.status = status;
  a.cols = cols; a.group = group; a.find = find; a.apply = apply;
  a.c = make_qcfg(cfg);
  a.mse = (find && cfg->mse) ? 1 : 0;
  if (a.mse) {
    if ((cfg->qtype == LCB_Q_NVFP && cfg->zero_point) || axis != -1 || group <= 0 || cols % group != 0) {
      set_error("lcb_qdq: the mse clip search is implemented for INT / FP / MX / symmetric NVFP, axis -1, groups tiling the rows");
      return LCB_ERR_UNSUPPORTED;
    }
    LCB_REQUIRE(scales != nullptr && zeros != nullptr, "lcb_qdq: mse needs scales / zeros output buffers");
  }
  if (group == 0 || axis == -1) {
    a.nrows = batch * rows; a.rows = rows;
    a.G = group ? ceil_div(cols, group) : 1;
  } else {
    a.nrows = batch; a.rows = rows;
    a.G = ceil_div(rows, group);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == LCB_BF16 ? qdq_typed<__nv_bfloat16>(a, axis, batch, ws, ws_bytes, st)
                           : qdq_typed<float>(a, axis, batch, ws, ws_bytes, st);
}

extern "C" int lcb_nvfp_global_amax(const lcb_quant_cfg* cfg, int dtype, const void* x, int64_t batch, int64_t rows,
                                    int64_t cols, int axis, int64_t group, float* amax_out, void* stream) {
  int rc = check_cfg(cfg, group);
  if (rc != LCB_OK) return rc;
  LCB_REQUIRE(cfg->qtype == LCB_Q_NVFP && group > 0, "lcb_nvfp_global_amax is for NVFP with group > 0");
  LCB_REQUIRE(axis == -1, "lcb_nvfp_global_amax: only axis -1 (use lcb_qdq for axis -2)");
  LCB_REQUIRE(amax_out != nullptr && x != nullptr, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  QdqArgs a{};
  a.x = x; a.cols = cols; a.group = group; a.find = 1; a.apply = 0;
  a.c = make_qcfg(cfg);
  a.nrows = batch * rows; a.rows = rows; a.G = ceil_div(cols, group);
  LCB_CUDA(cudaMemsetAsync(amax_out, 0, sizeof(float), st));
  uint32_t* key = reinterpret_cast<uint32_t*>(amax_out);
  return dtype == LCB_BF16 ? launch_rowwise<__nv_bfloat16, true>(a, key, st) : launch_rowwise<float, true>(a, key, st);
}

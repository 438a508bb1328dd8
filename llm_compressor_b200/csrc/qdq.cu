// (c) Fused quantize-dequantize kernels for sm_100a.  HBM-bound streaming work: 128-bit
// coalesced loads/stores, sub-warp shuffle reductions for the group statistics, one pass over
// HBM wherever the reference's granularity allows it (group / token), two passes where a
// tensor-wide or column-wide statistic must exist first (per-tensor, per-channel, NVFP amax).
//
// Replaces ref: quantizers/{int,fp,mx,nvfp}_quant.py find_params/forward/fake_quantize and
// quantizers/utils.py _reshape_to_blocks/_undo_reshape_to_blocks/_quantize_elemwise_core.
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
  constexpr int GPW = 32 / LPG;  // groups per warp
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

  for (int64_t g0 = warp * GPW; g0 < total; g0 += nwarps * GPW) {
    const int64_t gid = g0 + sub;
    const bool active = gid < total;
    const int64_t r = active ? gid / a.G : 0;
    const int64_t b = active ? gid - r * a.G : 0;
    const int64_t col = b * a.group + (int64_t)sl * VEC;
    const bool inb = active && col < a.cols;
    float v[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.0f;
    if (inb) load16<T>(x + r * a.cols + col, v);
    float s, z;
    if (a.find) {
      float mx = -INFINITY, mn = INFINITY, amax = 0.0f;
      stats_of<T, VEC>(v, mx, mn, amax);  // out-of-range lanes contribute the zero padding
#pragma unroll
      for (int o = 1; o < LPG; o <<= 1) {
        mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        amax = nan_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      }
      if constexpr (NVFP_PASS1) {
        float vb, zb;
        nvfp_block_stat<DT>(mx, mn, amax, a.c.zero_point, vb, zb);
        if (active) local_amax = fmaxf(local_amax, fabsf(vb));
        continue;
      }
      find_params<DT, DT>(a.c, mx, mn, amax, nv_g, s, z);
      if (active && sl == 0) {
        flag_nan_scale(a, s);
        if (sc != nullptr) sc[gid] = from_f<T>(s);
        if (zr != nullptr) zr[gid] = from_f<T>(z);
      }
    } else {
      s = active ? to_f<T>(sc[gid]) : 1.0f;
      z = active ? to_f<T>(zr[gid]) : 0.0f;
    }
    if (a.apply && inb) {
      const int64_t off = r * a.cols + col;
      apply_vec<T, VEC>(a, v, s, z, out + off, a.codes ? a.codes + off : nullptr);
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
  if (vec_ok && lpg >= 1 && lpg <= 32 && (lpg & (lpg - 1)) == 0) {
    const int gpw = 32 / (int)lpg;
    const int grid = grid_for(total, (int64_t)8 * gpw, 8);
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
    if (a.find) {
      LCB_CUDA(cudaMemsetAsync(keys, 0, 16, st));
      tensor_stats_kernel<T><<<grid_for(n, 256 * 16, 4), 256, 0, st>>>(static_cast<const T*>(a.x), n, keys);
      LCB_LAUNCH_CHECK();
    }
    if (a.apply || a.scales || a.zeros) {
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
    dim3 grid((unsigned)ceil_div(a.cols, 64), (unsigned)a.G, (unsigned)batch), block(32, 8);
    colgroup_stats_kernel<T><<<grid, block, 0, st>>>(a, sc, zr, keys);
    LCB_LAUNCH_CHECK();
    if (nvfp) {
      const float* g = a.nv_amax ? a.nv_amax : reinterpret_cast<const float*>(keys);
      colgroup_nvfp_finalize_kernel<T><<<grid_for(nparams, 256, 4), 256, 0, st>>>(a, sc, nparams, g);
      LCB_LAUNCH_CHECK();
    }
  }
  if (a.apply) {
    colgroup_apply_kernel<T><<<grid_for(a.nrows * a.rows * a.cols, 256 * 8, 8), 256, 0, st>>>(a, sc, zr);
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
    bytes += 2 * (size_t)(batch * ceil_div(rows, group) * cols) * es + 32;
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

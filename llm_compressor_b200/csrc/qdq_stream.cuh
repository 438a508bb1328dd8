// Streaming fake-quant kernels for the common geometry: bf16 tensor, groups of 16..512 elements along
// the last axis that tile the rows exactly, find_params + fake_quantize in one pass, no code output.
//
// ncu on the previous fast path showed it issue-bound (75-78 % issue slots, 27-43 thread instructions
// per element) instead of HBM-bound.  This version cuts the instruction count ~3x:
//   * one 256-bit load / store per lane (16 bf16), lanes of a group adjacent in a warp, two chunks in
//     flight per lane; a lane owns a whole NVFP block, two lanes an MX block, eight an int4-g128 group,
//     so the group parameters are computed by few lanes and reduced with log2(LPG) shuffles;
//   * everything that torch rounds to bf16 after each op runs as packed bf16x2 arithmetic (one
//     rounding per op, two elements per instruction): abs / max / min statistics, + z, clamp, rint by
//     magic number, - z, * s, and the element rounding of e2m1 / e4m3 / e5m2 done on the bit patterns
//     (add half an ulp of the kept mantissa, mask), with the e2m1 sub-unit range as two threshold
//     compares (0x3E7F is the reference's bf16(|y| + 0.5) tie quirk, see qdq_fast.cuh);
//   * x / s as x * rcp(s) snapped onto bf16 ties (exact for 8-bit significands, qdq_fast.cuh);
//   * groups whose scale is not finite (NaN / inf inputs) and lanes in the sub-normal range of the fp8
//     formats leave the fast path and run the op-by-op reference arithmetic of qmath.cuh.
// Results are bit-identical to the generic kernels (tests/test_qdq_gpu.py runs every golden case and
// exhaustive element sweeps through both).
#pragma once

#include "qdq_apply.cuh"

namespace lcb {

// KIND: element rounding; LPG: lanes (16-element chunks) per group, power of two <= 32.
template <int KIND, int LPG>
__global__ void __launch_bounds__(256) qdq_stream_kernel(QdqArgs a) {
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
  __nv_bfloat16* sc = static_cast<__nv_bfloat16*>(a.scales);
  __nv_bfloat16* zr = static_cast<__nv_bfloat16*>(a.zeros);
  const bool zp = a.c.zero_point != 0;
  const int64_t chunks = a.nrows * a.cols / 16;
  float nv_g = 0.0f;
  if (a.c.qtype == LCB_Q_NVFP) nv_g = *a.nv_amax;
  // Not persistent on purpose: a CTA handles 2 x 256 chunks and retires, so CTAs in different phases
  // (loading / computing / storing) share an SM and the memory system always has requests in flight
  // (a flat copy kernel measures 6.7 TB/s on this part against 5.9-6.4 TB/s for grid-stride loops).
  const int64_t base = (int64_t)blockIdx.x * 512 + threadIdx.x;
  W8 v[2];
  int64_t c[2];
  bool act[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    c[u] = base + u * 256;
    act[u] = c[u] < chunks;
    if (act[u]) {
      v[u] = ldg256(x + c[u] * 16);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[u].w[i] = 0u;
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    Stat2 st = stats16(v[u], zp);
#pragma unroll
    for (int o = 1; o < LPG; o <<= 1) stat_shfl_xor(st, o, zp);
    float mx, mn, amax, s, z;
    stat_finish(st, zp, mx, mn, amax);
    if constexpr (KIND == FK_INT4 || KIND == FK_INT8) int_params_fast(mx, mn, amax, zp, a.c.f, s, z);
    else find_params<LCB_BF16, LCB_BF16>(a.c, mx, mn, amax, nv_g, s, z);
    if (!act[u]) continue;
    if ((c[u] & (LPG - 1)) == 0) {
      flag_nan_scale(a, s);
      const int64_t gid = c[u] / LPG;
      if (sc != nullptr) sc[gid] = __float2bfloat16_rn(s);
      if (zr != nullptr) zr[gid] = __float2bfloat16_rn(z);
    }
    // finite parameters <=> every element of the group is finite and the fast arithmetic is exact
    const bool finite = (__float_as_uint(s) & 0x7f800000u) != 0x7f800000u && (__float_as_uint(z) & 0x7f800000u) != 0x7f800000u;
    if (finite) apply16<KIND>(v[u], s, z, zp);
    else v[u] = apply16_generic(a.c, v[u], s, z);
    stg256(out + c[u] * 16, v[u]);
  }
}

// NVFP pass 1: whole-tensor maximum of the block statistics (ref: nvfp_quant.py:87).  Flat launch, 4 x 256
// chunks per CTA.
template <int LPG>
__global__ void __launch_bounds__(256) nvfp_amax_stream_kernel(QdqArgs a, uint32_t* amax_key) {
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  const bool zp = a.c.zero_point != 0;
  const int64_t chunks = a.nrows * a.cols / 16;
  const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  float local = 0.0f;
  uint32_t am_sym = 0u;  // symmetric: the tensor amax is the maximum of |x| -- no per-block work at all
  W8 v[4];
  bool act[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t c = base + u * 256;
    act[u] = c < chunks;
    if (act[u]) {
      v[u] = ldg256(x + c * 16);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[u].w[i] = 0u;
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (!zp) {
      const Stat2 st = stats16(v[u], false);
      am_sym = bf22u(__hmax2_nan(u2bf2(am_sym), u2bf2(st.amax)));
    } else {
      Stat2 st = stats16(v[u], true);
#pragma unroll
      for (int o = 1; o < LPG; o <<= 1) stat_shfl_xor(st, o, true);
      float mx, mn, amax, vb, zb;
      stat_finish(st, true, mx, mn, amax);
      nvfp_block_stat<LCB_BF16>(mx, mn, amax, 1, vb, zb);
      if (act[u]) local = fmaxf(local, fabsf(vb));
      if (act[u] && vb != vb) local = __uint_as_float(0x7fc00000u);
    }
  }
  if (!zp) {
    const __nv_bfloat162 p = u2bf2(am_sym);
    local = __bfloat162float(__hmax_nan(p.x, p.y));
  }
  // NaN must win the reduction like torch.amax: positive-sign NaN patterns order above every finite key
  __shared__ uint32_t red[8];
  uint32_t key = __float_as_uint(local);
  if (local != local) key = 0x7fc00000u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) key = max(key, red[i]);
    atomicMax(amax_key, key);
  }
}

// One CTA per group for long groups (per-token rows): the group stays in registers between the
// reduction and the quantisation, CPT 16-element chunks per thread, blockDim = ceil(group / (16 * CPT)).
template <int KIND, int CPT>
__global__ void __launch_bounds__(256) qdq_rowstream_kernel(QdqArgs a) {
  __shared__ uint32_t red[2][8];
  __shared__ float bc[2];
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(a.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
  __nv_bfloat16* sc = static_cast<__nv_bfloat16*>(a.scales);
  __nv_bfloat16* zr = static_cast<__nv_bfloat16*>(a.zeros);
  const bool zp = a.c.zero_point != 0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t gid = blockIdx.x;
  const int cpg = (int)(a.group / 16);  // chunks per group
  const int64_t off = gid * a.group;    // groups tile the tensor
  float nv_g = 0.0f;
  if (a.c.qtype == LCB_Q_NVFP) nv_g = *a.nv_amax;
  W8 v[CPT];
  bool act[CPT];
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    const int ch = u * blockDim.x + tid;
    act[u] = ch < cpg;
    // slots beyond the group re-read its first chunk: duplicates do not change max / min
    v[u] = ldg256(x + off + (int64_t)(act[u] ? ch : 0) * 16);
  }
  Stat2 st = stats16(v[0], zp);
#pragma unroll
  for (int u = 1; u < CPT; ++u) {
    const Stat2 t = stats16(v[u], zp);
    stat_combine(st, t, zp);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) stat_shfl_xor(st, o, zp);
  if (lane == 0) { red[0][wid] = st.mxmn; red[1][wid] = st.amax; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    Stat2 t;
    t.mxmn = lane < nw ? red[0][lane] : red[0][0];
    t.amax = lane < nw ? red[1][lane] : red[1][0];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) stat_shfl_xor(t, o, zp);
    if (lane == 0) {
      float mx, mn, amax, s, z;
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
  const float s = bc[0], z = bc[1];
  const bool finite = (__float_as_uint(s) & 0x7f800000u) != 0x7f800000u && (__float_as_uint(z) & 0x7f800000u) != 0x7f800000u;
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    if (!act[u]) continue;
    if (finite) apply16<KIND>(v[u], s, z, zp);
    else v[u] = apply16_generic(a.c, v[u], s, z);
    stg256(out + off + (int64_t)(u * blockDim.x + tid) * 16, v[u]);
  }
}

}  // namespace lcb

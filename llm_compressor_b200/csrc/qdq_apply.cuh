// Per-chunk building blocks of the streaming fake-quant: a 256-bit register chunk of 16 bf16 values (W8), its packed
// statistics and the packed quantise-dequantise arithmetic (apply16*).  Shared by the streaming kernels (qdq_stream.cuh),
// the flat two-pass kernels (qdq.cu) and the operand prologue of the fused activation-QDQ GEMM (qgemm.cu).
#pragma once

#include "qdq_fast.cuh"

namespace lcb {

struct W8 {
  uint32_t w[8];
};

__device__ __forceinline__ W8 ldg256(const void* p) {
  W8 r;
  asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                 "=r"(r.w[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(void* p, const W8& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]),
               "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7])
               : "memory");
}

// x * r snapped onto the bf16 tie grid (finite operands only: no NaN guard, see div_snap)
__device__ __forceinline__ float mul_snap(float x, float r) {
  return __uint_as_float((__float_as_uint(__fmul_rn(x, r)) + 4u) & 0xfffffff8u);
}

// packed element rounding of the float formats; A2 = two bf16 values of A = (x - z) / s, all finite.
// Returns false when a lane needs the reference arithmetic (fp8 sub-normal range).
template <int KIND>
__device__ __forceinline__ uint32_t core2(uint32_t A2) {
  const uint32_t ab = A2 & 0x7fff7fffu;
  uint32_t res;
  if constexpr (KIND == FK_E2M1) {
    uint32_t nb = (ab + 0x00200020u) & 0xffc0ffc0u;                       // 1 mantissa bit, ties away
    nb = bf22u(__hmin2(u2bf2(nb), u2bf2(0x40c040c0u)));                   // saturate at 6
    // |A| < 1: grid {0, 0.5, 1}: 0.5 * ([|A| >= 0.2490234] + [|A| >= 0.75])
    const __nv_bfloat162 c1 = __hge2(u2bf2(ab), u2bf2(0x3e7f3e7fu));
    const __nv_bfloat162 c2 = __hge2(u2bf2(ab), u2bf2(0x3f403f40u));
    const uint32_t sub = bf22u(__hmul2_rn(__hadd2_rn(c1, c2), u2bf2(0x3f003f00u)));
    const uint32_t small = __hlt2_mask(u2bf2(ab), u2bf2(0x3f803f80u));
    res = (sub & small) | (nb & ~small);
  } else if constexpr (KIND == FK_E4M3) {
    const uint32_t nb = (ab + 0x00080008u) & 0xfff0fff0u;                 // 3 mantissa bits
    res = bf22u(__hmin2(u2bf2(nb), u2bf2(0x43e043e0u)));                  // 448
  } else {
    const uint32_t nb = (ab + 0x00100010u) & 0xffe0ffe0u;                 // 2 mantissa bits
    res = bf22u(__hmin2(u2bf2(nb), u2bf2(0x47604760u)));                  // 57344
  }
  return res | (A2 & 0x80008000u);
}
// lanes of A2 in (0, 2^min_exp): the fixed-step range of the fp8 formats
template <int KIND>
__device__ __forceinline__ bool needs_subnormal(uint32_t A2) {
  if constexpr (KIND == FK_E4M3 || KIND == FK_E5M2) {
    const uint32_t ab = A2 & 0x7fff7fffu;
    const uint32_t lim = (KIND == FK_E4M3) ? 0x3c803c80u : 0x38803880u;  // 2^-6, 2^-14
    const uint32_t lt = __hlt2_mask(u2bf2(ab), u2bf2(lim));
    const uint32_t nz = __hgt2_mask(u2bf2(ab), u2bf2(0u));
    return (lt & nz) != 0u;
  } else {
    return false;
  }
}

template <int KIND>
__device__ __forceinline__ void apply16(W8& v, float s, float z, bool zp) {
  const float r = rcp_fast(s);
  const __nv_bfloat162 s2 = u2bf2(dup_bf(s));
  const __nv_bfloat162 z2 = u2bf2(dup_bf(z));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t w = v.w[i];
    if constexpr (KIND == FK_INT4 || KIND == FK_INT8) {
      const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
      __nv_bfloat162 q = u2bf2(pack_bf2(mul_snap(x0, r), mul_snap(x1, r)));
      q = __hadd2_rn(q, z2);
      if constexpr (KIND == FK_INT4) {
        q = __hmin2(__hmax2(q, u2bf2(0xc0e0c0e0u)), u2bf2(0x40e040e0u));  // clamp first: the bounds are integers
        const __nv_bfloat162 magic = u2bf2(0x43404340u);                     // 192: ulp(bf16) == 1 in [128, 256)
        q = __hsub2_rn(__hadd2_rn(q, magic), magic);                         // rint, ties to even
      } else {
        q = __hmin2(__hmax2(q, u2bf2(0xc2fec2feu)), u2bf2(0x42fe42feu));  // +-127
        const uint32_t qb = bf22u(q);
        const float MG = 12582912.0f;  // 1.5 * 2^23
        const float q0 = __fsub_rn(__fadd_rn(__uint_as_float(qb << 16), MG), MG);
        const float q1 = __fsub_rn(__fadd_rn(__uint_as_float(qb & 0xffff0000u), MG), MG);
        q = u2bf2(pack_bf2(q0, q1));
      }
      v.w[i] = bf22u(__hmul2_rn(__hsub2_rn(q, z2), s2));
    } else {
      if (zp) w = bf22u(__hsub2_rn(u2bf2(w), z2));  // a = bf16(x - z)
      const float a0 = __uint_as_float(w << 16), a1 = __uint_as_float(w & 0xffff0000u);
      const uint32_t A2 = pack_bf2(mul_snap(a0, r), mul_snap(a1, r));  // A = bf16(a / s)
      uint32_t q2;
      if (needs_subnormal<KIND>(A2)) {
        constexpr int EB = (KIND == FK_E4M3) ? 4 : 5, MB = (KIND == FK_E4M3) ? 5 : 4;
        q2 = pack_bf2(core_bits<EB, MB>(__uint_as_float(A2 << 16)), core_bits<EB, MB>(__uint_as_float(A2 & 0xffff0000u)));
      } else {
        q2 = core2<KIND>(A2);
      }
      v.w[i] = bf22u(__hadd2_rn(__hmul2_rn(u2bf2(q2), s2), z2));  // bf16(bf16(q * s) + z)
    }
  }
}

// Per-tensor INT scaling: the scale is a 0-dim fp32 tensor in the reference (int_quant.py:93-102 on a 0-dim amax), so
// x / s is RNE_bf16(RN_fp32(float(x) / s_fp32)) with an arbitrary 24-bit divisor: the exact-tie argument of div_snap does
// not hold.  q0 = x * (1/s) is within 2 ulp of the quotient; whenever q0 lies within 16 ulp of a bf16 rounding tie the
// true IEEE quotient is recomputed (about one element in 2^11), everything else cannot round differently.  The zero
// point is an integer of magnitude <= 254 (exact in bf16), so + z / - z are the packed bf16 ops of apply16; only the
// final * s goes through fp32 again.  All operands finite (the caller falls back to the generic kernel otherwise).
template <int KIND>
__device__ __forceinline__ void apply16_tensor(W8& v, float s, float rinv, float z) {
  static_assert(KIND == FK_INT4 || KIND == FK_INT8, "per-tensor fast path: INT formats");
  const __nv_bfloat162 z2 = u2bf2(dup_bf(z));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t w = v.w[i];
    const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
    float q0 = __fmul_rn(x0, rinv), q1 = __fmul_rn(x1, rinv);
    if (((__float_as_uint(q0) & 0xffffu) - 0x7ff0u) < 0x20u) q0 = __fdiv_rn(x0, s);
    if (((__float_as_uint(q1) & 0xffffu) - 0x7ff0u) < 0x20u) q1 = __fdiv_rn(x1, s);
    __nv_bfloat162 q = u2bf2(pack_bf2(q0, q1));
    q = __hadd2_rn(q, z2);
    float d0, d1;
    if constexpr (KIND == FK_INT4) {
      q = __hmin2(__hmax2(q, u2bf2(0xc0e0c0e0u)), u2bf2(0x40e040e0u));
      const __nv_bfloat162 magic = u2bf2(0x43404340u);
      q = __hsub2_rn(__hsub2_rn(__hadd2_rn(q, magic), magic), z2);
      const uint32_t qb = bf22u(q);
      d0 = __uint_as_float(qb << 16); d1 = __uint_as_float(qb & 0xffff0000u);
    } else {
      q = __hmin2(__hmax2(q, u2bf2(0xc2fec2feu)), u2bf2(0x42fe42feu));
      const uint32_t qb = bf22u(q);
      const float MG = 12582912.0f;
      const float r0 = __fsub_rn(__fadd_rn(__uint_as_float(qb << 16), MG), MG);
      const float r1 = __fsub_rn(__fadd_rn(__uint_as_float(qb & 0xffff0000u), MG), MG);
      const uint32_t db = bf22u(__hsub2_rn(u2bf2(pack_bf2(r0, r1)), z2));
      d0 = __uint_as_float(db << 16); d1 = __uint_as_float(db & 0xffff0000u);
    }
    v.w[i] = pack_bf2(__fmul_rn(d0, s), __fmul_rn(d1, s));
  }
}

// op-by-op reference arithmetic for a chunk whose group parameters are not finite
// (by value on purpose: a reference would force the caller's registers / kernel parameters into local memory)
static __device__ __noinline__ W8 apply16_generic(QCfg c, W8 v, float s, float z) {
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    float code;
    const float o0 = fake_quant<LCB_BF16>(c, __uint_as_float(v.w[i] << 16), s, z, code);
    const float o1 = fake_quant<LCB_BF16>(c, __uint_as_float(v.w[i] & 0xffff0000u), s, z, code);
    v.w[i] = (__float_as_uint(o0) >> 16) | (__float_as_uint(o1) & 0xffff0000u);
  }
  return v;
}

// statistics of 16 packed values: pair of (max, -min) for asymmetric, |x| maximum otherwise
__device__ __forceinline__ Stat2 stats16(const W8& v, bool zp) {
  Stat2 s;
  if (zp) {
    __nv_bfloat162 mx = __hmax2_nan(u2bf2(v.w[0]), u2bf2(v.w[1])), mn = __hmin2_nan(u2bf2(v.w[0]), u2bf2(v.w[1]));
#pragma unroll
    for (int i = 2; i < 8; ++i) {
      mx = __hmax2_nan(mx, u2bf2(v.w[i]));
      mn = __hmin2_nan(mn, u2bf2(v.w[i]));
    }
    const __nv_bfloat16 m1 = __hmax_nan(mx.x, mx.y);
    const __nv_bfloat16 m2 = __hneg(__hmin_nan(mn.x, mn.y));
    s.mxmn = bf22u(__halves2bfloat162(m1, m2));
    s.amax = 0;
  } else {
    __nv_bfloat162 am = __hmax2_nan(u2bf2(v.w[0] & 0x7fff7fffu), u2bf2(v.w[1] & 0x7fff7fffu));
#pragma unroll
    for (int i = 2; i < 8; ++i) am = __hmax2_nan(am, u2bf2(v.w[i] & 0x7fff7fffu));
    s.amax = bf22u(am);
    s.mxmn = 0;
  }
  return s;
}

}  // namespace lcb

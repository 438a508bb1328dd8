// Bandwidth-oriented fast paths of the fused fake-quant for bf16 tensors (axis -1).
//
// Same results as the generic arithmetic in qmath.cuh (bit for bit: the parity tests run both),
// with ~5x fewer instructions per element so that the kernel is HBM-bound instead of issue-bound:
//   * statistics with packed bf16x2 NaN-propagating max/min (2 elements per instruction);
//   * x / s without a division: q0 = x * rcp(s) is within 2 fp32 ulp of the quotient; for bf16 x
//     and bf16 s (8-bit significands) the exact quotient is either EXACTLY on a bf16 rounding
//     tie or at least 1/(255*512) ~ 7.6e-6 (relative) away from every tie, so when q0 lands
//     within 4 ulp of a tie the quotient IS that tie and q0 is snapped onto it; the following
//     round-to-nearest-even to bf16 then equals torch's RNE_bf16(RN_fp32(x / s)) in all cases;
//   * packed bf16x2 add / mul / min / max for the ops whose operands are bf16 on both sides
//     (one rounding to bf16 per op, exactly what torch does), magic-number rint for int4;
//   * the element rounding of the float formats on the bit pattern (add half, mask) instead of
//     the scale / floor / rescale sequence.
#pragma once

#include "qmath.cuh"

namespace lcb {

enum FastKind { FK_INT4 = 0, FK_INT8 = 1, FK_E2M1 = 2, FK_E4M3 = 3, FK_E5M2 = 4 };

__device__ __forceinline__ __nv_bfloat162 u2bf2(uint32_t u) { return *reinterpret_cast<__nv_bfloat162*>(&u); }
__device__ __forceinline__ uint32_t bf22u(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {  // cvt.rn.bf16x2.f32 (RNE)
  return bf22u(__floats2bfloat162_rn(lo, hi));
}
__device__ __forceinline__ uint32_t dup_bf(float v) { return pack_bf2(v, v); }

// reciprocal good to ~1 ulp; s is a clamped scale (>= 1e-5) or NaN / inf
__device__ __forceinline__ float rcp_fast(float s) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(s));
  return r;
}

// RNE_bf16(RN_fp32(x / s)) for bf16-valued x and s, returned as the fp32 to be rounded to bf16:
// q0 = x * r with r ~ 1/s is within 2 ulp of the quotient.  A true quotient is either exactly a
// bf16 tie (a multiple of 2^15 ulp, so also of 8) or >= 64 ulp away from every tie; rounding q0 to
// the nearest multiple of 8 ulp therefore puts the tie cases exactly on the tie (where the
// following RNE conversion breaks it to even, like torch) and cannot move any other value across one.
__device__ __forceinline__ float div_snap(float x, float r) {
  const float q0 = __fmul_rn(x, r);
  const uint32_t b = __float_as_uint(q0);
  const float snapped = __uint_as_float((b + 4u) & 0xfffffff8u);
  return (q0 != q0) ? q0 : snapped;  // the +4 would carry a NaN pattern into the sign bit
}
// statistics of 8 packed bf16 values; pair = (max, -min) and amax as bf16x2 lanes to be reduced
struct Stat2 {
  uint32_t mxmn;  // lo: max, hi: -min  (bf16x2)
  uint32_t amax;  // both lanes: partial |x| maxima (bf16x2)
};
__device__ __forceinline__ Stat2 stats8(const uint4& v, bool zp) {
  Stat2 s;
  const __nv_bfloat162 a = u2bf2(v.x), b = u2bf2(v.y), c = u2bf2(v.z), d = u2bf2(v.w);
  if (zp) {
    const __nv_bfloat162 mx = __hmax2_nan(__hmax2_nan(a, b), __hmax2_nan(c, d));
    const __nv_bfloat162 mn = __hmin2_nan(__hmin2_nan(a, b), __hmin2_nan(c, d));
    const __nv_bfloat16 m1 = __hmax_nan(mx.x, mx.y);
    const __nv_bfloat16 m2 = __hneg(__hmin_nan(mn.x, mn.y));
    s.mxmn = bf22u(__halves2bfloat162(m1, m2));
    s.amax = 0;
  } else {
    const __nv_bfloat162 am = __hmax2_nan(__hmax2_nan(__habs2(a), __habs2(b)), __hmax2_nan(__habs2(c), __habs2(d)));
    s.amax = bf22u(am);
    s.mxmn = 0;
  }
  return s;
}
__device__ __forceinline__ void stat_combine(Stat2& s, const Stat2& o, bool zp) {
  if (zp) s.mxmn = bf22u(__hmax2_nan(u2bf2(s.mxmn), u2bf2(o.mxmn)));
  else s.amax = bf22u(__hmax2_nan(u2bf2(s.amax), u2bf2(o.amax)));
}
__device__ __forceinline__ void stat_shfl_xor(Stat2& s, int off, bool zp) {
  if (zp) s.mxmn = bf22u(__hmax2_nan(u2bf2(s.mxmn), u2bf2(__shfl_xor_sync(0xffffffffu, s.mxmn, off))));
  else s.amax = bf22u(__hmax2_nan(u2bf2(s.amax), u2bf2(__shfl_xor_sync(0xffffffffu, s.amax, off))));
}
__device__ __forceinline__ void stat_finish(const Stat2& s, bool zp, float& mx, float& mn, float& amax) {
  if (zp) {
    const __nv_bfloat162 p = u2bf2(s.mxmn);
    mx = __bfloat162float(p.x);
    mn = -__bfloat162float(p.y);
    amax = 0.0f;
  } else {
    const __nv_bfloat162 p = u2bf2(s.amax);
    amax = __bfloat162float(__hmax_nan(p.x, p.y));
    mx = amax; mn = -amax;
  }
}

// INT find_params in bf16 (== int_params<LCB_BF16, LCB_BF16>) without divisions.  The divisors 7 / 14 /
// 127 / 254 and the bf16 scale all have <= 8 significant bits, so div_snap is exact (see above).
__device__ __forceinline__ void int_params_fast(float mx, float mn, float amax, int zp, const Fmt& f, float& s, float& z) {
  const float c = scale_floor<LCB_BF16>();
  if (zp) {
    const float range = R<LCB_BF16>(__fsub_rn(mx, mn));
    const float r14 = (f.mbits == 4) ? (1.0f / 14.0f) : (1.0f / 254.0f);
    const float s0 = R<LCB_BF16>(div_snap(range, r14));
    float t;
    if (s0 >= 1e-30f) t = R<LCB_BF16>(div_snap(mn, rcp_fast(s0)));
    else t = R<LCB_BF16>(__fdiv_rn(mn, s0));  // zero / denormal / NaN scale: exact IEEE semantics
    z = rintf(R<LCB_BF16>(__fsub_rn(-f.qmax, t)));
    s = clamp_min_nan(s0, c);
  } else {
    const float r7 = (f.mbits == 4) ? (1.0f / 7.0f) : (1.0f / 127.0f);
    s = clamp_min_nan(R<LCB_BF16>(div_snap(amax, r7)), c);
    z = 0.0f;
  }
}

// element rounding of a float format on the bf16-valued fp32 A (== elem_core<LCB_BF16>)
template <int EB, int MB>
__device__ __forceinline__ float core_bits(float A) {
  constexpr int M = MB - 2;                       // mantissa bits kept
  constexpr int MIN_EXP = 2 - (1 << (EB - 1));
  constexpr uint32_t MAXN = (EB == 2) ? 0x40c00000u : (EB == 4 ? 0x43e00000u : 0x47600000u);  // 6, 448, 57344
  const uint32_t b = __float_as_uint(A);
  const uint32_t ab = b & 0x7fffffffu;
  // normal range of the target format: add half an ulp of the kept mantissa, drop the rest (ties away)
  uint32_t nb = (ab + (1u << (22 - M))) & (0xffffffffu << (23 - M));
  // below 2^MIN_EXP: fixed grid of step 2^(MIN_EXP - M)
  const float y = __fmul_rn(__uint_as_float(ab), exp2i(M - MIN_EXP));
  // the reference rounds (|y| + 0.5) to bf16 before the floor: 0.5 - 2^-9 + 0.5 ties up to 1.0
  const float sub = __fmul_rn(floorf(R<LCB_BF16>(__fadd_rn(y, 0.5f))), exp2i(MIN_EXP - M));
  const uint32_t lim = (uint32_t)(MIN_EXP + 127) << 23;
  uint32_t rb = (ab >= lim) ? nb : __float_as_uint(sub);
  rb = min(rb, MAXN);
  rb = (ab >= 0x7f580000u) ? 0x7fc00000u : rb;  // reference quirk: bf16 log2 overflow -> NaN (qmath.cuh)
  rb = (ab >= 0x7f800000u) ? ab : rb;           // inf passes through, NaN stays NaN
  return __uint_as_float(rb | (b & 0x80000000u));
}

// Quantise-dequantise 8 packed bf16 values with scale s, zero z (bf16-valued floats), r ~ 1/s.
template <int KIND>
__device__ __forceinline__ uint4 apply8(const uint4& in, float s, float z, float r, bool zp) {
  const uint32_t w[4] = {in.x, in.y, in.z, in.w};
  uint32_t o[4];
  const __nv_bfloat162 s2 = u2bf2(dup_bf(s));
  const __nv_bfloat162 z2 = u2bf2(dup_bf(z));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x0 = __uint_as_float(w[i] << 16), x1 = __uint_as_float(w[i] & 0xffff0000u);
    if constexpr (KIND == FK_INT4) {
      // q = clamp(rint(bf16(bf16(x/s) + z)), -7, 7); out = bf16(bf16(q - z) * s)
      __nv_bfloat162 q = u2bf2(pack_bf2(div_snap(x0, r), div_snap(x1, r)));
      q = __hadd2_rn(q, z2);
      const __nv_bfloat162 lo = u2bf2(0xc0e0c0e0u), hi = u2bf2(0x40e040e0u);  // -7, 7
      q = __hmin2_nan(__hmax2_nan(q, lo), hi);                                   // clamp first: bounds are integers
      const __nv_bfloat162 magic = u2bf2(0x43404340u);                           // 192: ulp(bf16) == 1 in [128, 256)
      q = __hsub2_rn(__hadd2_rn(q, magic), magic);                                     // rint, ties to even
      o[i] = bf22u(__hmul2_rn(__hsub2_rn(q, z2), s2));
    } else if constexpr (KIND == FK_INT8) {
      __nv_bfloat162 q = u2bf2(pack_bf2(div_snap(x0, r), div_snap(x1, r)));
      q = __hadd2_rn(q, z2);
      const __nv_bfloat162 lo = u2bf2(0xc2fec2feu), hi = u2bf2(0x42fe42feu);  // -127, 127
      q = __hmin2_nan(__hmax2_nan(q, lo), hi);
      const uint32_t qb = bf22u(q);
      const float MG = 12582912.0f;  // 1.5 * 2^23: fp32 magic rint (ties to even)
      const float q0 = __fsub_rn(__fadd_rn(__uint_as_float(qb << 16), MG), MG);
      const float q1 = __fsub_rn(__fadd_rn(__uint_as_float(qb & 0xffff0000u), MG), MG);
      // (q - z) is an integer of magnitude <= 254: exact in bf16 when |.| <= 256 -> packed sub is exact
      o[i] = bf22u(__hmul2_rn(__hsub2_rn(u2bf2(pack_bf2(q0, q1)), z2), s2));
    } else {
      float a0 = x0, a1 = x1;
      if (zp) {  // a = bf16(x - z)
        const uint32_t t = bf22u(__hsub2_rn(u2bf2(w[i]), z2));
        a0 = __uint_as_float(t << 16); a1 = __uint_as_float(t & 0xffff0000u);
      }
      const uint32_t Ab = pack_bf2(div_snap(a0, r), div_snap(a1, r));  // A = bf16(a / s)
      float q0, q1;
      if constexpr (KIND == FK_E2M1) {
        q0 = core_bits<2, 3>(__uint_as_float(Ab << 16)); q1 = core_bits<2, 3>(__uint_as_float(Ab & 0xffff0000u));
      } else if constexpr (KIND == FK_E4M3) {
        q0 = core_bits<4, 5>(__uint_as_float(Ab << 16)); q1 = core_bits<4, 5>(__uint_as_float(Ab & 0xffff0000u));
      } else {
        q0 = core_bits<5, 4>(__uint_as_float(Ab << 16)); q1 = core_bits<5, 4>(__uint_as_float(Ab & 0xffff0000u));
      }
      // out = bf16(bf16(q * s) + z); q is bf16-exact, so the packed multiply rounds once like torch
      __nv_bfloat162 t = __hmul2_rn(u2bf2(pack_bf2(q0, q1)), s2);
      t = __hadd2_rn(t, z2);  // z == +0 for symmetric: only the sign of a zero result can change, as in torch
      o[i] = bf22u(t);
    }
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace lcb

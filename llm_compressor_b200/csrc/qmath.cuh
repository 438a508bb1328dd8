// Per-group parameter search and per-element fake-quant arithmetic, shared by the fake-quant
// kernels (qdq.cu) and the GPTQ block kernels (solver.cu).
//
// Every function states the reference lines it reproduces (under
// /root/reference/llm_compressor/quantization/quantizers/).  Arithmetic contract: each primitive
// torch op is one IEEE fp32 op (no FMA contraction -> __f*_rn intrinsics) followed by R<DT>().
#pragma once

#include "common.cuh"

namespace lcb {

struct Fmt {
  int ebits, mbits, emax, min_exp;
  float max_norm;  // largest normal of the element format (formats.py:83-86)
  float qmax;      // INT: max_norm * 2^(mbits-2)  (int_quant.py:55-57)
  float scale_m;   // 2^(mbits-2)
};

// formats.py:41-92
__host__ __device__ inline Fmt make_fmt(int elem) {
  Fmt f{};
  switch (elem) {
    case LCB_E_INT4: f.ebits = 0; f.mbits = 4; f.emax = 0; f.max_norm = 1.75f; break;
    case LCB_E_INT8: f.ebits = 0; f.mbits = 8; f.emax = 0; f.max_norm = 127.0f / 64.0f; break;
    case LCB_E_FP4_E2M1: f.ebits = 2; f.mbits = 3; f.emax = 2; f.max_norm = 6.0f; break;
    case LCB_E_FP8_E4M3: f.ebits = 4; f.mbits = 5; f.emax = 8; f.max_norm = 448.0f; break;
    default: f.ebits = 5; f.mbits = 4; f.emax = 15; f.max_norm = 57344.0f; break;
  }
  f.min_exp = f.ebits ? 2 - (1 << (f.ebits - 1)) : 0;
  f.scale_m = (float)(1 << (f.mbits - 2));
  f.qmax = f.max_norm * f.scale_m;
  return f;
}

struct QCfg {
  int qtype;
  int zero_point;
  float scale_emax;  // MX: 2^(scale_ebits-1) - 1
  Fmt f;
};

// ---------------------------------------------------------------------------------------------
// utils.py:218-284 _quantize_elemwise_core(round="nearest", saturate_normals=True).
//
// The reference derives the private exponent as floor(log2(|A|)) evaluated in A's dtype, which
// for bf16 can round up to the next integer when |A| is just below a power of two.  Inside the
// saturating range of every supported format the result is identical to using the exact binary
// exponent (the value rounds to that power of two on either grid; proof in DESIGN.md, checked
// exhaustively over all 65536 bf16 inputs by tests), so the exponent is taken from the bits.
template <int DT>
__device__ __forceinline__ float elem_core(float A, const Fmt& f) {
  float out;
  if (f.ebits != 0) {
    uint32_t ab = __float_as_uint(A) & 0x7fffffffu;
    if (ab >= 0x7f800000u) return A;  // +-inf pass through (utils.py:280-281), NaN propagates
    if constexpr (DT == LCB_BF16) {
      // |A| >= 2^127 * 1.6875: the reference's bf16 log2 rounds to 128, 2^128 overflows to inf and
      // the rescaling yields 0 * inf = NaN (80 of the 65536 bf16 patterns).  Reproduced on purpose.
      if (ab >= 0x7f580000u) return __uint_as_float(0x7fc00000u);
    }
    int e = (int)(ab >> 23) - 127;
    e = max(e, f.min_exp);
    e = min(e, 30);  // anything this large saturates to max_norm below
    // y = A / 2^e * 2^(mbits-2): exact power-of-two scalings
    float y = __fmul_rn(A, exp2i(f.mbits - 2 - e));
    float t = R<DT>(__fadd_rn(fabsf(y), 0.5f));
    float fl = floorf(t);
    out = copysignf(__fmul_rn(fl, exp2i(e - (f.mbits - 2))), A);
  } else {
    float y = R<DT>(__fmul_rn(A, f.scale_m));
    float t = R<DT>(__fadd_rn(fabsf(y), 0.5f));
    float fl = copysignf(floorf(t), y);
    fl = (y != y) ? y : fl;
    out = R<DT>(__fdiv_rn(fl, f.scale_m));
    if (fabsf(A) == INFINITY) return A;
  }
  return clamp_nan(out, -f.max_norm, f.max_norm);
}

// ---------------------------------------------------------------------------------------------
// find_params bodies.  DT: tensor dtype (first tensor-tensor op), PDT: dtype of the parameter
// arithmetic; PDT != DT only for per-tensor quantisation, where max/min are 0-dim tensors and
// torch promotes `0-dim bf16 (op) 0-dim fp32 buffer` to fp32.

// int_quant.py:90-112,164
// CLAMP = false: the scale before clamp(min=1e-5), as the mse clip search uses it (int_quant.py:129-135)
template <int DT, int PDT, bool CLAMP = true>
__device__ __forceinline__ void int_params(float mx, float mn, float amax, int zp, const Fmt& f, float& s, float& z) {
  if (zp) {
    float range = R<DT>(__fsub_rn(mx, mn));
    float s0 = R<PDT>(__fdiv_rn(range, __fsub_rn(f.qmax, -f.qmax)));
    float t = R<PDT>(__fdiv_rn(mn, s0));
    z = rintf(R<PDT>(__fsub_rn(-f.qmax, t)));
    s = CLAMP ? clamp_min_nan(s0, scale_floor<PDT>()) : s0;
  } else {
    const float s0 = R<PDT>(__fdiv_rn(amax, f.qmax));
    s = CLAMP ? clamp_min_nan(s0, scale_floor<PDT>()) : s0;
    z = 0.0f;
  }
}

// fp_quant.py:102-124,176
template <int DT, int PDT, bool CLAMP = true>
__device__ __forceinline__ void fp_params(float mx, float mn, float amax, int zp, const Fmt& f, float& s, float& z) {
  float s0;
  if (zp) {
    float range = R<DT>(__fsub_rn(mx, mn));
    s0 = R<PDT>(__fdiv_rn(range, __fmul_rn(2.0f, f.max_norm)));
    z = R<DT>(__fmul_rn(R<DT>(__fadd_rn(mx, mn)), 0.5f));
  } else {
    s0 = R<PDT>(__fdiv_rn(amax, f.max_norm));
    z = 0.0f;
  }
  s = CLAMP ? clamp_min_nan(s0, scale_floor<PDT>()) : s0;
}

// mx_quant.py:88-101: 2^clamp(floor(log2(v)) - emax_elem).  log2 is evaluated in the tensor
// dtype (bf16 rounds the logarithm before the floor).
template <int DT>
__device__ __forceinline__ float mx_shared_scale(float v, const QCfg& c) {
  float add = (v == 0.0f) ? 1.17549435e-38f : 0.0f;
  float t = R<DT>(__fadd_rn(v, add));
  float se = floorf(R<DT>(log2f(t)));
  se = R<DT>(__fsub_rn(se, (float)c.f.emax));
  se = (se > c.scale_emax) ? c.scale_emax + 1.0f : se;
  se = (se < -c.scale_emax) ? -c.scale_emax : se;
  return R<DT>(exp2f(se));  // integer argument: exact (2^128 -> inf, 2^-127 -> denormal)
}

template <int DT, bool CLAMP = true>
__device__ __forceinline__ void mx_params(float mx, float mn, float amax, const QCfg& c, float& s, float& z) {
  float v = amax;
  z = 0.0f;
  if (c.zero_point) {
    z = R<DT>(__fmul_rn(R<DT>(__fadd_rn(mx, mn)), 0.5f));
    v = R<DT>(__fsub_rn(mx, z));
  }
  const float s0 = mx_shared_scale<DT>(v, c);
  s = CLAMP ? clamp_min_nan(s0, scale_floor<DT>()) : s0;
}

// nvfp_quant.py:85-111.  Block statistic whose |.| maximum over the tensor is the global amax.
template <int DT>
__device__ __forceinline__ void nvfp_block_stat(float mx, float mn, float amax, int zp, float& v, float& z) {
  v = amax;
  z = 0.0f;
  if (zp) {
    z = R<DT>(__fmul_rn(R<DT>(__fadd_rn(mx, mn)), 0.5f));
    v = R<DT>(__fsub_rn(mx, z));
  }
}
// s32 = g / (448 * 6) stays fp32 (0-dim / 0-dim); the fp8 block scale is rounded in DT
template <int DT, bool CLAMP = true>
__device__ __forceinline__ float nvfp_scale(float v, float g_amax, const Fmt& f) {
  const Fmt f8 = make_fmt(LCB_E_FP8_E4M3);
  float s32 = __fdiv_rn(g_amax, __fmul_rn(f8.max_norm, f.max_norm));
  float m = R<DT>(__fdiv_rn(v, __fmul_rn(s32, f.max_norm)));
  float s8 = elem_core<DT>(m, f8);
  const float s0 = R<DT>(__fmul_rn(s8, s32));
  return CLAMP ? clamp_min_nan(s0, scale_floor<DT>()) : s0;
}

// parameters before the final clamp (candidates of the mse clip search); NVFP: symmetric only, nv_gamax = the
// (shrunken) whole-tensor amax
template <int DT>
__device__ __forceinline__ void find_params_unclamped(const QCfg& c, float mx, float mn, float amax, float nv_gamax,
                                                      float& s, float& z) {
  switch (c.qtype) {
    case LCB_Q_INT: int_params<DT, DT, false>(mx, mn, amax, c.zero_point, c.f, s, z); break;
    case LCB_Q_FP: fp_params<DT, DT, false>(mx, mn, amax, c.zero_point, c.f, s, z); break;
    case LCB_Q_MX: mx_params<DT, false>(mx, mn, amax, c, s, z); break;
    default: s = nvfp_scale<DT, false>(amax, nv_gamax, c.f); z = 0.0f; break;
  }
}

template <int DT, int PDT>
__device__ __forceinline__ void find_params(const QCfg& c, float mx, float mn, float amax, float nv_gamax, float& s,
                                            float& z) {
  switch (c.qtype) {
    case LCB_Q_INT: int_params<DT, PDT>(mx, mn, amax, c.zero_point, c.f, s, z); break;
    case LCB_Q_FP: fp_params<DT, PDT>(mx, mn, amax, c.zero_point, c.f, s, z); break;
    case LCB_Q_MX: mx_params<DT>(mx, mn, amax, c, s, z); break;
    default: {
      float v;
      nvfp_block_stat<DT>(mx, mn, amax, c.zero_point, v, z);
      s = nvfp_scale<DT>(v, nv_gamax, c.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fake_quantize bodies.  `code` receives the integer code / grid value.

// int_quant.py:210-212
template <int DT>
__device__ __forceinline__ float int_fq(float x, float s, float z, const Fmt& f, float& code) {
  float q = R<DT>(__fdiv_rn(x, s));
  q = R<DT>(__fadd_rn(q, z));
  q = clamp_nan(rintf(q), -f.qmax, f.qmax);
  code = q;
  return R<DT>(__fmul_rn(R<DT>(__fsub_rn(q, z)), s));
}

// fp_quant.py:222-234 (same body in mx_quant.py:189-201, nvfp_quant.py:188-200)
template <int DT>
__device__ __forceinline__ float fp_fq(float x, float s, float z, const Fmt& f, float& code) {
  float a = R<DT>(__fsub_rn(x, z));
  a = R<DT>(__fdiv_rn(a, s));
  float q = elem_core<DT>(a, f);
  code = q;
  return R<DT>(__fadd_rn(R<DT>(__fmul_rn(q, s)), z));
}

template <int DT>
__device__ __forceinline__ float fake_quant(const QCfg& c, float x, float s, float z, float& code) {
  return (c.qtype == LCB_Q_INT) ? int_fq<DT>(x, s, z, c.f, code) : fp_fq<DT>(x, s, z, c.f, code);
}

// Bit pattern of a grid value in its storage format (low bits of a byte).
//   INT: two's complement int8 of the integer code.
//   MX-INT (ebits == 0 under MX): two's complement of q * 2^(mbits-2).
//   FP: sign | exponent | mantissa, bias = 1 - min_exp, subnormals below 2^min_exp.
__device__ __forceinline__ uint8_t encode_code(const QCfg& c, float q) {
  const Fmt& f = c.f;
  if (c.qtype == LCB_Q_INT) return (uint8_t)(int8_t)(int)q;
  if (f.ebits == 0) return (uint8_t)(int8_t)(int)(q * f.scale_m);
  uint32_t b = __float_as_uint(q);
  uint32_t sign = b >> 31;
  float aq = fabsf(q);
  int mant_bits = f.mbits - 2;
  int e = (int)((b & 0x7fffffffu) >> 23) - 127;
  uint32_t field, mant;
  if (!(aq >= exp2i(f.min_exp))) {  // zero / subnormal (NaN maps to 0)
    field = 0;
    mant = (uint32_t)(aq * exp2i(mant_bits - f.min_exp));
  } else {
    field = (uint32_t)(e - f.min_exp + 1);
    mant = (uint32_t)(aq * exp2i(mant_bits - e)) - (1u << mant_bits);
  }
  return (uint8_t)((sign << (f.ebits + mant_bits)) | (field << mant_bits) | mant);
}

}  // namespace lcb

/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the compression hot path of pjh5672/llm-compressor.
 *
 * This file is a scalar, single-purpose restatement of the reference's PyTorch arithmetic.
 * It is only ever called from tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py.  The product (llm_compressor_b200/) never links or
 * calls it; the product fails loudly when its CUDA library is missing.
 *
 * Parity status: the reference ships no golden vectors or asserting tests for this path
 * (SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
 * generated in the build container by oracle/gen_golden.py (imports /root/reference) and
 * committed under tests/golden/.  tests/test_oracle_golden.py checks this file against them.
 *
 * Value convention: tensors are passed as float32 arrays.  With dt == ORC_BF16 every input is
 * a bf16-representable float and every primitive op is followed by one round-to-nearest-even
 * to bf16 (this is what torch does for bf16 tensors: opmath in fp32, result rounded), see
 * SURVEY.md section 8a note N8.  With dt == ORC_F32 ops are IEEE fp32 without contraction.
 *
 * Reference lines restated (all under /root/reference/llm_compressor/):
 *   quantization/quantizers/int_quant.py:80-212   (INT find_params / fake_quantize)
 *   quantization/quantizers/fp_quant.py:92-234    (FP)
 *   quantization/quantizers/mx_quant.py:74-201    (MX)
 *   quantization/quantizers/nvfp_quant.py:72-200  (NVFP)
 *   quantization/quantizers/utils.py:85-167,218-284 (block reshape/zero pad, elementwise core)
 *   quantization/quantizers/formats.py:41-92      (format parameters)
 *   pruning/wanda/core.py:116-126, pruning/ria/core.py:118-126, pruning/magnitude/core.py:38-43
 *   pruning/sparsegpt/core.py:201-203             (mask selection)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_F32 0
#define ORC_BF16 1

#define Q_INT 0
#define Q_FP 1
#define Q_MX 2
#define Q_NVFP 3

#define F_INT4 1
#define F_INT8 2
#define F_FP4_E2M1 3
#define F_FP8_E4M3 4
#define F_FP8_E5M2 5

/* ---------------------------------------------------------------- rounding helpers */

static float bf16_rne(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) { /* inf / nan: keep class */
    if (u & 0x007fffffu) u |= 0x00400000u;
    u &= 0xffff0000u;
  } else {
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    u &= 0xffff0000u;
  }
  memcpy(&v, &u, 4);
  return v;
}

static inline float R(float v, int dt) { return dt == ORC_BF16 ? bf16_rne(v) : v; }

/* volatile temporaries forbid the C compiler from fusing mul+add (build also uses -ffp-contract=off) */
static inline float f_add(float a, float b) { volatile float r = a + b; return r; }
static inline float f_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float f_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float f_div(float a, float b) { volatile float r = a / b; return r; }

/* torch.clamp semantics: NaN propagates */
static inline float clampf(float v, float lo, float hi) {
  if (v != v) return v;
  if (v < lo) return lo;
  if (v > hi) return hi;
  return v;
}
static inline float clamp_min(float v, float lo) {
  if (v != v) return v;
  return v < lo ? lo : v;
}
/* torch amax / amin propagate NaN */
static inline float nmax(float a, float b) { if (a != a) return a; if (b != b) return b; return a > b ? a : b; }
static inline float nmin(float a, float b) { if (a != a) return a; if (b != b) return b; return a < b ? a : b; }

/* log2 in the tensor dtype: fp32 result correctly rounded (via double), then rounded to bf16 */
static inline float log2_dt(float v, int dt) { return R((float)log2((double)v), dt); }

/* ---------------------------------------------------------------- format table (formats.py:41-92) */
typedef struct { int ebits, mbits, emax; float max_norm; } fmt_t;

static fmt_t get_fmt(int elem) {
  fmt_t f;
  switch (elem) {
    case F_INT4: f.ebits = 0; f.mbits = 4; f.emax = 0; f.max_norm = 1.75f; break;
    case F_INT8: f.ebits = 0; f.mbits = 8; f.emax = 0; f.max_norm = 127.0f / 64.0f; break;
    case F_FP4_E2M1: f.ebits = 2; f.mbits = 3; f.emax = 2; f.max_norm = 6.0f; break;
    case F_FP8_E4M3: f.ebits = 4; f.mbits = 5; f.emax = 8; f.max_norm = 448.0f; break;
    default: f.ebits = 5; f.mbits = 4; f.emax = 15; f.max_norm = 57344.0f; break;
  }
  return f;
}

/* ---------------------------------------------------------------- utils.py:218-284 */
static float elem_core(float A, fmt_t f, int dt) {
  float out = A;
  float scale_m = ldexpf(1.0f, f.mbits - 2);
  if (f.ebits != 0) {
    float absA = fabsf(A);
    float t = R(f_add(absA, (A == 0.0f) ? 1.0f : 0.0f), dt);
    float pe = floorf(log2_dt(t, dt));
    float min_exp = (float)(2 - (1 << (f.ebits - 1)));
    pe = (pe != pe) ? pe : (pe < min_exp ? min_exp : pe); /* clip(min=) keeps NaN */
    float p2 = R(powf(2.0f, pe), dt);                     /* 2**exp in dtype */
    out = R(f_div(out, p2), dt);
    out = R(f_mul(out, scale_m), dt);
    /* round half away from zero: sign * floor(|y| + 0.5) */
    float sg = (out > 0.0f) ? 1.0f : ((out < 0.0f) ? -1.0f : out /* 0 or NaN */);
    float fl = floorf(R(f_add(fabsf(out), 0.5f), dt));
    out = R(f_mul(sg, fl), dt);
    out = R(f_div(out, scale_m), dt);
    out = R(f_mul(out, p2), dt);
  } else {
    out = R(f_mul(out, scale_m), dt);
    float sg = (out > 0.0f) ? 1.0f : ((out < 0.0f) ? -1.0f : out);
    float fl = floorf(R(f_add(fabsf(out), 0.5f), dt));
    out = R(f_mul(sg, fl), dt);
    out = R(f_div(out, scale_m), dt);
  }
  out = clampf(out, -f.max_norm, f.max_norm);
  if (A == INFINITY) out = INFINITY;
  if (A == -INFINITY) out = -INFINITY;
  return out;
}

/* exported for exhaustive tests of the elementwise core */
void orc_elem_core(const float* a, float* out, long n, int elem, int dt) {
  fmt_t f = get_fmt(elem);
  for (long i = 0; i < n; ++i) out[i] = elem_core(a[i], f, dt);
}

static inline float scale_floor(int dt) { return dt == ORC_BF16 ? bf16_rne(1e-5f) : 1e-5f; }

/* ---------------------------------------------------------------- per-group parameter search */
typedef struct { float mx, mn, amax; } stats_t;

static stats_t group_stats(const float* x, long n_valid, long n_group) {
  stats_t s;
  s.mx = -INFINITY; s.mn = INFINITY; s.amax = 0.0f;
  for (long i = 0; i < n_valid; ++i) {
    s.mx = nmax(s.mx, x[i]); s.mn = nmin(s.mn, x[i]); s.amax = nmax(s.amax, fabsf(x[i]));
  }
  if (n_valid < n_group) { /* zero padding takes part in min/max (utils.py:119-132) */
    s.mx = nmax(s.mx, 0.0f); s.mn = nmin(s.mn, 0.0f);
  }
  return s;
}

/* int_quant.py:90-112,164 */
/* dt: dtype of the tensor (first tensor-tensor op); pdt: dtype of the parameter arithmetic.  They
 * differ only for per-tensor quantisation, where max/min are 0-dim tensors and torch's type
 * promotion (0-dim bf16 op 0-dim fp32 buffer -> fp32) keeps scales / zeros in fp32. */
static void int_params(stats_t st, int zero_point, float qmax, int dt, int pdt, float* s_out, float* z_out, int clamp) {
  float s, z;
  if (zero_point) {
    float range = R(f_sub(st.mx, st.mn), dt);
    s = R(f_div(range, f_sub(qmax, -qmax)), pdt);
    float t = R(f_div(st.mn, s), pdt);
    z = rintf(R(f_sub(-qmax, t), pdt));
  } else {
    s = R(f_div(st.amax, qmax), pdt);
    z = 0.0f;
  }
  *s_out = clamp ? clamp_min(s, scale_floor(pdt)) : s;
  *z_out = z;
}

/* fp_quant.py:102-124,176 */
static void fp_params(stats_t st, int zero_point, float max_norm, int dt, int pdt, float* s_out, float* z_out, int clamp) {
  float s, z;
  if (zero_point) {
    float range = R(f_sub(st.mx, st.mn), dt);
    s = R(f_div(range, f_mul(2.0f, max_norm)), pdt);
    z = R(f_div(R(f_add(st.mx, st.mn), dt), 2.0f), dt); /* 0-dim / python scalar stays in dt */
  } else {
    s = R(f_div(st.amax, max_norm), pdt);
    z = 0.0f;
  }
  *s_out = clamp ? clamp_min(s, scale_floor(pdt)) : s;
  *z_out = z;
}

/* mx_quant.py:88-101 */
static float mx_scale(float val, int emax_elem, int scale_ebits, int dt) {
  float scale_emax = (float)((1 << (scale_ebits - 1)) - 1);
  float add = R(f_mul(1.17549435e-38f, (val == 0.0f) ? 1.0f : 0.0f), dt);
  float t = R(f_add(val, add), dt);
  float se = floorf(log2_dt(t, dt));
  se = R(f_sub(se, (float)emax_elem), dt);
  if (se > scale_emax) se = scale_emax + 1.0f;
  if (se < -scale_emax) se = -scale_emax;
  return R(powf(2.0f, se), dt);
}

static void mx_params(stats_t st, int zero_point, fmt_t f, int scale_ebits, int dt, float* s_out, float* z_out, int clamp) {
  float s, z;
  if (zero_point) {
    z = R(f_div(R(f_add(st.mx, st.mn), dt), 2.0f), dt);
    s = mx_scale(R(f_sub(st.mx, z), dt), f.emax, scale_ebits, dt);
  } else {
    z = 0.0f;
    s = mx_scale(st.amax, f.emax, scale_ebits, dt);
  }
  *s_out = clamp ? clamp_min(s, scale_floor(dt)) : s;
  *z_out = z;
}

/* ---------------------------------------------------------------- fake_quantize */
static float int_fq(float x, float s, float z, float qmax, int dt, float* code) {
  float q = R(f_div(x, s), dt);
  q = R(f_add(q, z), dt);
  q = clampf(rintf(q), -qmax, qmax);
  if (code) *code = q;
  return R(f_mul(R(f_sub(q, z), dt), s), dt);
}

static float fp_fq(float x, float s, float z, fmt_t f, int dt, float* code) {
  float a = R(f_sub(x, z), dt);
  a = R(f_div(a, s), dt);
  float q = elem_core(a, f, dt);
  if (code) *code = q;
  return R(f_add(R(f_mul(q, s), dt), z), dt);
}

/*
 * Row-major [rows, cols]; groups of `group` consecutive elements along cols (ragged tail is
 * zero padded for the statistics).  scales / zeros are [rows, G], G = ceil(cols / group).
 * params_given != 0: scales/zeros are inputs.  codes (optional) receives the integer code
 * (INT) or the unit-scale grid value (FP/MX/NVFP) as float.
 * nv_global: NVFP only.  NaN -> computed over the whole [rows, cols] input.
 * pdt: dtype of the scale / zero arithmetic (== dt except per-tensor, see int_params).
 */
int orc_qdq_rows(const float* x, float* out, float* scales, float* zeros, float* codes,
                 long rows, long cols, long group, int qtype, int elem, int zero_point,
                 int scale_ebits, int dt, int pdt, int params_given, float nv_global, int mse) {
  if (group <= 0) return -1;
  if (mse && ((qtype == Q_NVFP && zero_point) || cols % group != 0)) return -2;
  fmt_t f = get_fmt(elem);
  long G = (cols + group - 1) / group;
  float qmax = f.max_norm * ldexpf(1.0f, f.mbits - 2); /* int_quant.py:55-57 */
  float* blockmax = NULL;
  float* zbuf = NULL;
  float s32 = 0.0f, gmax_saved = 0.0f;
  fmt_t f8 = get_fmt(F_FP8_E4M3);

  if (!params_given && qtype == Q_NVFP) {
    /* nvfp_quant.py:85-111: block maxima first, then one whole-tensor amax */
    blockmax = (float*)malloc(sizeof(float) * (size_t)(rows * G));
    zbuf = (float*)malloc(sizeof(float) * (size_t)(rows * G));
    float g = 0.0f;
    for (long r = 0; r < rows; ++r)
      for (long b = 0; b < G; ++b) {
        long c0 = b * group, nv = (cols - c0 < group) ? cols - c0 : group;
        stats_t st = group_stats(x + r * cols + c0, nv, group);
        float z = 0.0f, v = st.amax;
        if (zero_point) {
          z = R(f_div(R(f_add(st.mx, st.mn), dt), 2.0f), dt);
          v = R(f_sub(st.mx, z), dt);
        }
        blockmax[r * G + b] = v; zbuf[r * G + b] = z;
        g = nmax(g, fabsf(v));
      }
    if (nv_global == nv_global) g = nv_global;
    gmax_saved = g;
    s32 = f_div(g, f_mul(f8.max_norm, f.max_norm)); /* fp32 even for bf16 tensors (0-dim / 0-dim) */
  }

  for (long r = 0; r < rows; ++r) {
    for (long b = 0; b < G; ++b) {
      long c0 = b * group, nv = (cols - c0 < group) ? cols - c0 : group;
      const float* xg = x + r * cols + c0;
      float s, z;
      if (params_given) {
        s = scales[r * G + b]; z = zeros[r * G + b];
      } else {
        if (qtype == Q_NVFP && mse) {
          /* nvfp_quant.py:113-148 (symmetric): candidate i shrinks every block maximum by p, so the tensor-wide
           * amax inside _get_scales shrinks to R(p * g) too (rounding is monotone) */
          float g = gmax_saved;
          float best = INFINITY;
          float v = blockmax[r * G + b];
          z = 0.0f;
          {
            float m = R(f_div(v, f_mul(s32, f.max_norm)), dt);
            s = R(f_mul(elem_core(m, f8, dt), s32), dt);
          }
          for (int it = 0; it < 80; ++it) {
            float p = (float)(1.0 - (double)it / 100.0);
            float v1 = R(f_mul(v, p), dt), g1 = R(f_mul(g, p), dt);
            float s32_1 = f_div(g1, f_mul(f8.max_norm, f.max_norm));
            float m1 = R(f_div(v1, f_mul(s32_1, f.max_norm)), dt);
            float sc1 = R(f_mul(elem_core(m1, f8, dt), s32_1), dt);
            float acc = 0.0f;
            for (long i = 0; i < nv; ++i) {
              float dq = fp_fq(xg[i], sc1, 0.0f, f, dt, NULL);
              float d = fabsf(R(f_sub(dq, xg[i]), dt));
              acc = f_add(acc, R(powf(d, R(2.4f, dt)), dt));
            }
            float err = R(acc, dt);
            if (err < best) { best = err; s = sc1; }
          }
          s = clamp_min(s, scale_floor(dt));
        } else if (qtype == Q_NVFP) {
          float m = R(f_div(blockmax[r * G + b], f_mul(s32, f.max_norm)), dt);
          float s8 = elem_core(m, f8, dt);
          s = clamp_min(R(f_mul(s8, s32), dt), scale_floor(dt));
          z = zbuf[r * G + b];
        } else {
          stats_t st = group_stats(xg, nv, group);
          if (!mse) {
            if (qtype == Q_INT) int_params(st, zero_point, qmax, dt, pdt, &s, &z, 1);
            else if (qtype == Q_FP) fp_params(st, zero_point, f.max_norm, dt, pdt, &s, &z, 1);
            else mx_params(st, zero_point, f, scale_ebits, dt, &s, &z, 1);
          } else {
            /* _clip_range (int_quant.py:115-162, fp_quant.py:127-174, mx_quant.py:114-149): 80 shrink steps,
             * candidate parameters NOT clamped, error |dq - x|^2.4 op by op in the tensor dtype, summed in
             * fp32 and rounded once; strict < keeps the first minimum; clamp(min=1e-5) at the very end. */
            if (!zero_point) { st.mx = st.amax; st.mn = -st.amax; }
            float best = INFINITY;
            if (qtype == Q_INT) int_params(st, zero_point, qmax, dt, pdt, &s, &z, 0);
            else if (qtype == Q_FP) fp_params(st, zero_point, f.max_norm, dt, pdt, &s, &z, 0);
            else mx_params(st, zero_point, f, scale_ebits, dt, &s, &z, 0);
            for (int it = 0; it < 80; ++it) {
              float p = (float)(1.0 - (double)it / 100.0);
              stats_t s1;
              s1.mx = R(f_mul(st.mx, p), dt); s1.mn = R(f_mul(st.mn, p), dt); s1.amax = s1.mx;
              float sc1, z1;
              if (qtype == Q_INT) int_params(s1, zero_point, qmax, dt, pdt, &sc1, &z1, 0);
              else if (qtype == Q_FP) fp_params(s1, zero_point, f.max_norm, dt, pdt, &sc1, &z1, 0);
              else mx_params(s1, zero_point, f, scale_ebits, dt, &sc1, &z1, 0);
              float acc = 0.0f;
              for (long i = 0; i < nv; ++i) {
                float dq = (qtype == Q_INT) ? int_fq(xg[i], sc1, z1, qmax, dt, NULL) : fp_fq(xg[i], sc1, z1, f, dt, NULL);
                float d = fabsf(R(f_sub(dq, xg[i]), dt));
                /* pow_(2.4) on a bf16 tensor: torch casts the python exponent to the tensor dtype first
                 * (bf16(2.4) = 2.40625) [measured against the reference on CPU: 8.0e-7 vs 8.3e-7 at 2.93e-3] */
                acc = f_add(acc, R(powf(d, R(2.4f, dt)), dt));
              }
              float err = R(acc, dt);
              if (err < best) { best = err; s = sc1; z = z1; }
            }
            s = clamp_min(s, scale_floor(pdt));
          }
        }
        if (scales) scales[r * G + b] = s;
        if (zeros) zeros[r * G + b] = z;
      }
      if (out) {
        for (long i = 0; i < nv; ++i) {
          float* cd = codes ? codes + r * cols + c0 + i : NULL;
          out[r * cols + c0 + i] = (qtype == Q_INT) ? int_fq(xg[i], s, z, qmax, dt, cd)
                                                    : fp_fq(xg[i], s, z, f, dt, cd);
        }
      }
    }
  }
  free(blockmax); free(zbuf);
  return 0;
}

/* ---------------------------------------------------------------- masks */
typedef struct { float v; int64_t i; } kv_t;
static int kv_cmp(const void* a, const void* b) {
  const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
  int nx = x->v != x->v, ny = y->v != y->v;
  if (nx != ny) return nx - ny;
  if (x->v < y->v) return -1;
  if (x->v > y->v) return 1;
  return (x->i > y->i) - (x->i < y->i);
}
/* torch.sort order: NaN compares greater than every number */
static int f_cmp(const void* a, const void* b) {
  float x = *(const float*)a, y = *(const float*)b;
  int nx = x != x, ny = y != y;
  if (nx || ny) return nx - ny;
  return (x > y) - (x < y);
}

/* wanda/core.py:116-126: metric fp32 = |W| * sqrt(scaler_row); per row the first int(K*ratio)
 * entries of a stable ascending sort are pruned. */
void orc_mask_wanda(const float* w, const float* scaler_row, uint8_t* mask, long rows, long cols, double ratio) {
  long k = (long)((double)cols * ratio);
  kv_t* buf = (kv_t*)malloc(sizeof(kv_t) * (size_t)cols);
  for (long r = 0; r < rows; ++r) {
    for (long c = 0; c < cols; ++c) {
      buf[c].v = f_mul(fabsf(w[r * cols + c]), sqrtf(scaler_row[c]));
      buf[c].i = c;
      mask[r * cols + c] = 0;
    }
    qsort(buf, (size_t)cols, sizeof(kv_t), kv_cmp);
    for (long j = 0; j < k; ++j) mask[r * cols + buf[j].i] = 1;
  }
  free(buf);
}

/* value at sorted index idx of an array (copy + qsort; oracle sizes are small) */
static float kth_sorted(const float* v, long n, long idx) {
  float* t = (float*)malloc(sizeof(float) * (size_t)n);
  memcpy(t, v, sizeof(float) * (size_t)n);
  qsort(t, (size_t)n, sizeof(float), f_cmp);
  float r = t[idx];
  free(t);
  return r;
}

/* magnitude/core.py:38-43 (also the threshold rule of sparsegpt/core.py:201-203 on a given metric) */
float orc_mask_threshold(const float* metric, uint8_t* mask, long n, double ratio) {
  long idx = (long)((double)n * ratio);
  if (idx >= n) idx = n - 1;
  float th = kth_sorted(metric, n, idx);
  for (long i = 0; i < n; ++i) mask[i] = metric[i] <= th;
  return th;
}

/* ria/core.py:118-126.  dt describes W's dtype (bf16 sums are rounded once). */
float orc_mask_ria(const float* w, const float* scaler_row, uint8_t* mask, long rows, long cols,
                   double ratio, float alpha, int dt) {
  double* cs = (double*)calloc((size_t)cols, sizeof(double));
  double* rs = (double*)calloc((size_t)rows, sizeof(double));
  for (long r = 0; r < rows; ++r)
    for (long c = 0; c < cols; ++c) { double a = fabsf(w[r * cols + c]); cs[c] += a; rs[r] += a; }
  float* m = (float*)malloc(sizeof(float) * (size_t)(rows * cols));
  for (long r = 0; r < rows; ++r)
    for (long c = 0; c < cols; ++c) {
      float a = fabsf(w[r * cols + c]);
      float t0 = R(f_div(a, R((float)cs[c], dt)), dt);
      float t1 = R(f_div(a, R((float)rs[r], dt)), dt);
      float base = R(f_add(t0, t1), dt);
      float sq = sqrtf(scaler_row[c]);
      float fac = (alpha == 0.5f) ? sqrtf(sq) : ((alpha == 1.0f) ? sq : powf(sq, alpha));
      m[r * cols + c] = f_mul(base, fac);
    }
  float th = orc_mask_threshold(m, mask, rows * cols, ratio);
  free(m); free(cs); free(rs);
  return th;
}

/* wanda/core.py:92-105 running row-norm: s *= n/(n+1); n += 1; s += ||x_k||^2 / n (x: [T, K]) */
void orc_rownorm_accum(float* s, const float* x, long T, long K, long nsamples_before) {
  /* python: s *= n/(n+1) -> double ratio cast to fp32 scalar; / n -> fp32 divide */
  float ratio = (float)((double)nsamples_before / (double)(nsamples_before + 1));
  float n1 = (float)(nsamples_before + 1);
  for (long k = 0; k < K; ++k) {
    double acc = 0.0;
    for (long t = 0; t < T; ++t) { double v = x[t * K + k]; acc += v * v; }
    float nrm = sqrtf((float)acc);
    s[k] = f_add(f_mul(s[k], ratio), f_div(f_mul(nrm, nrm), n1));
  }
}

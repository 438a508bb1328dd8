/*
 * lcb200.h -- C ABI of the B200-native (sm_100a) compression hot path.
 *
 * Drop-in boundary for the data-parallel compression path of pjh5672/llm-compressor
 * (reference tree: /root/reference/llm_compressor, cited below as ref:<file>:<lines>).
 * The reference has no FFI of its own for this path -- it is pure PyTorch -- so each entry point
 * replaces a group of tensor-op sequences inside one reference function.  The Python host side
 * (llm_compressor_b200/*.py) mirrors the reference's quantizer / solver API on top of these
 * calls through ctypes; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / CUDA types in the signatures.  `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`.
 *   - no hidden device allocation: scratch memory is passed in (`ws`, `ws_bytes`); each call has a
 *     `*_ws_bytes` query.  Calls are asynchronous on `stream`.  Two pieces of library state exist and are
 *     the only exceptions to re-entrancy: (1) lcb_set_gemm_mode() is a PROCESS-GLOBAL switch read by the
 *     solver entry points (set it before concurrent use, not during); (2) the look-ahead schedules of
 *     lcb_gptq_update / lcb_sparsegpt_update / lcb_gptaq_p and of the exact-mode lcb_chol_inv_upper use three
 *     library-owned non-blocking side streams + events, created once per host thread and device and joined
 *     back into `stream` before the call returns -- concurrent calls must come from different host threads
 *     (or be serialised on one), each with its own workspace.
 *   - return value: LCB_OK or a negative LCB_ERR_*; lcb_last_error() gives a thread-local
 *     message.  Numerical trouble found on the device (NaN scales, non-SPD Hessian) is reported
 *     through the optional device word `status` (bit mask LCB_ST_*), which the caller reads
 *     when it wants reference-equivalent exceptions (ref: int_quant.py:165, gptq/core.py:213-221).
 *   - matrices are row-major and contiguous unless a leading dimension is given.
 */
#ifndef LCB200_H_
#define LCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCB_ABI_VERSION 1

#define LCB_OK 0
#define LCB_ERR_INVALID (-1)     /* bad argument / geometry */
#define LCB_ERR_UNSUPPORTED (-2) /* valid in the reference but not implemented here yet */
#define LCB_ERR_CUDA (-3)        /* CUDA runtime / driver error, see lcb_last_error() */
#define LCB_ERR_WORKSPACE (-4)   /* ws_bytes too small */

/* device status bits */
#define LCB_ST_NAN_SCALE 1u /* a scale is NaN  (ref: int_quant.py:165 assert) */
#define LCB_ST_NOT_SPD 2u   /* Cholesky met a non-positive pivot (ref: gptq/core.py:213-221 retry) */

/* storage / arithmetic dtype of a tensor: every primitive op rounds to this type, like torch */
#define LCB_F32 0
#define LCB_BF16 1
#define LCB_F64 2 /* lcb_hadamard_rows only */

/* quantizer families (ref: quantization/quant.py:36-63) */
#define LCB_Q_INT 0  /* ref: quantizers/int_quant.py  */
#define LCB_Q_FP 1   /* ref: quantizers/fp_quant.py   */
#define LCB_Q_MX 2   /* ref: quantizers/mx_quant.py   */
#define LCB_Q_NVFP 3 /* ref: quantizers/nvfp_quant.py */

/* element formats, numbered like ElemFormat (ref: quantizers/formats.py:11-16) */
#define LCB_E_INT4 1
#define LCB_E_INT8 2
#define LCB_E_FP4_E2M1 3
#define LCB_E_FP8_E4M3 4
#define LCB_E_FP8_E5M2 5

typedef struct lcb_quant_cfg {
  int32_t qtype;       /* LCB_Q_* */
  int32_t elem;        /* LCB_E_* */
  int32_t zero_point;  /* asymmetric */
  int32_t scale_ebits; /* MX shared-exponent bits (ref: mx_quant.py:63), 8 */
  int32_t mse;         /* find_params with the clip search (ref: int_quant.py:115-162): 80 shrink steps, |.|^2.4 error;
                          INT / FP / MX / symmetric NVFP, axis -1, cols % group == 0; else LCB_ERR_UNSUPPORTED */
  int32_t reserved[3];
} lcb_quant_cfg;

/* mode bits of lcb_qdq */
#define LCB_QDQ_FIND 1  /* compute scales / zeros from x (find_params) */
#define LCB_QDQ_APPLY 2 /* write the fake-quantised tensor (fake_quantize) */

int lcb_abi_version(void);
const char* lcb_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench accounting) */
uint64_t lcb_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * (c) fused quantize-dequantize.
 * Replaces <Quantizer>.find_params / .forward / .fake_quantize together with
 * _reshape_to_blocks / _undo_reshape_to_blocks / _quantize_elemwise_core
 * (ref: int_quant.py:80-212, fp_quant.py:92-234, mx_quant.py:74-201, nvfp_quant.py:72-200,
 *  utils.py:85-167,218-284).
 *
 * x, out : [batch, rows, cols] contiguous, dtype `dtype`.
 * axis   : -1 groups of `group` consecutive elements along cols (ragged tail zero padded for
 *             the statistics, like utils.py:119-132); -2 groups of `group` rows, per column.
 * group  : > 0 group length; 0 = per tensor (INT / FP only; axis ignored).
 *          Per-token is group = cols with axis -1, per-channel group = rows with axis -2.
 * scales, zeros : block shaped, [batch*rows, G] (axis -1, G = ceil(cols/group)) or
 *          [batch, G, cols] (axis -2, G = ceil(rows/group)); dtype `dtype`; per tensor: one
 *          float32 each (torch keeps 0-dim parameters in fp32).  Outputs when mode has
 *          LCB_QDQ_FIND (may be NULL), inputs otherwise.
 * codes  : optional uint8 [batch, rows, cols]: the integer code (INT, two's complement int8) or
 *          the element's bit pattern in its fp4 / fp8 format (low bits).
 * NVFP needs the amax of the WHOLE tensor passed in (ref: nvfp_quant.py:87); with row-sharded
 * multi-GPU use lcb_nvfp_global_amax + an all-reduce(MAX) and pass the result in nv_amax
 * (device float, NULL = compute locally).
 */
size_t lcb_qdq_ws_bytes(const lcb_quant_cfg* cfg, int dtype, int64_t batch, int64_t rows, int64_t cols,
                        int axis, int64_t group);
int lcb_qdq(const lcb_quant_cfg* cfg, int dtype, int mode, const void* x, void* out, int64_t batch,
            int64_t rows, int64_t cols, int axis, int64_t group, void* scales, void* zeros, uint8_t* codes,
            const float* nv_amax, void* ws, size_t ws_bytes, uint32_t* status, void* stream);
/* phase 1 of NVFP alone: amax_out[0] = max over blocks of |block maximum| (float32) */
int lcb_nvfp_global_amax(const lcb_quant_cfg* cfg, int dtype, const void* x, int64_t batch, int64_t rows,
                         int64_t cols, int axis, int64_t group, float* amax_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * (a) calibration statistics.
 * lcb_hessian_accum replaces the body of cache_hessian_weight / Wrapper.cache_hessian_weight
 * (ref: gptq/core.py:103-119, sparsegpt/core.py:85-101) and, with dxxt / x_fp given,
 * cache_hessian_dxxt_weight (ref: gptaq/core.py:116-141):
 *     H = beta*H + alpha * X^T X          dXXT = beta*dXXT + alpha * (X_fp - X)^T X
 * X, X_fp: [tokens, k] bf16 row-major (token-major, as the hook receives them).
 * H, dXXT: [k, k] float32.  The reference's hook is beta = n/(n+1), alpha = 2/(n+1) (running
 * mean).  The faster equivalent keeps raw sums: beta = 1, alpha = 1 on every call and ONE
 * lcb_hessian_finalize(scale = 2/n) before the solver; with upper_only != 0 (requires beta == 1)
 * only the tiles touching the upper triangle of the symmetric X^T X are computed and
 * lcb_hessian_finalize(symmetric_from_upper = 1) mirrors them.  Token-sharded ranks all-reduce the
 * raw sums before finalising.
 * Requires k % 8 == 0, k <= 16384.  tcgen05 path: bf16 x bf16 -> fp32 in TMEM, TMA reduce-add into H.
 */
size_t lcb_hessian_ws_bytes(int64_t tokens, int64_t k);
int lcb_hessian_accum(float* H, float* dxxt, const void* x, const void* x_fp, int64_t tokens, int64_t k,
                      float alpha, float beta, int upper_only, void* ws, size_t ws_bytes, void* stream);
/* H *= scale; with symmetric_from_upper: H[i][j] = H[j][i] = scale * H[min(i,j)][max(i,j)] */
/* Raw-sum accumulation of up to 8 hook inputs in ONE launch:  H += alpha * sum_i X_i^T X_i  (every X_i [tokens, k] bf16,
 * xs = HOST array of `count` device pointers).  Equivalent to `count` lcb_hessian_accum(beta = 1) calls; each tile keeps
 * one tensor-core accumulation chain over all count * tokens tokens, so H crosses L2 / HBM once per launch instead of
 * once per hook call (K = 8192: the computed half of H is 140 MB > L2) and the per-launch prologue / drain is shared.
 * The chain is `count` times longer: the tensor core's truncating fp32 accumulation gives relF ~ 4e-6 at 8192 tokens
 * (1.6e-6 at 2048), a nearly uniform scale the solvers are invariant to. */
int lcb_hessian_accum_multi(float* H, const void* const* xs, int count, int64_t tokens, int64_t k, float alpha,
                            int upper_only, void* stream);
int lcb_hessian_finalize(float* H, int64_t k, float scale, int symmetric_from_upper, void* stream);
/* Token-sharded ranks (new work, the reference is single-device): only the upper triangle of the raw sums is
 * meaningful, so the all-reduce can carry HALF the bytes.  lcb_hessian_pack_upper copies the upper 32 x 32 blocks of S [k, k]
 * (block row bi, block column bj >= bi, nb = ceil(k / 32); block (bi, bj) at float 1024 * (bi*nb - bi*(bi-1)/2 + bj - bi),
 * row-major, zero beyond k) into `packed` (lcb_hessian_packed_floats(k) floats); after the all-reduce of `packed`,
 * lcb_hessian_finalize_packed writes H[i][j] = H[j][i] = scale * packed(min(i,j), max(i,j)) -- bit for bit what
 * lcb_hessian_finalize(symmetric_from_upper = 1) makes of the same sums. */
size_t lcb_hessian_packed_floats(int64_t k);
int lcb_hessian_pack_upper(const float* S, int64_t k, float* packed, void* stream);
int lcb_hessian_finalize_packed(const float* packed, float* H, int64_t k, float scale, void* stream);
/* ref: wanda/core.py:92-105, ria/core.py:94-107:  s = beta*s + alpha * sum_t X[t,:]^2 */
int lcb_rownorm_accum(float* s, const void* x, int64_t tokens, int64_t k, float alpha, float beta, void* stream);

/* ------------------------------------------------------------------------------------------
 * (b) layer solvers.
 * lcb_hessian_dead_fix: dead = diag(H) == 0; H[dead, dead] = 1 (ref: gptq/core.py:175-176); the
 * caller zeroes W[:, dead] (and dXXT[:, dead]) itself.  dead: uint8 [k] out (may be NULL).
 *
 * lcb_chol_inv_upper replaces damping + cholesky -> cholesky_inverse -> cholesky(upper)
 * (ref: gptq/core.py:207-224) and the act-order permutation of H (ref: gptq/core.py:181-201):
 * with Hp[i][j] = H[perm[i]][perm[j]] (perm NULL = identity) and Hp += damp * mean(diag(H)) * I,
 * U [k, k] (may alias H) receives the upper triangular factor with Hp^-1 = U^T U (strictly lower
 * part zeroed).
 * U is obtained as the inverse of the reverse-ordered Cholesky factor (H = R R^T, U = R^-1):
 * one blocked potrf + one blocked trtri instead of the reference's potrf + potri + potrf.
 * A non-positive pivot sets LCB_ST_NOT_SPD in *status (U is then garbage; the caller calls again
 * with the larger damping, reproducing the reference's retry).
 */
int lcb_hessian_dead_fix(float* H, int64_t k, uint8_t* dead, void* stream);
size_t lcb_chol_ws_bytes(int64_t k);
/* development aid: with LCB_CHOL_TRACE=1 in the environment the tile-task kernel leaves 8 x uint64 per task
 * (type|row|col, then globaltimer ns at fetch / accumulate-done / diagonal-flag / factor-done / end) at this byte
 * offset of the workspace. */
size_t lcb_chol_trace_offset(int64_t k);
int lcb_chol_inv_upper(const float* H, float* U, int64_t k, const int64_t* perm, float damp, void* ws,
                       size_t ws_bytes, uint32_t* status, void* stream);

/* Block loop of gptq.update_weight / gptaq.update_weight (ref: gptq/core.py:226-265,
 * gptaq/core.py:274-319) on already permuted data:
 *   W      [n, k] float32 in/out scratch (permuted weight, dead columns zeroed)
 *   Q      [n, k] float32 out (dequantised result, still permuted)
 *   U      [k, k] float32 upper factor from lcb_chol_inv_upper
 *   P      [k, k] float32 or NULL (GPTAQ: alpha * triu(dXXT U^T, 1) U)
 *   scales, zeros [n, G] float32 (G = k/group; per-row / per-tensor: G = 1), from lcb_qdq(FIND)
 *   keep   [n, k] uint8 or NULL: 1 where the original weight was non-zero (MASK)
 *   group  >0: grouped branch (group | block); -1 or 0: per-column branch
 */
size_t lcb_gptq_ws_bytes(int64_t n, int64_t k, int block);
int lcb_gptq_update(const lcb_quant_cfg* cfg, float* W, float* Q, const float* U, const float* P, const float* scales,
                    const float* zeros, const uint8_t* keep, int64_t n, int64_t k, int64_t group, int block,
                    void* ws, size_t ws_bytes, void* stream);
/* Entry / exit of update_weight around the block loop, one pass over the weight each:
 * lcb_gptq_gather  (ref: gptq/core.py:164-201): Wp[:, j] = float(W[:, col_perm[j]]), keep = (that != 0), then the
 *   dead columns (dead[col] != 0, from lcb_hessian_dead_fix) are zeroed in Wp.  W: [n, k] `dtype`; col_perm [k]
 *   int64 or NULL (identity: no act-order); dead [k] uint8 or NULL.
 * lcb_gptq_scatter (ref: gptq/core.py:267-278): out[:, col_perm[j]] = cast(Q[:, j]) -- inverse permutation and the
 *   cast back to the weight dtype. */
int lcb_gptq_gather(const void* W, int dtype, const int64_t* col_perm, const uint8_t* dead, float* Wp, uint8_t* keep,
                    int64_t n, int64_t k, void* stream);
int lcb_gptq_scatter(const float* Q, const int64_t* col_perm, void* out, int dtype, int64_t n, int64_t k, void* stream);
/* P = alpha * triu(dXXT @ U^T, 1) @ U   (ref: gptaq/core.py:272); dXXT is left untouched */
size_t lcb_gptaq_p_ws_bytes(int64_t k);
int lcb_gptaq_p(float* P, float* dxxt, const float* U, int64_t k, float alpha, void* ws, size_t ws_bytes, void* stream);

/* Block loop of sparsegpt.prune_weight (ref: sparsegpt/core.py:192-218); W [n,k] fp32 in/out. */
size_t lcb_sparsegpt_ws_bytes(int64_t n, int64_t k, int block);
int lcb_sparsegpt_update(float* W, const float* U, double sparsity, int64_t n, int64_t k, int block, void* ws,
                         size_t ws_bytes, void* stream);

/* Row-sharded variant: this rank holds n_local of the n_total output rows; the per-block threshold
 * (ref: sparsegpt/core.py:202, the int(numel * sparsity)-th smallest saliency of the [n_total, 128] block) is found
 * by the same 4-pass radix select, with `reduce(hist, 256, user, stream)` called once per pass to sum the uint32
 * histograms over the ranks (4 * k/128 calls, each must be ordered on `stream`; e.g. an NCCL all-reduce).  Every rank
 * then applies the identical threshold to its rows: masks are bit-identical to the unsharded call.
 * The callback returns 0 on success. */
typedef int (*lcb_reduce_u32_fn)(void* device_u32, int64_t count, void* user, void* stream);
int lcb_sparsegpt_update_sharded(float* W, const float* U, double sparsity, int64_t n_local, int64_t n_total, int64_t k,
                                 int block, void* ws, size_t ws_bytes, lcb_reduce_u32_fn reduce, void* reduce_user,
                                 void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense contractions of the solvers (lazy-batch update ref: gptq/core.py:265, gptaq/core.py:272,319,
 * sparsegpt/core.py:218; Cholesky trailing updates behind ref: gptq/core.py:213-224).
 * lcb_set_gemm_mode selects how they are computed for every later call in this process and returns
 * the previous mode:
 *   1 (default)  tcgen05 tensor cores with the 3xTF32 split (x = hi + lo, three MMAs, fp32 TMEM
 *                accumulator): ~2^-21 relative error per product instead of fp32's 2^-24
 *   0            exact fp32 FFMA (SIMT) -- bit-reproduces the reference's fp32 results on the
 *                golden cases; used by the parity tests as the anchor
 * lcb_tgemm_nt exposes the tensor-core GEMM itself:  C[m,n] (+)= alpha * A[m,kd] * B[n,kd]^T, all
 * row-major fp32 with leading dimensions in elements; kd % 4 == 0, ldc % 4 == 0, C 16 B aligned.
 * kchain > 0 cuts the reduction into accumulation chains of kchain columns that are combined by fp32
 * reduce-adds at L2 (the tensor core's own accumulator truncates; error grows with the chain length).
 */
int lcb_set_gemm_mode(int mode);
size_t lcb_tgemm_ws_bytes(int64_t m, int64_t n, int64_t kd);
int lcb_tgemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t m, int64_t n,
                 int64_t kd, float alpha, int accumulate, int kchain, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (d) mask selection.  mask: uint8 [n, k], 1 = prune.
 * lcb_mask_wanda : ref: wanda/core.py:116-126 (per row, k_prune = int(k*ratio) smallest by
 *                  (|W|*sqrt(s), column index)).
 * lcb_mask_magnitude : ref: magnitude/core.py:38-43 (global threshold, <=).
 * lcb_mask_ria   : ref: ria/core.py:118-126.
 * W: [n, k] dtype `dtype`.  scaler_row: [k] float32.
 */
size_t lcb_mask_ws_bytes(int64_t n, int64_t k);
int lcb_mask_wanda(const void* W, int dtype, const float* scaler_row, uint8_t* mask, int64_t n, int64_t k,
                   double ratio, void* ws, size_t ws_bytes, void* stream);
int lcb_mask_magnitude(const void* W, int dtype, uint8_t* mask, int64_t n, int64_t k, double ratio, void* ws,
                       size_t ws_bytes, void* stream);
int lcb_mask_ria(const void* W, int dtype, const float* scaler_row, uint8_t* mask, int64_t n, int64_t k, double ratio,
                 float alpha, void* ws, size_t ws_bytes, void* stream);
/* Phase API behind lcb_mask_magnitude / lcb_mask_ria for row-sharded weights.  Exact global k-th
 * smallest of scores spread over ranks, MSB-first radix select, 4 passes:
 *     lcb_select_init(state, kth_global)
 *     for pass in 0..3:  lcb_select_hist(local scores, n_local, state, pass)
 *                        all-reduce(SUM) of the 256 uint32 counters at byte offset LCB_SELECT_HIST_OFFSET of state
 *                        lcb_select_scan(state, pass, thresh)         -- same result on every rank
 *     lcb_mask_le(local scores, thresh, mask)                         -- mask = score <= thresh (ref: `<=`)
 * state: lcb_select_state_bytes() bytes of device memory.
 * RIA (ref: ria/core.py:118-126): lcb_ria_sums gives the UNROUNDED fp32 column sums of |W| over the local rows
 * (all-reduce(SUM) them) and the local row sums rounded to W's dtype; lcb_ria_metric rounds the total column sums to
 * W's dtype and builds the fp32 scores. */
#define LCB_SELECT_HIST_OFFSET 16
size_t lcb_select_state_bytes(void);
int lcb_select_init(void* state, int64_t kth, void* stream);
int lcb_select_hist(const float* vals, int64_t n, void* state, int pass, void* stream);
int lcb_select_scan(void* state, int pass, float* thresh, void* stream);
int lcb_metric_magnitude(const void* W, int dtype, float* metric, int64_t numel, void* stream);
int lcb_ria_sums(const void* W, int dtype, float* colsum_partial, float* rowsum, int64_t n, int64_t k, void* stream);
int lcb_ria_metric(const void* W, int dtype, const float* colsum, const float* rowsum, const float* scaler_row,
                   float* metric, int64_t n, int64_t k, float alpha, void* stream);
int lcb_mask_le(const float* metric, const float* thresh, uint8_t* mask, int64_t numel, void* stream);
/* W[mask] = 0 in place */
int lcb_apply_mask(void* W, int dtype, const uint8_t* mask, int64_t numel, void* stream);

/* ------------------------------------------------------------------------------------------
 * Randomised Hadamard rotation (ref: spinquant/hadamard_utils.py:88-111 matmul_hadU,
 * rotation_utils.py:40-45 random_hadamard_matrix, :57-113 rotate_*, hadamard_utils.py:135-172
 * apply_exact_had_to_linear).  For every row of x [rows, n] (contiguous):
 *     y = T(signs .* x) / divisor,   T = (H_K (x) I_L)(I_K (x) H_L),  n = K * L,  L = 2^m
 * i.e. x @ (diag(signs) * matmul_hadU(I)) computed as a fast Walsh-Hadamard transform instead of the reference's
 * dense fp64 GEMM.  hadk_bits: HOST array; K <= 64: K words, bit a of word a' set <=> H_K[a'][a] == -1 (K == 1: null);
 * 64 < K <= 172 (the had108 .. had172 blocks): 3 words per row, bit (a & 63) of word 3 a' + (a >> 6);
 * signs: device [n] of +-1 or null; divisor: float32(sqrt(n)) for the reference's normalisation;
 * acc64 != 0 accumulates in fp64 (reference precision), else fp32.  x / y dtypes: LCB_F32 / LCB_BF16 / LCB_F64;
 * in place (x == y) is allowed.  Column transforms (R^T @ W) are row transforms of the transpose. */
int lcb_hadamard_rows(const void* x, int dtype_in, void* y, int dtype_out, int64_t rows, int64_t n, const float* signs,
                      const uint64_t* hadk_bits, int K, double divisor, int acc64, void* stream);

/* ------------------------------------------------------------------------------------------
 * Packed export.  The reference stores only fake-quantised bf16 weights (ref: models/llama.py:210-230);
 * its integer / fp4 codes are the intermediate `q` of fake_quantize (int_quant.py:210-212, utils.py:263-272), which
 * lcb_qdq returns as one uint8 per element.  lcb_pack4 packs 4-bit codes two per byte (element 2i in the low nibble),
 * lcb_unpack4 restores the uint8 codes (is_signed: sign-extend INT4 two's complement, else zero-extend fp4). */
int lcb_pack4(const uint8_t* codes, uint8_t* packed, int64_t numel, void* stream);
int lcb_unpack4(const uint8_t* packed, uint8_t* codes, int64_t numel, int is_signed, void* stream);

/* ---- fused activation fake-quant + GEMM of the calibration forwards (ref: modules/qlinear.py:86-88
 * `F.linear(input_quantizer(x), W, b)`):  y[m, n] = QDQ(x)[m, k] @ w[n, k]^T (+ bias), all bf16, fp32 accumulation on
 * tcgen05.  cfg != NULL: INT4 / INT8 activation quantiser (symmetric or asymmetric) whose scales / zeros [m, k / group]
 * (bf16, from lcb_qdq in FIND mode; group = k for per-token, else a multiple of 64 dividing k) are applied to each
 * activation tile in shared memory in the GEMM's operand prologue -- the quantised activation never goes to HBM and is
 * bit-identical to lcb_qdq's output.  cfg == NULL: plain bf16 GEMM.  Results equal F.linear's up to the fp32
 * accumulation order (outputs differ by at most one bf16 ulp). */
int lcb_qlinear_fwd(const lcb_quant_cfg* cfg, const void* x, const void* w, const void* bias, void* y, int64_t m, int64_t n,
                    int64_t k, int64_t group, const void* scales, const void* zeros, void* stream);

/* ---- numerical profile of a fake-quant op on the device (ref: quantizers/base.py:30-113 record_stats, which copies
 * both tensors to the CPU and sorts one).  lcb_profile_stats writes 8 floats to `stats`: min / max of x, min / max of
 * QDQ(x), and the sum over elements of ((x - min x) / (max x - min x) - (q - min q) / (max q - min q))^2 -- the
 * ingredients of the reference's Max, QDQ(Max), ClipError and SQNR columns.  The PC99% column is an exact order
 * statistic: lcb_profile_to_f32 + the lcb_select_* phases. */
size_t lcb_profile_ws_bytes(void);
int lcb_profile_stats(const void* x, const void* q, int dtype, int64_t n, float* stats, void* ws, size_t ws_bytes,
                      void* stream);
int lcb_profile_to_f32(const void* x, int dtype, float* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LCB200_H_ */

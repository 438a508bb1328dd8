"""Thin torch-tensor front end of the C ABI (include/lcb200.h): pointers, sizes, workspace and the
current CUDA stream are taken from torch; all arithmetic happens in liblcb200.so.

Nothing here falls back to PyTorch math: a missing library or a non-CUDA tensor raises.
"""
import ctypes

import torch

from . import _lib
from .quantizers import DeferredStatus, _ptr, _stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.LcbError("liblcb200 needs CUDA tensors (no CPU fallback); got device %s" % t.device)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _x2d(x):
    """hook input -> [tokens, k] bf16 contiguous (ref: gptq/core.py:104-111 reshape + t())"""
    if x.dim() == 2:
        x = x.unsqueeze(0)
    batch = x.shape[0]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.bfloat16:
        # the tcgen05 Hessian kernel multiplies bf16 operands exactly; fp16 / fp32 activations would need a hi + lo split
        # (four Gram products) that is not built -- fail loudly instead of rounding the activations to bf16
        raise NotImplementedError("calibration activations must be bfloat16 (got %s): load the model in bf16 "
                                  "(drivers check this before calibration starts)" % x2.dtype)
    return x2.contiguous(), batch


def hessian_add(H, x2d, alpha, beta, dxxt=None, xfp2d=None, upper_only=False):
    """H = beta*H + alpha*X^T X  (and dXXT likewise with (X_fp - X)^T X).  x2d: [tokens, k] bf16.
    upper_only (beta must be 1): only the tiles touching the upper triangle of H are accumulated;
    call hessian_finalize(H, scale, symmetric=True) once at the end."""
    _need_cuda(H, x2d, dxxt, xfp2d)
    L = _lib.lib()
    tokens, k = x2d.shape
    assert H.dtype == torch.float32 and H.is_contiguous() and tuple(H.shape) == (k, k)
    ws = None
    ws_bytes = 0
    if dxxt is not None:
        assert dxxt.dtype == torch.float32 and dxxt.is_contiguous() and tuple(dxxt.shape) == (k, k)
        ws_bytes = L.lcb_hessian_ws_bytes(tokens, k)
        ws = _ws(ws_bytes, H.device)
    with torch.cuda.device(H.device):
        rc = L.lcb_hessian_accum(_ptr(H), _ptr(dxxt), _ptr(x2d), _ptr(xfp2d), tokens, k, float(alpha), float(beta),
                                 1 if upper_only else 0, _ptr(ws), ws_bytes, _stream(H.device))
    _lib.check(rc, "lcb_hessian_accum")


def hessian_finalize(H, scale, symmetric):
    """H *= scale, mirroring the upper triangle into the lower one when `symmetric`."""
    _need_cuda(H)
    with torch.cuda.device(H.device):
        rc = _lib.lib().lcb_hessian_finalize(_ptr(H), H.shape[0], float(scale), 1 if symmetric else 0, _stream(H.device))
    _lib.check(rc, "lcb_hessian_finalize")
    return H


def hessian_pack_upper(H):
    """The upper 32 x 32 blocks of the raw sums H [k, k], end to end in one contiguous fp32 buffer (layout: lcb200.h) --
    what token-sharded ranks all-reduce instead of the whole matrix (half the bytes)."""
    _need_cuda(H)
    assert H.dtype == torch.float32 and H.is_contiguous() and H.dim() == 2 and H.shape[0] == H.shape[1]
    L = _lib.lib()
    packed = torch.empty(L.lcb_hessian_packed_floats(H.shape[0]), dtype=torch.float32, device=H.device)
    with torch.cuda.device(H.device):
        rc = L.lcb_hessian_pack_upper(_ptr(H), H.shape[0], _ptr(packed), _stream(H.device))
    _lib.check(rc, "lcb_hessian_pack_upper")
    return packed


def hessian_finalize_packed(packed, H, scale):
    """H[i][j] = H[j][i] = scale * packed(min(i,j), max(i,j)): hessian_finalize(H, scale, True) reading packed sums."""
    _need_cuda(H, packed)
    L = _lib.lib()
    assert packed.dtype == torch.float32 and packed.is_contiguous() and packed.numel() == L.lcb_hessian_packed_floats(H.shape[0])
    with torch.cuda.device(H.device):
        rc = L.lcb_hessian_finalize_packed(_ptr(packed), _ptr(H), H.shape[0], float(scale), _stream(H.device))
    _lib.check(rc, "lcb_hessian_finalize_packed")
    return H


def hessian_accum_raw(H, x, nsamples, dxxt=None, x_fp=None):
    """Lazy form of the hook: raw sums H += X^T X (upper-triangle tiles only), dXXT += (X_fp-X)^T X.
    Equal to the reference's running mean after hessian_finalize(H, 2/n, True)."""
    x2d, batch = _x2d(x)
    xfp2d = _x2d(x_fp)[0] if x_fp is not None else None
    hessian_add(H, x2d, 1.0, 1.0, dxxt, xfp2d, upper_only=True)
    return nsamples + batch


MAX_DEFER = 8  # MAX_SAMPLES of csrc/hessian.cu


class HessianAccumulator:
    """Raw-sum accumulation of H += X^T X over hook calls, `defer` calls per kernel launch (lcb_hessian_accum_multi).
    Why: every launch streams the computed half of H through L2 (K = 8192: 140 MB, i.e. through HBM) and pays the
    pipeline fill / accumulator drain of the tensor-core kernel; with `defer` hook inputs per launch each tile runs ONE
    accumulation chain over all of them.  No copy is made: the accumulator keeps a REFERENCE to each deferred input
    until it is flushed, so the caller must not modify those tensors in place before `flush()` (hook inputs are
    intermediate activations that nothing writes to afterwards; use defer=1 if in doubt).
    Measured on B200 (Llama-3.2-3B shapes, 128 x 2048 tokens, lock-step tile schedule of csrc/hessian.cu): Hessian stage
    0.72 s at defer=1, 0.60 s at 2, 0.56 s at 4, 0.52 s at 8 (executed tensor rate 1.02 -> 1.40 PFLOP/s).  The price is the
    chain length: the tensor core's fp32 accumulator truncates, relF(H) grows from ~5e-6 (2048 tokens) to ~2e-5 (8192) and
    ~5e-5 (16384) at K = 8192; GPTQ's layer-output SQNR moves by < 0.001 dB (tests/test_solvers_gpu.py), default 4.
    `flush()` (called by solvers.finalize_hessian) launches whatever is pending and returns the sample count."""

    def __init__(self, H, defer=4):
        self.H, self.defer, self.n = H, max(1, min(int(defer), MAX_DEFER)), 0
        self._pending = []
        self._versions = []

    def add(self, x):
        x2d, batch = _x2d(x)
        self.n += batch
        if self._pending and self._pending[0].shape != x2d.shape:
            self._launch()
        self._pending.append(x2d)
        self._versions.append(x2d._version)
        if len(self._pending) >= self.defer:
            self._launch()
        return self.n

    def _launch(self):
        xs = self._pending
        if not xs:
            return
        for t, v in zip(xs, self._versions):
            if t._version != v:   # deferred inputs are held by reference: an in-place write since add() would corrupt H silently
                raise RuntimeError("HessianAccumulator: a deferred hook input was modified in place before the launch "
                                   "(use HESSIAN_DEFER = 1 for models that overwrite activations in place)")
        self._pending = []
        self._versions = []
        H = self.H
        _need_cuda(H, *xs)
        tokens, k = xs[0].shape
        ptrs = (ctypes.c_void_p * len(xs))(*[t.data_ptr() for t in xs])
        with torch.cuda.device(H.device):
            rc = _lib.lib().lcb_hessian_accum_multi(_ptr(H), ctypes.cast(ptrs, ctypes.c_void_p), len(xs), tokens, k, 1.0, 1,
                                                    _stream(H.device))
        _lib.check(rc, "lcb_hessian_accum_multi")
        for t in xs:  # the launch is asynchronous: keep the inputs alive on this stream until it has run
            t.record_stream(torch.cuda.current_stream(H.device))

    def flush(self):
        self._launch()
        return self.n


def hessian_accum(H, x, nsamples, dxxt=None, x_fp=None):
    """One forward-hook update with the reference's running-mean semantics
    (ref: gptq/core.py:113-119): H *= n/(n+b); n += b; H += (2/n) X^T X.  Returns the new n."""
    x2d, batch = _x2d(x)
    xfp2d = _x2d(x_fp)[0] if x_fp is not None else None
    n_new = nsamples + batch
    hessian_add(H, x2d, 2.0 / n_new, nsamples / n_new, dxxt, xfp2d)
    return n_new


def rownorm_accum(s, x, nsamples):
    """ref: wanda/core.py:102-105: s *= n/(n+b); n += b; s += ||x_k||^2 / n."""
    x2d, batch = _x2d(x)
    _need_cuda(s, x2d)
    tokens, k = x2d.shape
    n_new = nsamples + batch
    with torch.cuda.device(s.device):
        rc = _lib.lib().lcb_rownorm_accum(_ptr(s), _ptr(x2d), tokens, k, 1.0 / n_new, nsamples / n_new, _stream(s.device))
    _lib.check(rc, "lcb_rownorm_accum")
    return n_new


def dead_fix(H):
    """dead = diag(H) == 0; H[dead, dead] = 1 (ref: gptq/core.py:175-176). Returns dead (bool [k])."""
    _need_cuda(H)
    k = H.shape[0]
    dead = torch.empty(k, dtype=torch.uint8, device=H.device)
    with torch.cuda.device(H.device):
        rc = _lib.lib().lcb_hessian_dead_fix(_ptr(H), k, _ptr(dead), _stream(H.device))
    _lib.check(rc, "lcb_hessian_dead_fix")
    return dead.bool()


class PendingFactor:
    """Outcome of a deferred lcb_chol_inv_upper: `resolve()` returns False when the factorisation succeeded and True
    when it had to be redone with the reference's second, larger damping (U was overwritten: dependents must be
    recomputed); raises like torch.linalg.cholesky when that fails too (ref: gptq/core.py:213-221)."""

    def __init__(self, st, src, U, perm, percdamp, ws, ws_bytes):
        self.st, self.src, self.U, self.perm, self.percdamp, self.ws, self.ws_bytes = st, src, U, perm, percdamp, ws, ws_bytes
        self._done = None

    def resolve(self):
        if self._done is None:
            self._done = False
            if self.st.value() & _lib.ST_NOT_SPD:
                _chol_call(self.src, self.U, self.perm, self.percdamp * (11.0 + 10.0 * self.percdamp), self.ws, self.ws_bytes,
                           check=True)
                self._done = True
            self.src = self.ws = None
        return self._done


def _chol_call(src, U, perm, damp, ws, ws_bytes, check):
    st = DeferredStatus(src.device)
    with torch.cuda.device(src.device):
        rc = _lib.lib().lcb_chol_inv_upper(_ptr(src), _ptr(U), src.shape[0], _ptr(perm), float(damp), _ptr(ws), ws_bytes,
                                           _ptr(st.word), _stream(src.device))
    _lib.check(rc, "lcb_chol_inv_upper")
    st.arm()
    if check and (st.value() & _lib.ST_NOT_SPD):
        raise RuntimeError("linalg.cholesky: the Hessian is not positive-definite even after 10x damping")
    return st


def chol_inv_upper(H, perm=None, percdamp=0.01, out=None, defer=False):
    """U with (H[perm][:, perm] + damp*mean(diag)*I)^-1 = U^T U, including the reference's retry
    with 10x damping on a failed factorisation (ref: gptq/core.py:207-224).  H is not modified
    unless `out is H`.
    defer=True returns (U, PendingFactor): the status word is read back without draining the stream -- enqueue the
    work that depends on U, then call `.resolve()`; it reports whether the retry had to run (rare)."""
    _need_cuda(H, perm)
    L = _lib.lib()
    k = H.shape[0]
    assert H.dtype == torch.float32 and H.is_contiguous()
    if perm is not None:
        perm = perm.to(device=H.device, dtype=torch.int64).contiguous()
    U = out if out is not None else torch.empty_like(H)
    ws_bytes = L.lcb_chol_ws_bytes(k)
    ws = _ws(ws_bytes, H.device)
    src = H
    if U.data_ptr() == H.data_ptr():
        src = H.clone()  # keep the input for a possible retry
    st = _chol_call(src, U, perm, percdamp, ws, ws_bytes, check=False)
    pending = PendingFactor(st, src, U, perm, percdamp, ws, ws_bytes)
    if defer:
        return U, pending
    pending.resolve()
    return U


def gptq_block_update(cfg, W, U, scales, zeros, keep, group, P=None, block=128):
    """Block loop of update_weight on permuted data (ref: gptq/core.py:226-265). Returns Q."""
    _need_cuda(W, U, scales, zeros, keep, P)
    L = _lib.lib()
    n, k = W.shape
    for t in (W, U, scales, zeros):
        assert t.dtype == torch.float32 and t.is_contiguous()
    Q = torch.empty_like(W)
    ws_bytes = L.lcb_gptq_ws_bytes(n, k, block)
    ws = _ws(ws_bytes, W.device)
    with torch.cuda.device(W.device):
        rc = L.lcb_gptq_update(ctypes.byref(cfg), _ptr(W), _ptr(Q), _ptr(U), _ptr(P), _ptr(scales), _ptr(zeros),
                               _ptr(keep), n, k, int(group), block, _ptr(ws), ws_bytes, _stream(W.device))
    _lib.check(rc, "lcb_gptq_update")
    return Q


def gptq_gather(W, col_perm=None, dead=None):
    """W [n, k] (bf16 / fp32 weight) -> (Wp fp32 [n, k] with columns permuted and dead columns zeroed, keep uint8):
    `W.float()`, `MASK = W != 0`, `W[:, dead] = 0` and the act-order gather of gptq/core.py:164-201 in one pass."""
    _need_cuda(W, col_perm, dead)
    assert W.dim() == 2 and W.is_contiguous()
    n, k = W.shape
    Wp = torch.empty((n, k), dtype=torch.float32, device=W.device)
    keep = torch.empty((n, k), dtype=torch.uint8, device=W.device)
    if col_perm is not None:
        col_perm = col_perm.to(torch.int64).contiguous()
    if dead is not None:
        dead = dead.to(torch.uint8).contiguous() if dead.dtype != torch.bool else dead.contiguous().view(torch.uint8)
    with torch.cuda.device(W.device):
        rc = _lib.lib().lcb_gptq_gather(_ptr(W), _wdt(W), _ptr(col_perm), _ptr(dead), _ptr(Wp), _ptr(keep), n, k,
                                         _stream(W.device))
    _lib.check(rc, "lcb_gptq_gather")
    return Wp, keep


def gptq_scatter(Q, col_perm, dtype):
    """out[:, col_perm[j]] = Q[:, j] cast to `dtype` (inverse act-order permutation + cast, gptq/core.py:267-278)."""
    _need_cuda(Q, col_perm)
    assert Q.dtype == torch.float32 and Q.is_contiguous()
    n, k = Q.shape
    out = torch.empty((n, k), dtype=dtype, device=Q.device)
    if col_perm is not None:
        col_perm = col_perm.to(torch.int64).contiguous()
    with torch.cuda.device(Q.device):
        rc = _lib.lib().lcb_gptq_scatter(_ptr(Q), _ptr(col_perm), _ptr(out), _wdt(out), n, k, _stream(Q.device))
    _lib.check(rc, "lcb_gptq_scatter")
    return out


def gptaq_p(dxxt, U, alpha):
    """P = alpha * triu(dXXT @ U^T, 1) @ U (ref: gptaq/core.py:272); dxxt is left untouched."""
    _need_cuda(dxxt, U)
    L = _lib.lib()
    k = U.shape[0]
    P = torch.empty_like(U)
    ws_bytes = L.lcb_gptaq_p_ws_bytes(k)
    ws = _ws(ws_bytes, U.device)
    with torch.cuda.device(U.device):
        rc = L.lcb_gptaq_p(_ptr(P), _ptr(dxxt), _ptr(U), k, float(alpha), _ptr(ws), ws_bytes, _stream(U.device))
    _lib.check(rc, "lcb_gptaq_p")
    return P


def sparsegpt_update(W, U, sparsity, block=128, n_total=None, reduce=None):
    """Block loop of prune_weight (ref: sparsegpt/core.py:192-218); W [n,k] fp32 updated in place.
    Row-sharded form: W holds this rank's n of the n_total output rows and `reduce(hist)` sums an
    int32[256] device tensor over the ranks in place (called 4 times per 128-column block, on the current stream);
    the per-block threshold is then the exact global one and the masks equal the unsharded call's bit for bit."""
    _need_cuda(W, U)
    L = _lib.lib()
    n, k = W.shape
    assert W.dtype == torch.float32 and W.is_contiguous() and U.is_contiguous()
    ws_bytes = L.lcb_sparsegpt_ws_bytes(n, k, block)
    ws = _ws(ws_bytes, W.device)
    with torch.cuda.device(W.device):
        if reduce is None:
            assert n_total is None or n_total == n
            rc = L.lcb_sparsegpt_update(_ptr(W), _ptr(U), float(sparsity), n, k, block, _ptr(ws), ws_bytes,
                                        _stream(W.device))
        else:
            base, failure = ws.data_ptr(), []

            def _cb(ptr, count, _user, _stream_):
                try:
                    off = int(ptr) - base
                    assert 0 <= off and off % 4 == 0 and off + 4 * count <= ws_bytes
                    reduce(ws[off:off + 4 * count].view(torch.int32))
                    return 0
                except BaseException as e:  # never let an exception cross the C frame
                    failure.append(e)
                    return 1

            cb = _lib.REDUCE_U32_FN(_cb)
            rc = L.lcb_sparsegpt_update_sharded(_ptr(W), _ptr(U), float(sparsity), n, int(n_total), k, block, _ptr(ws),
                                                ws_bytes, cb, None, _stream(W.device))
            if failure:
                raise failure[0]
    _lib.check(rc, "lcb_sparsegpt_update")
    return W


def _wdt(W):
    if W.dtype == torch.bfloat16:
        return _lib.BF16
    if W.dtype == torch.float32:
        return _lib.F32
    raise NotImplementedError("weights must be bfloat16 or float32, got %s" % W.dtype)


def _mask_call(fn_name, W, scaler_row, ratio, alpha=None):
    _need_cuda(W, scaler_row)
    L = _lib.lib()
    assert W.dim() == 2 and W.is_contiguous()
    n, k = W.shape
    mask = torch.empty((n, k), dtype=torch.uint8, device=W.device)
    ws_bytes = L.lcb_mask_ws_bytes(n, k) if fn_name != "lcb_mask_wanda" else 256
    ws = _ws(ws_bytes, W.device)
    if scaler_row is not None:
        scaler_row = scaler_row.to(torch.float32).contiguous()
    with torch.cuda.device(W.device):
        if fn_name == "lcb_mask_wanda":
            rc = L.lcb_mask_wanda(_ptr(W), _wdt(W), _ptr(scaler_row), _ptr(mask), n, k, float(ratio), _ptr(ws), ws_bytes,
                                  _stream(W.device))
        elif fn_name == "lcb_mask_magnitude":
            rc = L.lcb_mask_magnitude(_ptr(W), _wdt(W), _ptr(mask), n, k, float(ratio), _ptr(ws), ws_bytes,
                                      _stream(W.device))
        else:
            rc = L.lcb_mask_ria(_ptr(W), _wdt(W), _ptr(scaler_row), _ptr(mask), n, k, float(ratio), float(alpha),
                                _ptr(ws), ws_bytes, _stream(W.device))
    _lib.check(rc, fn_name)
    return mask.bool()


def mask_wanda(W, scaler_row, ratio):
    return _mask_call("lcb_mask_wanda", W, scaler_row, ratio)


def mask_magnitude(W, ratio):
    return _mask_call("lcb_mask_magnitude", W, None, ratio)


def mask_ria(W, scaler_row, ratio, alpha):
    return _mask_call("lcb_mask_ria", W, scaler_row, ratio, alpha)


# ---- phase API for row-sharded global thresholds (parallel.py drives the collectives in between)
SELECT_HIST_OFFSET = 16  # LCB_SELECT_HIST_OFFSET


def select_state(device):
    """Device state of one distributed radix select: (int32 buffer, view of its 256 histogram counters)."""
    nbytes = int(_lib.lib().lcb_select_state_bytes())
    buf = torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=device)
    return buf, buf[SELECT_HIST_OFFSET // 4: SELECT_HIST_OFFSET // 4 + 256]


def select_init(state, kth):
    _need_cuda(state)
    with torch.cuda.device(state.device):
        _lib.check(_lib.lib().lcb_select_init(_ptr(state), int(kth), _stream(state.device)), "lcb_select_init")


def select_hist(state, scores, pass_):
    _need_cuda(state, scores)
    assert scores.dtype == torch.float32 and scores.is_contiguous()
    with torch.cuda.device(state.device):
        _lib.check(_lib.lib().lcb_select_hist(_ptr(scores) if scores.numel() else None, scores.numel(), _ptr(state),
                                              int(pass_), _stream(state.device)), "lcb_select_hist")


def select_scan(state, pass_, thresh):
    _need_cuda(state, thresh)
    with torch.cuda.device(state.device):
        _lib.check(_lib.lib().lcb_select_scan(_ptr(state), int(pass_), _ptr(thresh), _stream(state.device)),
                   "lcb_select_scan")


def metric_magnitude(W):
    """|W| as fp32 scores (ref: magnitude/core.py:38-39)."""
    _need_cuda(W)
    assert W.is_contiguous()
    m = torch.empty(W.shape, dtype=torch.float32, device=W.device)
    with torch.cuda.device(W.device):
        _lib.check(_lib.lib().lcb_metric_magnitude(_ptr(W) if W.numel() else None, _wdt(W), _ptr(m) if W.numel() else None,
                                                   W.numel(), _stream(W.device)), "lcb_metric_magnitude")
    return m


def ria_sums(W):
    """(unrounded fp32 column sums of |W| over these rows, row sums rounded to W's dtype) -- ref: ria/core.py:118-121."""
    _need_cuda(W)
    assert W.dim() == 2 and W.is_contiguous()
    n, k = W.shape
    cs = torch.zeros(k, dtype=torch.float32, device=W.device)
    rs = torch.empty(n, dtype=torch.float32, device=W.device)
    if n == 0:  # empty row shard: contributes zero column sums
        return cs, rs
    with torch.cuda.device(W.device):
        _lib.check(_lib.lib().lcb_ria_sums(_ptr(W), _wdt(W), _ptr(cs), _ptr(rs), n, k, _stream(W.device)), "lcb_ria_sums")
    return cs, rs


def ria_metric(W, colsum, rowsum, scaler_row, alpha):
    _need_cuda(W, colsum, rowsum, scaler_row)
    n, k = W.shape
    m = torch.empty((n, k), dtype=torch.float32, device=W.device)
    if n == 0:
        return m
    scaler_row = scaler_row.to(torch.float32).contiguous()
    with torch.cuda.device(W.device):
        _lib.check(_lib.lib().lcb_ria_metric(_ptr(W), _wdt(W), _ptr(colsum), _ptr(rowsum), _ptr(scaler_row), _ptr(m), n, k,
                                             float(alpha), _stream(W.device)), "lcb_ria_metric")
    return m


def mask_le(scores, thresh):
    """scores <= thresh[0] (ref: the `<=` of ria/core.py:125, magnitude/core.py:42)."""
    _need_cuda(scores, thresh)
    mask = torch.empty(scores.shape, dtype=torch.uint8, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(_lib.lib().lcb_mask_le(_ptr(scores) if scores.numel() else None, _ptr(thresh),
                                          _ptr(mask) if scores.numel() else None, scores.numel(), _stream(scores.device)),
                   "lcb_mask_le")
    return mask.bool()


def apply_mask(W, mask):
    """W[mask] = 0 in place."""
    _need_cuda(W, mask)
    assert W.is_contiguous()
    m = mask.to(torch.uint8).contiguous()
    with torch.cuda.device(W.device):
        rc = _lib.lib().lcb_apply_mask(_ptr(W), _wdt(W), _ptr(m), W.numel(), _stream(W.device))
    _lib.check(rc, "lcb_apply_mask")
    return W


def set_gemm_mode(mode):
    """Solver GEMM precision for every later call: 1 = tcgen05 3xTF32 (default), 0 = exact fp32 FFMA.
    Returns the previous mode."""
    return int(_lib.lib().lcb_set_gemm_mode(int(mode)))


def tgemm_nt(A, B, C=None, alpha=1.0, accumulate=False, kchain=0):
    """C (+)= alpha * A @ B^T on the tcgen05 3xTF32 path (fp32 row-major operands)."""
    _need_cuda(A, B, C)
    L = _lib.lib()
    m, kd = A.shape
    n = B.shape[0]
    assert B.shape[1] == kd and A.dtype == torch.float32 and B.dtype == torch.float32
    assert A.stride(1) == 1 and B.stride(1) == 1
    if C is None:
        assert not accumulate
        C = torch.empty((m, n), dtype=torch.float32, device=A.device)
    assert C.stride(1) == 1 and C.dtype == torch.float32
    ws_bytes = L.lcb_tgemm_ws_bytes(m, n, kd)
    ws = _ws(ws_bytes, A.device)
    with torch.cuda.device(A.device):
        rc = L.lcb_tgemm_nt(_ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(C), C.stride(0), m, n, kd, float(alpha),
                            1 if accumulate else 0, int(kchain), _ptr(ws), ws_bytes, _stream(A.device))
    _lib.check(rc, "lcb_tgemm_nt")
    return C


def qlinear_forward(x, weight, bias=None, quantizer=None):
    """F.linear(quantizer(x), weight, bias) with the activation fake-quant fused into the operand prologue of a tcgen05
    GEMM (lcb_qlinear_fwd; ref: modules/qlinear.py:86-88).  `quantizer`: an INTQuantizer along the last axis, per token
    (group_size -1) or per group (a multiple of 64 dividing k), or None for the plain GEMM.  bf16 CUDA tensors."""
    _need_cuda(x, weight, bias)
    L = _lib.lib()
    k = x.shape[-1]
    n = weight.shape[0]
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.shape[1] == k
    x2 = x.reshape(-1, k).contiguous()
    m = x2.shape[0]
    w = weight.contiguous()
    b = bias.contiguous() if bias is not None else None
    y = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
    cfg = scales = zeros = None
    group = k
    if quantizer is not None:
        scales, zeros = quantizer.find_params(x2)            # find-only pass: [m, G, 1] in x.dtype
        group = k if quantizer.group_size in (-1, k) else int(quantizer.group_size)
        scales = scales.reshape(m, k // group).contiguous()
        zeros = zeros.reshape(m, k // group).contiguous()
        cfg = ctypes.byref(quantizer._cfg())
    with torch.cuda.device(x.device):
        rc = L.lcb_qlinear_fwd(cfg, _ptr(x2), _ptr(w), _ptr(b), _ptr(y), m, n, k, group, _ptr(scales), _ptr(zeros),
                               _stream(x.device))
    _lib.check(rc, "lcb_qlinear_fwd")
    return y.reshape(x.shape[:-1] + (n,))


def qlinear_fusable(x, weight, quantizer):
    """Can `F.linear(quantizer(x), weight)` run as one lcb_qlinear_fwd?"""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16):
        return False
    k = x.shape[-1]
    if type(quantizer).__name__ != "INTQuantizer" or quantizer.axes != -1 or getattr(quantizer, "mse", False) or quantizer.is_profile:
        return False
    gs = quantizer.group_size
    return k % 64 == 0 and (gs in (-1, k) or (isinstance(gs, int) and gs > 0 and gs % 64 == 0 and k % gs == 0))

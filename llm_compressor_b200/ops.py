"""Thin torch-tensor front end of the C ABI (include/lcb200.h): pointers, sizes, workspace and the
current CUDA stream are taken from torch; all arithmetic happens in liblcb200.so.

Nothing here falls back to PyTorch math: a missing library or a non-CUDA tensor raises.
"""
import ctypes

import torch

from . import _lib
from .quantizers import _ptr, _status_word, _stream


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
        raise NotImplementedError("calibration activations must be bfloat16 (got %s)" % x2.dtype)
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


def hessian_accum_raw(H, x, nsamples, dxxt=None, x_fp=None):
    """Lazy form of the hook: raw sums H += X^T X (upper-triangle tiles only), dXXT += (X_fp-X)^T X.
    Equal to the reference's running mean after hessian_finalize(H, 2/n, True)."""
    x2d, batch = _x2d(x)
    xfp2d = _x2d(x_fp)[0] if x_fp is not None else None
    hessian_add(H, x2d, 1.0, 1.0, dxxt, xfp2d, upper_only=True)
    return nsamples + batch


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


def chol_inv_upper(H, perm=None, percdamp=0.01, out=None):
    """U with (H[perm][:, perm] + damp*mean(diag)*I)^-1 = U^T U, including the reference's retry
    with 10x damping on a failed factorisation (ref: gptq/core.py:207-224).  H is not modified
    unless `out is H`."""
    _need_cuda(H, perm)
    L = _lib.lib()
    k = H.shape[0]
    assert H.dtype == torch.float32 and H.is_contiguous()
    if perm is not None:
        perm = perm.to(device=H.device, dtype=torch.int64).contiguous()
    U = out if out is not None else torch.empty_like(H)
    ws_bytes = L.lcb_chol_ws_bytes(k)
    ws = _ws(ws_bytes, H.device)
    status = _status_word(H.device)
    src = H
    if U.data_ptr() == H.data_ptr():
        src = H.clone()  # keep the input for a possible retry
    for damp in (percdamp, percdamp * (11.0 + 10.0 * percdamp)):
        # second value == the reference's "damp again by 10x on the already damped H"
        with torch.cuda.device(H.device):
            rc = L.lcb_chol_inv_upper(_ptr(src), _ptr(U), k, _ptr(perm), float(damp), _ptr(ws), ws_bytes,
                                         _ptr(status), _stream(H.device))
        _lib.check(rc, "lcb_chol_inv_upper")
        st = int(status.item())
        if st:
            status.zero_()
        if not (st & _lib.ST_NOT_SPD):
            return U
    raise RuntimeError("linalg.cholesky: the Hessian is not positive-definite even after 10x damping")


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


def sparsegpt_update(W, U, sparsity, block=128):
    """Block loop of prune_weight (ref: sparsegpt/core.py:192-218); W [n,k] fp32 updated in place."""
    _need_cuda(W, U)
    L = _lib.lib()
    n, k = W.shape
    assert W.dtype == torch.float32 and W.is_contiguous() and U.is_contiguous()
    ws_bytes = L.lcb_sparsegpt_ws_bytes(n, k, block)
    ws = _ws(ws_bytes, W.device)
    with torch.cuda.device(W.device):
        rc = L.lcb_sparsegpt_update(_ptr(W), _ptr(U), float(sparsity), n, k, block, _ptr(ws), ws_bytes,
                                    _stream(W.device))
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

"""TEST INFRASTRUCTURE ONLY -- Python front end of the CPU oracle.

Restates the compression hot path of pjh5672/llm-compressor on the CPU:
  * elementwise / reduction arithmetic (fake-quant, masks, row norms) in C (lc_oracle.c, exact
    dtype emulation, bit-exact contract);
  * the dense linear algebra of the layer solvers (Hessian, Cholesky chain, GPTQ / GPTAQ /
    SparseGPT block loops) in numpy float32 + LAPACK, statement by statement after
    quantization/calibrations/gptq/core.py:103-119,163-281,
    quantization/calibrations/gptaq/core.py:116-141,198-335 and
    pruning/sparsegpt/core.py:85-101,160-228 of the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.

Parity status: pinned against outputs of the imported reference (tests/golden/, produced by
oracle/gen_golden.py in the build container); the reference itself has no golden vectors.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblc_oracle.so")

F32, BF16 = 0, 1
QTYPES = {"int": 0, "fp": 1, "mx": 2, "nvfp": 3}
ELEMS = {"int4": 1, "int8": 2, "fp4_e2m1": 3, "fp8_e4m3": 4, "fp8_e5m2": 5}

_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "lc_oracle.c")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        u8 = ctypes.POINTER(ctypes.c_uint8)
        L.orc_qdq_rows.restype = ctypes.c_int
        L.orc_qdq_rows.argtypes = [fp, fp, fp, fp, fp, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int]
        L.orc_elem_core.restype = None
        L.orc_elem_core.argtypes = [fp, fp, ctypes.c_long, ctypes.c_int, ctypes.c_int]
        L.orc_mask_wanda.restype = None
        L.orc_mask_wanda.argtypes = [fp, fp, u8, ctypes.c_long, ctypes.c_long, ctypes.c_double]
        L.orc_mask_threshold.restype = ctypes.c_float
        L.orc_mask_threshold.argtypes = [fp, u8, ctypes.c_long, ctypes.c_double]
        L.orc_mask_ria.restype = ctypes.c_float
        L.orc_mask_ria.argtypes = [fp, fp, u8, ctypes.c_long, ctypes.c_long, ctypes.c_double,
                                   ctypes.c_float, ctypes.c_int]
        L.orc_rownorm_accum.restype = None
        L.orc_rownorm_accum.argtypes = [fp, fp, ctypes.c_long, ctypes.c_long, ctypes.c_long]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def _u8(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def bf16_round(a):
    """Round-to-nearest-even float32 -> bf16-representable float32 (numpy)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) & 0xFFFF0000).astype(np.uint32)
    special = (u & 0x7F800000) == 0x7F800000
    r = np.where(special, (u & 0xFFFF0000).astype(np.uint32), r)
    return r.view(np.float32).reshape(a.shape)


def elem_core(a, elem, dt):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty_like(a)
    lib().orc_elem_core(_fp(a), _fp(out), a.size, ELEMS[elem], dt)
    return out


def qdq(x, cfg, dt, scales=None, zeros=None, want_codes=False, only_params=False, mse=False):
    """Fake-quantize like `quantizer(x)` / `quantizer.find_params(x)` of the reference.

    x: float32 ndarray (bf16-representable when dt == BF16) of rank >= 2 (rank >= 1 per-tensor).
    cfg: reference config dict (type, format, group_size, axes, zero_point[, scale_ebits]).
    Returns (out, scales, zeros, codes); scales/zeros are block-shaped like the reference's
    ([..., G, 1] for axes=-1, [..., G, 1, C] for axes=-2, 0-dim for per-tensor).
    mse=True: find_params with the clip search (quantizer.mse = True; INT / FP / MX, axes=-1).
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    qt, el = QTYPES[cfg["type"]], ELEMS[cfg["format"]]
    zp = 1 if cfg.get("zero_point", False) else 0
    sebits = int(cfg.get("scale_ebits", 8))
    gs, axes = cfg["group_size"], cfg.get("axes", -1)
    if cfg["type"] in ("int", "fp"):
        if gs == -1:
            axes = -1
        elif gs == -2:
            axes = -2
    given = scales is not None and zeros is not None

    if gs == 0:  # per-tensor (int_quant.py:103-112): one group over everything, no padding
        assert cfg["type"] in ("int", "fp")
        flat = x.reshape(1, -1)
        group, view, back = flat.shape[1], flat, lambda a: a.reshape(x.shape)
        pshape = ()
    else:
        if axes == -1:
            group = x.shape[-1] if gs == -1 else gs
            view = x.reshape(-1, x.shape[-1])
            back = lambda a: a.reshape(x.shape)
            G = -(-x.shape[-1] // group)
            pshape = x.shape[:-1] + (G, 1)
        else:
            group = x.shape[-2] if gs == -2 else gs
            xt = np.ascontiguousarray(np.swapaxes(x, -1, -2))  # [..., C, R]
            view = xt.reshape(-1, xt.shape[-1])
            back = lambda a: np.ascontiguousarray(np.swapaxes(a.reshape(xt.shape), -1, -2))
            G = -(-x.shape[-2] // group)
            pshape = x.shape[:-2] + (G, 1, x.shape[-1])
    rows, cols = view.shape
    G = -(-cols // group)
    view = np.ascontiguousarray(view)
    s = np.empty((rows, G), np.float32)
    z = np.empty((rows, G), np.float32)
    if given:
        sb = np.asarray(scales, np.float32)
        zb = np.asarray(zeros, np.float32)
        if gs == 0:
            s[:] = sb.reshape(()); z[:] = zb.reshape(())
        elif axes == -1:
            s[:] = np.broadcast_to(sb.reshape(sb.shape[:-1]), x.shape[:-1] + (G,)).reshape(rows, G)
            z[:] = np.broadcast_to(zb.reshape(zb.shape[:-1]), x.shape[:-1] + (G,)).reshape(rows, G)
        else:
            # [..., G, 1, C] -> [..., C, G]
            sb = np.swapaxes(sb.reshape(sb.shape[:-2] + (sb.shape[-1],)), -1, -2)
            zb = np.swapaxes(zb.reshape(zb.shape[:-2] + (zb.shape[-1],)), -1, -2)
            s[:] = np.broadcast_to(sb, x.shape[:-2] + (x.shape[-1], G)).reshape(rows, G)
            z[:] = np.broadcast_to(zb, x.shape[:-2] + (x.shape[-1], G)).reshape(rows, G)
    out = None if only_params else np.empty_like(view)
    codes = np.empty_like(view) if (want_codes and not only_params) else None
    rc = lib().orc_qdq_rows(_fp(view), _fp(out), _fp(s), _fp(z), _fp(codes), rows, cols, group, qt, el, zp,
                            sebits, dt, F32 if gs == 0 else dt, 1 if given else 0, float("nan"), 1 if mse else 0)
    assert rc == 0, rc
    if gs == 0:
        s_out, z_out = s.reshape(()), z.reshape(())
    elif axes == -1:
        s_out, z_out = s.reshape(pshape), z.reshape(pshape)
    else:
        lead = x.shape[:-2]
        s_out = np.swapaxes(s.reshape(lead + (x.shape[-1], G)), -1, -2).reshape(pshape)
        z_out = np.swapaxes(z.reshape(lead + (x.shape[-1], G)), -1, -2).reshape(pshape)
    if gs != 0:
        s_out, z_out = np.ascontiguousarray(s_out), np.ascontiguousarray(z_out)
    return (None if out is None else back(out), s_out, z_out,
            None if codes is None else back(codes))


# ----------------------------------------------------------------------------- masks
def mask_wanda(w, scaler_row, ratio):
    w = np.ascontiguousarray(w, np.float32)
    m = np.zeros(w.shape, np.uint8)
    lib().orc_mask_wanda(_fp(w), _fp(np.ascontiguousarray(scaler_row, np.float32)), _u8(m), w.shape[0],
                         w.shape[1], float(ratio))
    return m.astype(bool)


def mask_threshold(metric, ratio):
    metric = np.ascontiguousarray(metric, np.float32)
    m = np.zeros(metric.shape, np.uint8)
    th = lib().orc_mask_threshold(_fp(metric), _u8(m), metric.size, float(ratio))
    return m.astype(bool), th


def mask_magnitude(w, ratio):
    return mask_threshold(np.abs(np.asarray(w, np.float32)), ratio)


def mask_ria(w, scaler_row, ratio, alpha, dt):
    w = np.ascontiguousarray(w, np.float32)
    m = np.zeros(w.shape, np.uint8)
    th = lib().orc_mask_ria(_fp(w), _fp(np.ascontiguousarray(scaler_row, np.float32)), _u8(m), w.shape[0],
                            w.shape[1], float(ratio), float(alpha), dt)
    return m.astype(bool), th


def rownorm_accum(s, x, nsamples):
    """wanda/core.py:92-105 for one sample x [T, K]; returns nsamples + 1 (s updated in place)."""
    x = np.ascontiguousarray(x, np.float32)
    assert s.dtype == np.float32 and s.flags.c_contiguous
    lib().orc_rownorm_accum(_fp(s), _fp(x), x.shape[0], x.shape[1], int(nsamples))
    return nsamples + 1


# ----------------------------------------------------------------------------- Hessian
def hessian_accum(H, x, nsamples, dXXT=None, x_fp=None):
    """gptq/core.py:103-119 (and gptaq/core.py:116-141 when dXXT/x_fp are given), batch of 1.

    x, x_fp: [T, K] float32 holding bf16 values. H, dXXT: [K, K] float32, updated in place.
    """
    tmp = 1
    H *= np.float32(nsamples / (nsamples + tmp))
    if dXXT is not None:
        dXXT *= np.float32(nsamples / (nsamples + tmp))
    nsamples += tmp
    c = np.float32(math.sqrt(2 / nsamples))
    inp = (c * x.astype(np.float32)).T  # [K, T]
    H += inp @ inp.T
    if dXXT is not None:
        dX = c * x_fp.astype(np.float32).T - inp
        dXXT += dX @ inp.T
    return nsamples


def _chol_upper_of_inverse(H):
    """gptq/core.py:213-224: cholesky -> cholesky_inverse -> cholesky(upper), fp32 LAPACK."""
    from scipy.linalg import lapack

    L, info = lapack.spotrf(H, lower=1, clean=1)
    if info != 0:
        raise np.linalg.LinAlgError("potrf info=%d" % info)
    Hi, info = lapack.spotri(L, lower=1)
    if info != 0:
        raise np.linalg.LinAlgError("potri info=%d" % info)
    Hi = np.tril(Hi) + np.tril(Hi, -1).T
    U, info = lapack.spotrf(np.ascontiguousarray(Hi, np.float32), lower=0, clean=1)
    if info != 0:
        raise np.linalg.LinAlgError("potrf(upper) info=%d" % info)
    return np.ascontiguousarray(U, np.float32)


def damp_and_factor(H, percdamp):
    """_adjust_dump + retry (gptq/core.py:207-224). H is modified in place like the reference."""
    K = H.shape[0]
    idx = np.arange(K)
    try:
        H[idx, idx] += np.float32(percdamp) * np.mean(np.diag(H), dtype=np.float32)
        return _chol_upper_of_inverse(H.copy())
    except np.linalg.LinAlgError:
        H[idx, idx] += np.float32(percdamp * 10) * np.mean(np.diag(H), dtype=np.float32)
        return _chol_upper_of_inverse(H.copy())


def _wq(w, cfg, s, z):
    """layer.weight_quantizer(w, scales=s, zeros=z) on fp32 data; s, z are [N] vectors."""
    N = w.shape[0]
    c = dict(cfg)
    c["group_size"], c["axes"] = -1, -1  # one group per row of the slice passed in
    if cfg["type"] in ("mx", "nvfp"):
        c["group_size"] = w.shape[1]
    out, _, _, _ = qdq(w, c, F32, scales=s.reshape(N, 1, 1), zeros=z.reshape(N, 1, 1))
    return out


def gptq_update(W, H, cfg, block_size=128, percdamp=0.01, actorder=True, dXXT=None, alpha=0.25,
                out_dtype_bf16=True):
    """gptq/core.py:163-281 (dXXT None) or gptaq/core.py:198-335 (dXXT given).

    W: [N, K] float32 (bf16-representable values when the layer is bf16). H (and dXXT): [K, K]
    float32, consumed.  cfg: weight quantizer config.  Returns the new weight as float32
    (rounded to bf16 when out_dtype_bf16).
    """
    W = np.array(W, np.float32, copy=True)
    N, K = W.shape
    MASK = W != 0
    gs = cfg["group_size"]
    dead = np.diag(H) == 0
    H[dead, dead] = 1
    W[:, dead] = 0
    if dXXT is not None:
        dXXT[:, dead] = 0
    per_col = gs in (0, -1)
    _, scales, zeros, _ = qdq(W, cfg, F32, only_params=True)
    if actorder:
        if per_col:
            perm = np.argsort(-np.diag(H), kind="stable")
            W = W[:, perm]; MASK = MASK[:, perm]; H = H[perm][:, perm]
            if dXXT is not None:
                dXXT = dXXT[perm][:, perm]
        else:
            ng = K // gs
            perm = np.argsort(-np.diag(H).reshape(-1, gs).sum(-1, dtype=np.float32), kind="stable")
            W = W.reshape(N, ng, gs)[:, perm, :].reshape(N, K)
            MASK = MASK.reshape(N, ng, gs)[:, perm, :].reshape(N, K)
            _, scales, zeros, _ = qdq(W, cfg, F32, only_params=True)
            H = H.reshape(ng, gs, ng, gs)[perm][:, :, perm, :].reshape(K, K)
            if dXXT is not None:
                dXXT = dXXT.reshape(ng, gs, ng, gs)[perm][:, :, perm, :].reshape(K, K)
        invperm = np.argsort(perm, kind="stable")
    H = np.ascontiguousarray(H, np.float32)
    Q = np.zeros_like(W)
    Hinv = damp_and_factor(H, percdamp)
    P = None
    if dXXT is not None:
        P = (np.float32(alpha) * np.triu(np.ascontiguousarray(dXXT, np.float32) @ Hinv.T, 1)) @ Hinv
    if per_col:
        s_row = np.broadcast_to(scales.reshape(-1), (N,)) if gs == 0 else scales.reshape(N)
        z_row = np.broadcast_to(zeros.reshape(-1), (N,)) if gs == 0 else zeros.reshape(N)
        s_row, z_row = np.ascontiguousarray(s_row), np.ascontiguousarray(z_row)
    else:
        s2, z2 = scales.reshape(N, K // gs), zeros.reshape(N, K // gs)
    for i1 in range(0, K, block_size):
        i2 = min(i1 + block_size, K)
        count = i2 - i1
        W1 = W[:, i1:i2].copy()
        M1 = MASK[:, i1:i2]
        Q1 = np.zeros_like(W1)
        E1 = np.zeros_like(W1)
        U1 = Hinv[i1:i2, i1:i2]
        P1 = P[i1:i2, i1:i2] if P is not None else None
        if per_col:
            for i in range(count):
                w = W1[:, i].copy()
                q = _wq(w.reshape(N, 1), cfg, s_row, z_row).reshape(N) * M1[:, i]
                Q1[:, i] = q
                e = (w - q) / U1[i, i]
                upd = np.outer(e, U1[i, i:])
                if P1 is not None:
                    upd = upd - np.outer(w, P1[i, i:])
                W1[:, i:] -= upd
                E1[:, i] = e
        else:
            for i in range(0, count, gs):
                g = (i1 + i) // gs
                w = W1[:, i:i + gs].copy()
                q = _wq(w, cfg, np.ascontiguousarray(s2[:, g]), np.ascontiguousarray(z2[:, g])) * M1[:, i:i + gs]
                Q1[:, i:i + gs] = q
                e = (w - q) / np.diag(U1)[i:i + gs]
                upd = e @ U1[i:i + gs, i:]
                if P1 is not None:
                    upd = upd - w @ P1[i:i + gs, i:]
                W1[:, i:] -= upd
                E1[:, i:i + gs] = e
        Q[:, i1:i2] = Q1
        if i2 < K:
            upd = E1 @ Hinv[i1:i2, i2:]
            if P is not None:
                upd = upd - W1 @ P[i1:i2, i2:]
            W[:, i2:] -= upd
    if actorder:
        if per_col:
            Q = Q[:, invperm]
        else:
            Q = Q.reshape(N, K // gs, gs)[:, invperm, :].reshape(N, K)
    Q = np.ascontiguousarray(Q, np.float32)
    return bf16_round(Q) if out_dtype_bf16 else Q


def sparsegpt_prune(W, H, sparsity, block_size=128, percdamp=0.01, out_dtype_bf16=True):
    """pruning/sparsegpt/core.py:160-228."""
    W = np.array(W, np.float32, copy=True)
    N, K = W.shape
    dead = np.diag(H) == 0
    H[dead, dead] = 1
    W[:, dead] = 0
    Hinv = damp_and_factor(np.ascontiguousarray(H, np.float32), percdamp)
    for i1 in range(0, K, block_size):
        i2 = min(i1 + block_size, K)
        count = i2 - i1
        W1 = W[:, i1:i2].copy()
        Q1 = np.zeros_like(W1)
        E1 = np.zeros_like(W1)
        U1 = Hinv[i1:i2, i1:i2]
        tmp = W1 ** 2 / (np.diag(U1).reshape(1, -1)) ** 2
        M1, _ = mask_threshold(tmp, sparsity)
        for i in range(count):
            w = W1[:, i].copy()
            q = w.copy()
            q[M1[:, i]] = 0
            Q1[:, i] = q
            e = (w - q) / U1[i, i]
            W1[:, i:] -= np.outer(e, U1[i, i:])
            E1[:, i] = e
        W[:, i1:i2] = Q1
        if i2 < K:
            W[:, i2:] -= E1 @ Hinv[i1:i2, i2:]
    return bf16_round(W) if out_dtype_bf16 else W


# ----------------------------------------------------------------------------- Hadamard rotation
def _hadk_table(K):
    """The reference's fixed K x K Hadamard block (hadamard_utils.py get_had<K>), read from the golden fixture that
    oracle/gen_golden_hadamard.py extracted from the reference itself."""
    z = np.load(os.path.join(os.path.dirname(_HERE), "tests", "golden", "hadamard.npz"))
    neg = np.unpackbits(z["had%d" % K])[: K * K].reshape(K, K).astype(bool)
    return np.where(neg, -1.0, 1.0)


def get_hadk(n):
    """ref: hadamard_utils.py:17-85 (same precedence; only the sizes present in the fixture)."""
    for K in (172, 156, 140, 108, 60, 52, 36, 28, 44, 40, 20, 12):
        if n % K == 0:
            return _hadk_table(K), K
    return None, 1


def matmul_hadU(X, transpose=False, signs=None):
    """ref: hadamard_utils.py:88-111: radix-2 butterflies on the power-of-two part (adjacent pairs first), then the
    K x K block across the leading index, then division by float32(sqrt(n)).  fp64 throughout.
    signs: optional [n] of +-1 multiplied onto X first (= X @ diag(signs) @ matmul_hadU(I), rotation_utils.py:40-45)."""
    X = np.asarray(X, np.float64)
    n = X.shape[-1]
    hadK, K = get_hadk(n)
    if hadK is not None and transpose:
        hadK = hadK.T
    v = X.reshape(-1, n).copy()
    if signs is not None:
        v = v * np.asarray(signs, np.float64)[None, :]
    B = v.shape[0]
    cur = v.reshape(B, n, 1)
    while cur.shape[1] > K:
        cur = cur.reshape(B, cur.shape[1] // 2, 2, cur.shape[2])
        nxt = np.empty_like(cur)
        nxt[:, :, 0, :] = cur[:, :, 0, :] + cur[:, :, 1, :]
        nxt[:, :, 1, :] = cur[:, :, 0, :] - cur[:, :, 1, :]
        cur = nxt.reshape(B, cur.shape[1], -1)
    if K > 1:
        cur = hadK.reshape(1, K, K) @ cur
    return cur.reshape(X.shape) / np.float64(np.sqrt(np.float32(n)))

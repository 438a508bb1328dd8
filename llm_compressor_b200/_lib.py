"""ctypes binding of liblcb200.so (the C ABI declared in include/lcb200.h).

The library is built in-tree by `build()` (nvcc, sm_100a only).  There is NO fallback: if the
shared object is missing or a call fails, an exception is raised.
"""
import ctypes
import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "liblcb200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]

F32, BF16 = 0, 1
Q_INT, Q_FP, Q_MX, Q_NVFP = 0, 1, 2, 3
ELEM = {"int4": 1, "int8": 2, "fp4_e2m1": 3, "fp8_e4m3": 4, "fp8_e5m2": 5}
QDQ_FIND, QDQ_APPLY = 1, 2
ST_NAN_SCALE, ST_NOT_SPD = 1, 2


class LcbError(RuntimeError):
    pass


class QuantCfg(ctypes.Structure):
    _fields_ = [
        ("qtype", ctypes.c_int32),
        ("elem", ctypes.c_int32),
        ("zero_point", ctypes.c_int32),
        ("scale_ebits", ctypes.c_int32),
        ("mse", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 3),
    ]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source under csrc/ for sm_100a and link liblcb200.so in-tree."""
    srcs = sorted(glob.glob(os.path.join(_CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(_CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(_PKG), "include", "lcb200.h")]
    objdir = os.path.join(_PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """Load liblcb200.so.  Raises LcbError when it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LcbError(
            "liblcb200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "this package has no CPU or PyTorch fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
    cfgp = ctypes.POINTER(QuantCfg)
    L.lcb_abi_version.restype = i32
    L.lcb_last_error.restype = ctypes.c_char_p
    L.lcb_qdq_ws_bytes.restype = sz
    L.lcb_qdq_ws_bytes.argtypes = [cfgp, i32, i64, i64, i64, i32, i64]
    L.lcb_qdq.restype = i32
    L.lcb_qdq.argtypes = [cfgp, i32, i32, vp, vp, i64, i64, i64, i32, i64, vp, vp, vp, vp, vp, sz, vp, vp]
    L.lcb_nvfp_global_amax.restype = i32
    L.lcb_nvfp_global_amax.argtypes = [cfgp, i32, vp, i64, i64, i64, i32, i64, vp, vp]
    for name, args, res in _OPTIONAL:
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
    if L.lcb_abi_version() != 1:
        raise LcbError("liblcb200.so ABI version mismatch")
    _lib = L
    return L


_vp, _i64, _i32, _sz, _f32, _f64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t, ctypes.c_float, ctypes.c_double
_cfgp = ctypes.POINTER(QuantCfg)
REDUCE_U32_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p)  # lcb_reduce_u32_fn
_OPTIONAL = [
    ("lcb_launch_count", [], ctypes.c_uint64),
    ("lcb_hessian_ws_bytes", [_i64, _i64], _sz),
    ("lcb_hessian_accum", [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _i32, _vp, _sz, _vp], _i32),
    ("lcb_hessian_finalize", [_vp, _i64, _f32, _i32, _vp], _i32),
    ("lcb_hessian_packed_floats", [_i64], _sz),
    ("lcb_hessian_pack_upper", [_vp, _i64, _vp, _vp], _i32),
    ("lcb_hessian_finalize_packed", [_vp, _vp, _i64, _f32, _vp], _i32),
    ("lcb_hessian_accum_multi", [_vp, _vp, _i32, _i64, _i64, _f32, _i32, _vp], _i32),
    ("lcb_rownorm_accum", [_vp, _vp, _i64, _i64, _f32, _f32, _vp], _i32),
    ("lcb_chol_ws_bytes", [_i64], _sz),
    ("lcb_chol_trace_offset", [_i64], _sz),
    ("lcb_qlinear_fwd", [_cfgp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp], _i32),
    ("lcb_profile_ws_bytes", [], _sz),
    ("lcb_profile_stats", [_vp, _vp, _i32, _i64, _vp, _vp, _sz, _vp], _i32),
    ("lcb_profile_to_f32", [_vp, _i32, _vp, _i64, _vp], _i32),
    ("lcb_hessian_dead_fix", [_vp, _i64, _vp, _vp], _i32),
    ("lcb_chol_inv_upper", [_vp, _vp, _i64, _vp, _f32, _vp, _sz, _vp, _vp], _i32),
    ("lcb_gptq_ws_bytes", [_i64, _i64, _i32], _sz),
    ("lcb_gptq_update", [_cfgp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _sz, _vp], _i32),
    ("lcb_gptq_gather", [_vp, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _vp], _i32),
    ("lcb_gptq_scatter", [_vp, _vp, _vp, _i32, _i64, _i64, _vp], _i32),
    ("lcb_gptaq_p_ws_bytes", [_i64], _sz),
    ("lcb_gptaq_p", [_vp, _vp, _vp, _i64, _f32, _vp, _sz, _vp], _i32),
    ("lcb_sparsegpt_ws_bytes", [_i64, _i64, _i32], _sz),
    ("lcb_sparsegpt_update", [_vp, _vp, _f64, _i64, _i64, _i32, _vp, _sz, _vp], _i32),
    ("lcb_mask_ws_bytes", [_i64, _i64], _sz),
    ("lcb_mask_wanda", [_vp, _i32, _vp, _vp, _i64, _i64, _f64, _vp, _sz, _vp], _i32),
    ("lcb_mask_magnitude", [_vp, _i32, _vp, _i64, _i64, _f64, _vp, _sz, _vp], _i32),
    ("lcb_mask_ria", [_vp, _i32, _vp, _vp, _i64, _i64, _f64, _f32, _vp, _sz, _vp], _i32),
    ("lcb_apply_mask", [_vp, _i32, _vp, _i64, _vp], _i32),
    ("lcb_sparsegpt_update_sharded", [_vp, _vp, _f64, _i64, _i64, _i64, _i32, _vp, _sz, _vp, _vp, _vp], _i32),
    ("lcb_select_state_bytes", [], _sz),
    ("lcb_select_init", [_vp, _i64, _vp], _i32),
    ("lcb_select_hist", [_vp, _i64, _vp, _i32, _vp], _i32),
    ("lcb_select_scan", [_vp, _i32, _vp, _vp], _i32),
    ("lcb_metric_magnitude", [_vp, _i32, _vp, _i64, _vp], _i32),
    ("lcb_ria_sums", [_vp, _i32, _vp, _vp, _i64, _i64, _vp], _i32),
    ("lcb_ria_metric", [_vp, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _vp], _i32),
    ("lcb_mask_le", [_vp, _vp, _vp, _i64, _vp], _i32),
    ("lcb_pack4", [_vp, _vp, _i64, _vp], _i32),
    ("lcb_unpack4", [_vp, _vp, _i64, _i32, _vp], _i32),
    ("lcb_hadamard_rows", [_vp, _i32, _vp, _i32, _i64, _i64, _vp, _vp, _i32, _f64, _i32, _vp], _i32),
    ("lcb_set_gemm_mode", [_i32], _i32),
    ("lcb_tgemm_ws_bytes", [_i64, _i64, _i64], _sz),
    ("lcb_tgemm_nt", [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _f32, _i32, _i32, _vp, _sz, _vp], _i32),
]

# every symbol include/lcb200.h declares (tests check that the library exports all of them)
DECLARED_SYMBOLS = ["lcb_abi_version", "lcb_last_error", "lcb_qdq_ws_bytes", "lcb_qdq", "lcb_nvfp_global_amax"] + [
    n for n, _, _ in _OPTIONAL]


def check(rc, what):
    if rc != 0:
        msg = lib().lcb_last_error()
        raise LcbError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))


def make_cfg(qtype, elem, zero_point, scale_ebits=8, mse=False):
    c = QuantCfg()
    c.qtype, c.elem, c.zero_point, c.scale_ebits = int(qtype), int(elem), int(bool(zero_point)), int(scale_ebits)
    c.mse = int(bool(mse))
    return c

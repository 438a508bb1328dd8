"""CPU: the C-ABI library loads and exports every symbol include/lcb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "lcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lcb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    fns = declared_functions()
    assert "lcb_qdq" in fns and "lcb_hessian_accum" in fns and len(fns) >= 15


def test_library_exports_every_declared_symbol():
    from llm_compressor_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [f for f in declared_functions() if not hasattr(L, f)]
    assert not missing, "not exported: %s" % missing
    assert sorted(_lib.DECLARED_SYMBOLS) == declared_functions()


def test_loads_and_reports_version_without_gpu():
    from llm_compressor_b200 import _lib
    assert _lib.lib().lcb_abi_version() == 1


def test_invalid_arguments_are_rejected_without_gpu():
    from llm_compressor_b200 import _lib
    L = _lib.lib()
    cfg = _lib.make_cfg(_lib.Q_NVFP, _lib.ELEM["int8"], False)
    rc = L.lcb_qdq(ctypes.byref(cfg), _lib.BF16, 3, None, None, 1, 4, 128, -1, 16, None, None, None, None, None, 0, None, None)
    assert rc == -1 and b"NVFP" in L.lcb_last_error()


def test_cpu_tensor_is_refused_loudly():
    import torch
    import llm_compressor_b200 as lc
    q = lc.FakeQuantizer.build(dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False))
    with pytest.raises(lc._lib.LcbError):
        q(torch.zeros(4, 128, dtype=torch.bfloat16))

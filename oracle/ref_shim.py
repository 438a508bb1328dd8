"""TEST INFRASTRUCTURE ONLY. Import shim for the reference (pjh5672/llm-compressor at /root/reference).

Imports the reference from /root/reference (build container) or from the unmodified copy in oracle/_ref/
(GPU box; see oracle/make_ref.py).  Used by oracle/gen_golden*.py to produce tests/golden/*.npz, by the
cross-checks of the oracle restatement, by bench.py's reference legs (`--impl reference`, `reference_eager_b200`:
the reference timed as the baseline, never as the product) and by tests/test_reference_dropin_gpu.py.  Nothing in
the product package imports this file.

The stubs replace import-time-only dependencies of the reference that are absent here
(ref: llm_compressor/utils/general.py:10 matplotlib; utils/parser.py:4 easydict;
quantization/calibrations/spinquant/hadamard_utils.py:3 fast_hadamard_transform).
"""
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


_ZIP = os.path.join(_HERE, "_ref", "llm_compressor_ref.zip")


def _default_root():
    """/root/reference in the build container; on the GPU box the archive of unmodified reference modules that
    oracle/make_ref.py left in oracle/_ref/ (git-ignored, travels with the snapshot; imported in place via zipimport)."""
    for cand in (os.environ.get("LC_REFERENCE_ROOT"), "/root/reference"):
        if cand and (os.path.isdir(os.path.join(cand, "llm_compressor")) or (cand.endswith(".zip") and os.path.isfile(cand))):
            return cand
    return _ZIP


REF_ROOT = _default_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "llm_compressor")) or (REF_ROOT.endswith(".zip") and os.path.isfile(REF_ROOT))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


def install():
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            _stub("matplotlib")
            sys.modules["matplotlib"].pyplot = _stub("matplotlib.pyplot")
    if "easydict" not in sys.modules:
        try:
            import easydict  # noqa: F401
        except Exception:
            _stub("easydict", EasyDict=EasyDict)
    if "fast_hadamard_transform" not in sys.modules:
        _stub("fast_hadamard_transform", hadamard_transform=None)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def build_quantizer(cfg):
    install()
    from llm_compressor.quantization.quant import FakeQuantizer

    return FakeQuantizer.build(dict(cfg))

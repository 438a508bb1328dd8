"""B200-native (sm_100a) implementation of the data-parallel compression hot path of
pjh5672/llm-compressor: calibration Hessians, GPTQ / GPTAQ / SparseGPT layer solvers, fused
fake-quant for INT / FP / MX / NVFP formats and Wanda / RIA / magnitude mask selection.

The numerics live in hand-written CUDA kernels behind the C ABI of include/lcb200.h
(liblcb200.so, built in-tree by `_lib.build()`); this package is the Python host side that
mirrors the reference's quantizer / solver interfaces.  No CPU or PyTorch fallback exists.
"""
from . import _lib  # noqa: F401
from .quantizers import (  # noqa: F401
    DummyQuantizer, ElemFormat, FakeQuantizer, FPQuantizer, INTQuantizer, MXQuantizer, NVFPQuantizer,
)

__version__ = "0.1.0"

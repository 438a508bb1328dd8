"""Numerical profile of a fake-quant op (`--profile`): the reference's `record_stats` on the GPU.

ref: llm_compressor/quantization/quantizers/base.py:30-113.  The reference copies x and QDQ(x) to the CPU, sorts x for the
99th percentile and reduces both for max / SQNR / clipping error.  Here the tensors stay where they are: two flat
reduction passes (lcb_profile_stats), an exact k-th order statistic by radix select for PC99% (lcb_profile_to_f32 +
lcb_select_*, no sort), ONE 32-byte read-back, and the same line appended to `<save_path>/stats.csv` in the reference's
format (`%46s,` + `%14.5g,` columns).  Arithmetic is fp32 per element (the reference computes in the tensor dtype on
the CPU, i.e. bf16 for bf16 activations: its SQNR column carries that rounding, ours does not)."""
import math
import os

import torch

from . import _lib
from .quantizers import _ptr, _stream

KEYS = ("Op Name", "PC99%", "Max", "QDQ(Max)", "SQNR", "ClipError", "Elem", "BPV")


def bits_per_value(qtype, fmt, group_size, zero_point, numel):
    """BPV column (ref: base.py:66-92); qtype in DUMMY / INT / FP / MX / NV, fmt the upper-case format name."""
    if qtype == "DUMMY":
        return 16
    zeros = 16 / group_size if zero_point else 0
    if qtype == "NV":
        return 4 + (16 / numel + 8 / group_size) + zeros
    four = fmt in ("INT4", "FP4_E2M1") if qtype == "MX" else fmt == ("INT4" if qtype == "INT" else "FP4_E2M1")
    return (4 if four else 8) + 16 / group_size + zeros


def device_stats(x, qdq_x):
    """(pc99, max x, max q, sqnr, clip error) of a CUDA tensor pair, computed on the device."""
    from . import parallel
    if not (x.is_cuda and qdq_x.is_cuda):
        raise _lib.LcbError("liblcb200 needs CUDA tensors (no CPU fallback)")
    x, q = x.detach().contiguous(), qdq_x.detach().contiguous()
    assert x.dtype == q.dtype and x.numel() == q.numel()
    L = _lib.lib()
    dt = _lib.BF16 if x.dtype == torch.bfloat16 else _lib.F32
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise NotImplementedError("profile: bfloat16 / float32 tensors")
    n = x.numel()
    stats = torch.empty(8, dtype=torch.float32, device=x.device)
    ws = torch.empty(int(L.lcb_profile_ws_bytes()), dtype=torch.uint8, device=x.device)
    xf = torch.empty(n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.lcb_profile_stats(_ptr(x), _ptr(q), dt, n, _ptr(stats), _ptr(ws), ws.numel(), _stream(x.device)),
                   "lcb_profile_stats")
        _lib.check(L.lcb_profile_to_f32(_ptr(x), dt, _ptr(xf), n, _stream(x.device)), "lcb_profile_to_f32")
    kth = round(0.99 * (n - 1))                                  # ref: extract_percentile (base.py:57-59)
    pc99 = parallel.select_kth_sharded(xf, kth, group=None)      # single process: exact k-th smallest, no sort
    host = torch.cat([stats[:5], pc99]).cpu()                    # the one read-back
    xmin, xmax, qmin, qmax, sq, p99 = (float(v) for v in host)
    sqnr = -10.0 * math.log10(sq / n + 1e-10)
    return p99, xmax, qmax, sqnr, xmax - qmax


def record_stats(quantizer, x, qdq_x):
    """Drop-in for BaseQuantizer.record_stats (same columns, same file format)."""
    pc99, maxval, qdq_max, sqnr, clip = device_stats(x, qdq_x)
    name = type(quantizer).__name__
    qtype = {"INTQuantizer": "INT", "FPQuantizer": "FP", "MXQuantizer": "MX", "NVFPQuantizer": "NV"}.get(name, "DUMMY")
    gs = quantizer.group_size if isinstance(getattr(quantizer, "group_size", None), int) and quantizer.group_size > 0 else x.shape[-1]
    bpv = bits_per_value(qtype, getattr(quantizer, "str_format", "BF16"), gs, bool(getattr(quantizer, "zero_point", False)), x.numel())
    vals = (quantizer.op_name, pc99, maxval, qdq_max, sqnr, clip, x.numel(), bpv)
    path = os.path.join(str(quantizer.save_path), "stats.csv")
    head = "" if os.path.exists(path) else ((("%46s," + "%14s," * (len(KEYS) - 1)) % KEYS).rstrip(",") + "\n")
    with open(path, "a") as f:
        f.write(head + (("%46s," + "%14.5g," * (len(vals) - 1)) % vals).rstrip(",") + "\n")
    return dict(zip(KEYS, vals))

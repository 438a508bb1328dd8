"""Fake-quant modules: host-side mirror of the reference quantizer API on top of lcb_qdq.

Same class names, constructor arguments, attributes and call signatures as
  ref: llm_compressor/quantization/quantizers/int_quant.py:13  INTQuantizer
  ref: llm_compressor/quantization/quantizers/fp_quant.py:21   FPQuantizer
  ref: llm_compressor/quantization/quantizers/mx_quant.py:21   MXQuantizer
  ref: llm_compressor/quantization/quantizers/nvfp_quant.py:21 NVFPQuantizer
  ref: llm_compressor/quantization/quantizers/dummy.py:9       DummyQuantizer
  ref: llm_compressor/quantization/quant.py:36                 FakeQuantizer.build
so a reference driver (rtn / gptq / awq ...) can use them unchanged.  All arithmetic runs in the
hand-written CUDA kernels of liblcb200.so (csrc/qdq.cu); there is no PyTorch fallback.
"""
import ctypes
from enum import Enum

import torch
import torch.nn as nn

from . import _lib


class ElemFormat(Enum):  # ref: quantizers/formats.py:11-16
    int4 = 1
    int8 = 2
    fp4_e2m1 = 3
    fp8_e4m3 = 4
    fp8_e5m2 = 5

    @staticmethod
    def from_str(s):
        assert s is not None, "String elem_format == None"
        s = s.lower()
        if hasattr(ElemFormat, s):
            return getattr(ElemFormat, s)
        raise Exception("Undefined elem format", s)


# ref: quantizers/formats.py:41-92 -> (ebits, mbits, emax, max_norm, min_norm)
_FORMAT_PARAMS = {
    ElemFormat.int4: (0, 4, 0, 1.75, 0),
    ElemFormat.int8: (0, 8, 0, 127.0 / 64.0, 0),
    ElemFormat.fp4_e2m1: (2, 3, 2, 6.0, 1.0),
    ElemFormat.fp8_e4m3: (4, 5, 8, 448.0, 2.0 ** -6),
    ElemFormat.fp8_e5m2: (5, 4, 15, 57344.0, 2.0 ** -14),
}


def _get_format_params(fmt):
    if isinstance(fmt, str):
        fmt = ElemFormat.from_str(fmt)
    return _FORMAT_PARAMS[fmt]


def _as_format(fmt):
    if isinstance(fmt, ElemFormat):
        return fmt
    if isinstance(fmt, str):
        return ElemFormat.from_str(fmt)
    # the reference's own enum (same member names) is accepted too
    return ElemFormat.from_str(str(fmt).split(".")[-1])


_status = {}


def _status_word(device):
    key = (device.type, device.index)
    t = _status.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=device)
        _status[key] = t
    return t


class DeferredStatus:
    """A device status word (LCB_ST_* bits) whose read-back does not drain the stream.

    The reference reads its error conditions synchronously (`assert not torch.isnan(scales).any()` int_quant.py:165;
    `torch.linalg.cholesky` raising, gptq/core.py:213-221).  Here the producing call gets a private word; `arm()`
    copies it to a pooled pinned host word on the current stream and records an event; `value()` waits for THAT event
    only, so work enqueued after `arm()` keeps the GPU busy while the host looks at the result."""
    _pool = {}

    def __init__(self, device):
        self.device = torch.device(device)
        free = DeferredStatus._pool.setdefault((self.device.type, self.device.index), [])
        if free:
            self.word, self.host, self.event = free.pop()
        else:
            self.word = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self.event = torch.cuda.Event()
        self.word.zero_()
        self._value = None

    def arm(self):
        self.host.copy_(self.word, non_blocking=True)
        self.event.record(torch.cuda.current_stream(self.device))
        return self

    def value(self):
        if self._value is None:
            self.event.synchronize()
            self._value = int(self.host[0])
            DeferredStatus._pool[(self.device.type, self.device.index)].append((self.word, self.host, self.event))
            self.word = self.host = self.event = None
        return self._value


def _dt(x):
    if x.dtype == torch.bfloat16:
        return _lib.BF16
    if x.dtype == torch.float32:
        return _lib.F32
    raise NotImplementedError("liblcb200 fake-quant supports bfloat16 and float32 tensors, got %s" % x.dtype)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _prod(xs):
    p = 1
    for v in xs:
        p *= int(v)
    return p


def qdq_raw(cfg, x, axis, group, find, apply, scales=None, zeros=None, codes=False, nv_amax=None, check_nan=True,
            blocked=False, status=None):
    """One lcb_qdq call on a CUDA tensor.  Returns (out, scales, zeros, codes).

    Geometry (not blocked): axis -1 -> rows = prod(shape[:-1]), cols = shape[-1];
    axis -2 -> batch = prod(shape[:-2]), rows = shape[-2], cols = shape[-1]; group 0 -> per tensor.
    blocked=True: x is already `[..., G, bs]` (axis -1) or `[..., G, bs, C]` (axis -2).
    Scales / zeros come back block-shaped like the reference's (`[..., G, 1]`, `[..., G, 1, C]`,
    0-dim fp32 for per-tensor).
    """
    if not x.is_cuda:
        raise _lib.LcbError("liblcb200 needs CUDA tensors (no CPU fallback); got device %s" % x.device)
    L = _lib.lib()
    x = x.contiguous()
    dt = _dt(x)
    shape = tuple(x.shape)
    if group == 0:
        batch, rows, cols = 1, 1, x.numel()
        pshape, pdtype, n_params = (), torch.float32, 1
    elif axis == -1:
        if x.dim() < 1:
            raise ValueError("need at least 1 dimension")
        batch, rows, cols = 1, _prod(shape[:-1]), shape[-1]
        G = -(-cols // group)
        pshape = shape[:-1] + ((1,) if blocked else (G, 1))
        pdtype, n_params = x.dtype, rows * G
    else:
        if x.dim() < 2:
            raise ValueError("axes=-2 needs at least 2 dimensions")
        batch, rows, cols = _prod(shape[:-2]), shape[-2], shape[-1]
        G = -(-rows // group)
        pshape = shape[:-2] + ((1, cols) if blocked else (G, 1, cols))
        pdtype, n_params = x.dtype, batch * G * cols
    if blocked:
        assert G == 1, "blocked input must hold exactly one group along the reduced axis"
    given = not find
    if given:
        scales = torch.broadcast_to(scales.to(device=x.device), pshape if group else ()).to(pdtype).contiguous()
        zeros = torch.broadcast_to(zeros.to(device=x.device), pshape if group else ()).to(pdtype).contiguous()
    else:
        scales = torch.empty(pshape, dtype=pdtype, device=x.device)
        zeros = torch.empty(pshape, dtype=pdtype, device=x.device)
    out = torch.empty_like(x) if apply else None
    cds = torch.empty(shape, dtype=torch.uint8, device=x.device) if (codes and apply) else None
    ws_bytes = L.lcb_qdq_ws_bytes(ctypes.byref(cfg), dt, batch, rows, cols, axis, group)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    deferred = status       # a DeferredStatus: the caller checks the NaN-scale bit later (no host sync here)
    if deferred is not None:
        status = deferred.word if find else None
    else:
        status = _status_word(x.device) if (find and check_nan) else None
    mode = (_lib.QDQ_FIND if find else 0) | (_lib.QDQ_APPLY if apply else 0)
    with torch.cuda.device(x.device):
        rc = L.lcb_qdq(ctypes.byref(cfg), dt, mode, _ptr(x), _ptr(out), batch, rows, cols, axis, group, _ptr(scales),
                       _ptr(zeros), _ptr(cds), _ptr(nv_amax), _ptr(ws), ws_bytes, _ptr(status), _stream(x.device))
    _lib.check(rc, "lcb_qdq")
    if deferred is not None:
        deferred.arm()
    elif status is not None:
        st = int(status.item())  # same host sync as the reference's assert (int_quant.py:165)
        if st:
            status.zero_()
            assert not (st & _lib.ST_NAN_SCALE), "NaN in quantization scales"
    return out, scales, zeros, cds


class BaseQuantizer(nn.Module):  # ref: quantizers/base.py:8
    _QTYPE = None
    _ALLOWED = ()

    def __init__(self, format, group_size=-1, axes=-1, zero_point=False, is_profile=False, **kwargs):
        super().__init__()
        self.is_profile = is_profile
        op_name = kwargs.get("op_name", None)
        self.op_name = op_name if op_name is not None else "None"
        self.save_path = kwargs.get("save_path", "./")
        self.format = _as_format(format)
        assert self.format in self._ALLOWED, f"Not support Format for {self.__class__.__name__}"
        ebits, mbits, emax, max_norm, min_norm = _get_format_params(self.format)
        self._register_format_buffers(ebits, mbits, emax, max_norm, min_norm)
        self.scale_ebits = kwargs.get("scale_ebits", 8)
        self.str_format = self.format.name.upper()
        self.mse = False
        self.check_nan = True
        self.configure(zero_point=zero_point, group_size=group_size, axes=axes)

    def _register_format_buffers(self, ebits, mbits, emax, max_norm, min_norm):
        self.register_buffer("ebits", torch.tensor(ebits))
        self.register_buffer("mbits", torch.tensor(mbits))
        self.register_buffer("emax", torch.tensor(emax))
        self.register_buffer("max_norm", torch.tensor(max_norm))
        self.register_buffer("min_norm", torch.tensor(min_norm))

    def configure(self, zero_point, group_size, axes):
        self.zero_point = zero_point
        self.group_size = group_size
        self.axes = axes

    def _cfg(self, mse=False):
        return _lib.make_cfg(self._QTYPE, self.format.value, self.zero_point, self.scale_ebits, mse=mse)

    def _resolve_group(self, x):
        """group_size -1 / -2 become concrete on first use, like the reference (int_quant.py:82-85)."""
        if self.group_size == -1:
            self.group_size = x.shape[-1]
        elif self.group_size == -2:
            self.group_size = x.shape[-2]
        if isinstance(self.group_size, (list, tuple)):
            raise NotImplementedError("2-D block quantisation is commented out in the reference (utils.py:86-91)")

    def _mse(self):
        """mse clip search (ref: int_quant.py:115-162): INT / FP / MX / symmetric NVFP along the last axis; the
        kernel refuses the rest (LCB_ERR_UNSUPPORTED)."""
        return bool(self.mse)

    def find_params(self, x, already_reshaped=False, status=None):
        """`status` (extension): a DeferredStatus that receives the NaN-scale bit instead of the synchronous assert."""
        if self.group_size != 0 and not already_reshaped:
            self._resolve_group(x)
        _, s, z, _ = qdq_raw(self._cfg(self._mse()), x, self.axes, self.group_size, True, False, check_nan=self.check_nan,
                             blocked=bool(already_reshaped and self.group_size != 0), status=status)
        return s, z

    def forward(self, x, **kwargs):
        scales = kwargs.pop("scales", None)
        zeros = kwargs.pop("zeros", None)
        if self.group_size != 0:
            self._resolve_group(x)
        if (scales is not None) and (zeros is not None):
            s, z = self._given_params(x, scales, zeros)
            out, _, _, _ = qdq_raw(self._cfg(), x, self.axes, self.group_size, False, True, scales=s, zeros=z)
        else:
            out, _, _, _ = qdq_raw(self._cfg(self._mse()), x, self.axes, self.group_size, True, True,
                                   check_nan=self.check_nan)
        if self.is_profile:
            self.record_stats(x=x, qdq_x=out)
        return out

    def _given_params(self, x, scales, zeros):
        """Caller-provided parameters are block-shaped (`[..., G, 1]`); drop nothing, just make
        them broadcastable against this call's block view."""
        if self.group_size == 0:
            return scales.reshape(()), zeros.reshape(())
        return scales, zeros

    def fake_quantize(self, x, scales, zeros):
        """x is block-shaped (`[..., G, bs]` / `[..., G, bs, C]`), like the reference's."""
        blocked = self.group_size != 0
        out, _, _, _ = qdq_raw(self._cfg(), x, self.axes, x.shape[self.axes] if blocked else 0, False, True,
                               scales=scales, zeros=zeros, blocked=blocked)
        return out

    def quantize_with_codes(self, x):
        """Extension: dequantised tensor, scales, zeros and the uint8 codes."""
        if self.group_size != 0:
            self._resolve_group(x)
        return qdq_raw(self._cfg(), x, self.axes, self.group_size, True, True, codes=True, check_nan=self.check_nan)

    def record_stats(self, x, qdq_x, **kwargs):  # ref: quantizers/base.py:30-113 (numerical profile, host side)
        from .profile import record_stats
        record_stats(self, x, qdq_x)

    def extra_repr(self):
        return f"Format: {self.str_format}, Axes: {self.axes}, Group: {self.group_size}, ZP: {self.zero_point}"


class INTQuantizer(BaseQuantizer):
    _QTYPE = _lib.Q_INT
    _ALLOWED = (ElemFormat.int4, ElemFormat.int8)

    def _register_format_buffers(self, ebits, mbits, emax, max_norm, min_norm):
        self.q_bits = mbits
        q_max = max_norm * 2 ** (mbits - 2)  # restricted symmetric range, also for asymmetric
        self.register_buffer("q_max", torch.tensor(q_max))
        self.register_buffer("q_min", torch.tensor(-q_max))

    def configure(self, zero_point, group_size, axes):
        assert (group_size != 1) or not zero_point, "Asymmetric quant with per-element quant. are exclusive."
        self.zero_point = zero_point
        self.group_size = group_size
        if group_size == -1:
            self.axes = -1
        elif group_size == -2:
            self.axes = -2
        elif isinstance(group_size, (list, tuple)):
            self.axes = -1
        else:
            self.axes = axes


class FPQuantizer(BaseQuantizer):
    _QTYPE = _lib.Q_FP
    _ALLOWED = (ElemFormat.fp4_e2m1, ElemFormat.fp8_e4m3, ElemFormat.fp8_e5m2)
    configure = INTQuantizer.configure


class MXQuantizer(BaseQuantizer):
    _QTYPE = _lib.Q_MX
    _ALLOWED = tuple(ElemFormat)

    def __init__(self, format, group_size=32, axes=-1, zero_point=False, is_profile=False, **kwargs):
        super().__init__(format, group_size, axes, zero_point, is_profile, **kwargs)
        self.scale_emax = 2 ** (self.scale_ebits - 1) - 1

    def _resolve_group(self, x):
        if not isinstance(self.group_size, int) or self.group_size <= 0:
            raise ValueError("MX / NVFP quantizers need a positive group_size")


class NVFPQuantizer(MXQuantizer):
    _QTYPE = _lib.Q_NVFP
    _ALLOWED = (ElemFormat.fp4_e2m1,)

    def __init__(self, format, group_size=16, axes=-1, zero_point=False, is_profile=False, **kwargs):
        super().__init__(format, group_size, axes, zero_point, is_profile, **kwargs)
        s_ebits, s_mbits, _, s_max_norm, _ = _get_format_params(ElemFormat.fp8_e4m3)
        self.register_buffer("s_ebits", torch.tensor(s_ebits))
        self.register_buffer("s_mbits", torch.tensor(s_mbits))
        self.register_buffer("s_max_norm", torch.tensor(s_max_norm))


    def global_amax(self, x):
        """Phase 1 alone (lcb_nvfp_global_amax): the [1] fp32 amax over the blocks of `x` (ref: nvfp_quant.py:87).  With
        row-sharded weights all-reduce(MAX) it (parallel.allreduce_max_) and pass it to forward(x, nv_amax=...)."""
        if not x.is_cuda:
            raise _lib.LcbError("liblcb200 needs CUDA tensors (no CPU fallback); got device %s" % x.device)
        x = x.contiguous()
        self._resolve_group(x)
        assert self.axes == -1, "global_amax: row-sharded use is along the last axis"
        rows, cols = _prod(tuple(x.shape[:-1])), x.shape[-1]
        out = torch.zeros(1, dtype=torch.float32, device=x.device)
        cfg = self._cfg()
        with torch.cuda.device(x.device):
            rc = _lib.lib().lcb_nvfp_global_amax(ctypes.byref(cfg), _dt(x), _ptr(x), 1, rows, cols, -1, self.group_size,
                                                 _ptr(out), _stream(x.device))
        _lib.check(rc, "lcb_nvfp_global_amax")
        return out

    def forward(self, x, **kwargs):
        nv_amax = kwargs.pop("nv_amax", None)
        if nv_amax is None:
            return super().forward(x, **kwargs)
        self._resolve_group(x)
        out, _, _, _ = qdq_raw(self._cfg(self._mse()), x, self.axes, self.group_size, True, True, nv_amax=nv_amax,
                               check_nan=self.check_nan)
        return out


class DummyQuantizer(nn.Module):  # ref: quantizers/dummy.py:9
    def __init__(self, is_profile=False, **kwargs):
        super().__init__()
        self.is_profile = is_profile
        op_name = kwargs.get("op_name", None)
        self.op_name = op_name if op_name is not None else "None"
        self.save_path = kwargs.get("save_path", "./")

    def configure(self):
        pass

    def find_params(self):
        pass

    def forward(self, x, **kwargs):
        return x

    def fake_quantize(self):
        pass

    def extra_repr(self):
        return "Format: BF16"


def create_fmt_ctx(fmt):  # ref: quantization/quant.py:19-31
    if fmt in ("int4", "int8", "fp4_e2m1", "fp8_e4m3", "fp8_e5m2"):
        return ElemFormat.from_str(fmt)
    raise RuntimeError(f"Invalid format, got {fmt}")


class FakeQuantizer(nn.Module):  # ref: quantization/quant.py:34-63
    @staticmethod
    def build(quant_config, **kwargs):
        quant_config_ = dict(quant_config)
        quant_type = quant_config_.get("type")
        is_profile = quant_config_.get("is_profile")
        op_name = kwargs.get("op_name", None)
        save_path = kwargs.get("save_path", "./")
        if quant_type is None:
            return DummyQuantizer(is_profile=is_profile, op_name=op_name, save_path=save_path)
        if quant_type == "int":
            quantizer = INTQuantizer
        elif quant_type == "fp":
            quantizer = FPQuantizer
        elif quant_type == "mx":
            quantizer = MXQuantizer
        elif quant_type == "nvfp":
            quantizer = NVFPQuantizer
        else:
            raise RuntimeError(f"Unknown Quant type. got {quant_type}")
        quant_config_.pop("type")
        quant_config_["format"] = create_fmt_ctx(quant_config_["format"])
        quant_config_.update(op_name=op_name, save_path=save_path)
        return quantizer(**quant_config_)

"""Packed export of a quantised Linear: real 4 / 8-bit codes + compact scales instead of the fake-quantised
bf16 tensor the reference saves (ref: models/llama.py:210-230; the codes are the intermediate `q` of
quantizers/int_quant.py:210-212 and utils.py:263-272).

    blob = pack_weight(W, quantizer)        # dict of tensors: codes (nibble-packed for 4-bit formats), scales, zeros, meta
    W_dq = unpack_weight(blob)              # == quantizer(W) bit for bit

Codes and scales come from the same lcb_qdq call that produces the fake-quantised tensor, so "integer codes
bit-exact" is checkable from the outside.  Scale storage: INT / FP keep the quantizer's scale and zero-point tensors
(weight dtype); MX stores the shared exponent as one E8M0 byte per block when the scale is a power of two (it is unless
the reference's clamp(min=1e-5) hit, mx_quant.py:151 -- those blobs keep the full scale); NVFP keeps the quantizer's
per-block scale tensor (weight dtype: the product s8 * s32 of nvfp_quant.py:87-100, already clamped) -- storing the
e4m3 factor and the fp32 per-matrix factor separately would not reproduce the clamp(min=1e-5) cases.
Decoding (`unpack_weight`) is plain torch arithmetic in the weight dtype, op for op the reference's `(q - z) * s` /
`q * s + z`; it is the checker of the format, not a hot path.
"""

import torch

from . import _lib
from .quantizers import INTQuantizer, MXQuantizer, NVFPQuantizer, _ptr, _stream

_FP4_LUT = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def pack4(codes):
    """uint8 codes (one per element, low nibble used) -> two per byte."""
    if not codes.is_cuda:
        raise _lib.LcbError("liblcb200 needs CUDA tensors (no CPU fallback)")
    codes = codes.contiguous()
    assert codes.dtype == torch.uint8 and codes.numel() % 2 == 0
    out = torch.empty(codes.numel() // 2, dtype=torch.uint8, device=codes.device)
    with torch.cuda.device(codes.device):
        _lib.check(_lib.lib().lcb_pack4(_ptr(codes), _ptr(out), codes.numel(), _stream(codes.device)), "lcb_pack4")
    return out


def unpack4(packed, numel, signed):
    if not packed.is_cuda:
        raise _lib.LcbError("liblcb200 needs CUDA tensors (no CPU fallback)")
    out = torch.empty(numel, dtype=torch.uint8, device=packed.device)
    with torch.cuda.device(packed.device):
        _lib.check(_lib.lib().lcb_unpack4(_ptr(packed.contiguous()), _ptr(out), numel, int(bool(signed)), _stream(packed.device)),
                   "lcb_unpack4")
    return out


def _is_four_bit(q):
    return q.format.name in ("int4", "fp4_e2m1")


def pack_weight(W, quantizer):
    """Quantise W [N, K] with `quantizer` (an llm_compressor_b200 quantizer, axes = -1) and return the packed blob."""
    assert W.dim() == 2 and quantizer.axes == -1
    if quantizer.group_size == 0:
        raise NotImplementedError("pack_weight: per-tensor scaling (group_size 0) has no per-row scale layout; use a per-row "
                                  "(-1) or grouped quantizer")
    dq, scales, zeros, codes = quantizer.quantize_with_codes(W)
    kind = ("nvfp" if isinstance(quantizer, NVFPQuantizer) else "mx" if isinstance(quantizer, MXQuantizer)
            else "int" if isinstance(quantizer, INTQuantizer) else "fp")
    blob = dict(kind=kind, format=quantizer.format.name, shape=tuple(W.shape), dtype=W.dtype, group_size=int(quantizer.group_size),
                zero_point=bool(quantizer.zero_point))
    four = _is_four_bit(quantizer)
    blob["codes"] = pack4(codes.reshape(-1)) if four else codes.reshape(-1).clone()
    blob["bits"] = 4 if four else 8
    s = scales.reshape(-1)
    blob["scales"] = s.clone()
    blob["zeros"] = zeros.reshape(-1).clone() if quantizer.zero_point else None
    if kind == "mx" and not quantizer.zero_point:
        e = torch.log2(s.float())
        if bool((e == e.round()).all()):  # pure powers of two: E8M0 byte per block (OCP MX), else keep the full scales
            blob["scales_e8m0"] = (e.round() + 127).to(torch.uint8)
            blob["scales"] = None
    return blob, dq


def unpack_weight(blob):
    """Dequantise a blob of pack_weight with torch arithmetic in the weight dtype (checker of the format)."""
    n, k = blob["shape"]
    dt = blob["dtype"]
    numel = n * k
    kind, fmt = blob["kind"], blob["format"]
    codes = unpack4(blob["codes"], numel, signed=(fmt == "int4")) if blob["bits"] == 4 else blob["codes"]
    dev = codes.device
    if blob.get("scales") is not None:
        s = blob["scales"]
    else:
        s = torch.exp2(blob["scales_e8m0"].float() - 127).to(dt)
    gs = blob["group_size"]
    gs = k if gs in (-1,) else gs
    G = -(-k // gs)
    s = s.reshape(n, G, 1).to(dt)
    if fmt.startswith("int"):
        q = codes.view(torch.int8).reshape(n, k).to(dt)
        if kind in ("mx", "nvfp"):
            # MX / NVFP element type int4 / int8 is a FIXED-POINT value: the code is q * 2^(mbits - 2) (qmath.cuh encode_code,
            # ref: utils.py:263-268 with ebits == 0), i.e. int4 codes are in units of 1/4 and int8 codes in units of 1/64
            q = q / (4.0 if fmt == "int4" else 64.0)
    elif fmt == "fp4_e2m1":
        lut = torch.tensor(_FP4_LUT, dtype=torch.float32, device=dev)
        c = codes.reshape(n, k).long()
        q = (lut[c & 7] * torch.where((c & 8) != 0, -1.0, 1.0)).to(dt)
    elif fmt == "fp8_e4m3":
        q = codes.view(torch.float8_e4m3fn).reshape(n, k).to(dt)
    else:
        q = codes.view(torch.float8_e5m2).reshape(n, k).to(dt)
    pad = G * gs - k
    if pad:
        q = torch.nn.functional.pad(q, (0, pad))
    q = q.reshape(n, G, gs)
    if blob["zero_point"]:
        z = blob["zeros"].reshape(n, G, 1).to(dt)
        out = (q - z) * s if kind == "int" else q * s + z  # ref: int_quant.py:212 / fp_quant.py:234
    else:
        out = q * s
    return out.reshape(n, G * gs)[:, :k].contiguous()

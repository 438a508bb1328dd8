"""Quantizer configuration strings of the reference CLI (`--weight int4-g[128]-zp-rw`, `--act-in mxfp4_e2m1-g[32]-rw` ...).

Mirror of the grammar accepted by QuantConfigParser.parse_config (ref: llm_compressor/utils/parser.py:61-110):

    <format>-g[<n>[,<n>...]]-[zp-]<rw|cw>

format: int4 | int8 | fp4_e2m1 | fp8_e4m3 | fp8_e5m2, optionally prefixed with `mx` (MX, scale_ebits = 8) or
`nv` (NVFP); group: one integer (0 per tensor, -1 per token / row, -2 per channel) or a list; `zp` selects the
asymmetric variant; `rw` / `cw` reduce along the last / second-to-last axis (axes -1 / -2).
The result is the dict FakeQuantizer.build expects (ref: quantization/quant.py:36-63).
"""
import re

_PATTERN = re.compile(r"(?P<format>[^-]+)-(?P<group>g\[-?\d+(?:,\d+)*\]+)-(?:(?P<zp>zp)-)?(?P<wise>rw|cw)$")


def quant_type(fmt):
    """ref: utils/parser.py:49-59 (substring tests in this order)"""
    for key, name in (("mx", "mx"), ("nvfp", "nvfp"), ("fp", "fp"), ("int", "int")):
        if key in fmt:
            return name
    raise RuntimeError(f"Invalid format, got {fmt}.")


def parse_quant_config(s, profile=False):
    if s is None:
        return {"type": None, "format": None, "group_size": None, "axes": None, "zero_point": None, "is_profile": profile}
    m = _PATTERN.match(s)
    if not m:
        raise RuntimeError(f"Cannot update Qconfig. No matched pattern, got {s}.")
    fmt = m.group("format")
    qtype = quant_type(fmt)
    cfg = {"type": qtype}
    if qtype == "mx":
        cfg["format"] = fmt.replace("mx", "")
        cfg["scale_ebits"] = 8
    elif qtype == "nvfp":
        cfg["format"] = fmt.replace("nv", "")
    else:
        cfg["format"] = fmt
    nums = re.search(r"g\[(.*?)\]", m.group("group")).group(1).split(",")
    cfg["group_size"] = int(nums[0]) if len(nums) == 1 else [int(n) for n in nums]
    cfg["axes"] = -1 if m.group("wise") == "rw" else -2
    cfg["zero_point"] = m.group("zp") == "zp"
    cfg["is_profile"] = profile
    return cfg


def build_quant_config(weight, act_in=None, act_out=None, head=None, profile=False):
    """{linear, matmul, head} -> {weight, act_in, act_out} dicts, like QuantConfigParser.build_cfg
    (ref: utils/parser.py:26-47; the matmul operands use the activation configs)."""
    p = lambda s: parse_quant_config(s, profile)  # noqa: E731
    return {
        "linear": {"weight": p(weight), "act_in": p(act_in), "act_out": p(act_out)},
        "matmul": {"act_in": p(act_in), "act_out": p(act_out)},
        "head": {"weight": p(head), "act_in": p(None), "act_out": p(None)},
    }

import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))] + [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), p) for p in ("tests", "oracle")]
import numpy as np, torch
import golden_io as gio
from util import t_from_bits, to_f32_np
import llm_compressor_b200 as lc
z, _ = gio.load("elem_core")
bits = np.arange(65536, dtype=np.uint32).astype(np.uint16).reshape(512, 128)
x = t_from_bits(bits, "cuda:0")
for fmt in ["fp4_e2m1", "fp8_e4m3", "fp8_e5m2"]:
    q = lc.FakeQuantizer.build(dict(type="fp", format=fmt, group_size=128, axes=-1, zero_point=False, is_profile=False)).to("cuda:0")
    one = torch.ones(512, 1, 1, dtype=torch.bfloat16, device="cuda:0")
    y = to_f32_np(q(x, scales=one, zeros=torch.zeros_like(one))).reshape(-1)
    ref = gio.bits_to_f32(z[fmt]).reshape(-1)
    bad = np.where(~((y == ref) | (np.isnan(y) & np.isnan(ref))))[0]
    xin = gio.bits_to_f32(bits.reshape(-1))
    print(fmt, len(bad), [(hex(i), float(xin[i]), float(y[i]), float(ref[i])) for i in bad[:8]], [(hex(i)) for i in bad[-4:]])

"""TEST INFRASTRUCTURE ONLY -- golden vectors for find_params with the mse clip search (quantizer.mse = True).

Runs the UNMODIFIED reference quantizers (imported from /root/reference through oracle/ref_shim.py) on seeded
synthetic weights and stores inputs + (scales, zeros) under tests/golden/mse.npz.  Build container only:

    python oracle/gen_golden_mse.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from gen_golden import GOLD, to_bits  # noqa: E402

CASES = [
    ("int4_g128_bf16", dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False), (64, 512), torch.bfloat16),
    ("int4_g128_zp_bf16", dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True), (64, 512), torch.bfloat16),
    ("int4_g128_zp_f32", dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True), (64, 512), torch.float32),
    ("int8_row_f32", dict(type="int", format="int8", group_size=-1, axes=-1, zero_point=False), (48, 384), torch.float32),
    ("fp8e4m3_g128_f32", dict(type="fp", format="fp8_e4m3", group_size=128, axes=-1, zero_point=False), (32, 512), torch.float32),
    ("fp4_g32_zp_bf16", dict(type="fp", format="fp4_e2m1", group_size=32, axes=-1, zero_point=True), (32, 256), torch.bfloat16),
    ("mxfp4_g32_f32", dict(type="mx", format="fp4_e2m1", group_size=32, axes=-1, zero_point=False), (32, 512), torch.float32),
    ("nvfp4_g16_bf16", dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False), (32, 256), torch.bfloat16),
    ("nvfp4_g16_f32", dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False), (32, 256), torch.float32),
    ("mxfp8_g32_bf16", dict(type="mx", format="fp8_e4m3", group_size=32, axes=-1, zero_point=False), (32, 512), torch.bfloat16),
]


def main():
    ref_shim.install()
    out, meta = {}, []
    for n, (name, cfg, shape, dtype) in enumerate(CASES):
        g = torch.Generator().manual_seed(100 + n)
        x = 0.02 * torch.randn(shape, generator=g)
        x = x * (1.0 + 4.0 * (torch.rand(shape, generator=g) > 0.97))   # a few outliers: clipping pays off
        x = x.to(dtype)
        q = ref_shim.build_quantizer(dict(cfg, is_profile=False))
        q.mse = True
        with torch.no_grad():
            s, z = q.find_params(x.clone())
            q2 = ref_shim.build_quantizer(dict(cfg, is_profile=False))
            q2.mse = True
            y = q2(x.clone())
        out[name + "/x"] = to_bits(x)
        out[name + "/scales"] = to_bits(s)
        out[name + "/zeros"] = to_bits(z)
        out[name + "/y"] = to_bits(y)
        meta.append([name, cfg, list(shape), "bf16" if dtype == torch.bfloat16 else "f32"])
        print(name, tuple(s.shape), float(s.float().mean()))
    out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "mse.npz"), **out)
    print("wrote", os.path.join(GOLD, "mse.npz"))


if __name__ == "__main__":
    main()

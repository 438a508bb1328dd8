"""One fake-quant configuration on a [65536, 3072] tensor, a few calls (for ncu captures).  usage: qdq_once.py NAME"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import llm_compressor_b200 as lc
CFG = {
    "int8_tensor": (dict(type="int", format="int8", group_size=0, axes=-1, zero_point=False), torch.bfloat16),
    "int8_channel": (dict(type="int", format="int8", group_size=-2, axes=-2, zero_point=False), torch.bfloat16),
    "int4_g128_fp32": (dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True), torch.float32),
    "nvfp4": (dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False), torch.bfloat16),
    "mxfp4": (dict(type="mx", format="fp4_e2m1", group_size=32, axes=-1, zero_point=False), torch.bfloat16),
    "int4_g128": (dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True), torch.bfloat16),
}
cfg, dt = CFG[sys.argv[1]]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
rows = 65536 if dt == torch.bfloat16 else 32768
x = (0.02 * torch.randn(rows, 3072, generator=g, device=dev)).to(dt)
q = lc.FakeQuantizer.build(dict(cfg, is_profile=False)).to(dev)
q.check_nan = False
for _ in range(3):
    y = q(x)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))

"""Fake-quant bandwidth probe: one format per invocation (so ncu can capture the kernel).
usage: python scripts/qdq_bw.py [name]   (default: all)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import llm_compressor_b200 as lc

CASES = {
    "int4_g128_zp": dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True),
    "int4_g128": dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False),
    "int8_g128": dict(type="int", format="int8", group_size=128, axes=-1, zero_point=False),
    "int8_tok": dict(type="int", format="int8", group_size=-1, axes=-1, zero_point=False),
    "int4_tok": dict(type="int", format="int4", group_size=-1, axes=-1, zero_point=False),
    "fp8_tok": dict(type="fp", format="fp8_e4m3", group_size=-1, axes=-1, zero_point=False),
    "mxfp4": dict(type="mx", format="fp4_e2m1", group_size=32, axes=-1, zero_point=False),
    "mxfp8": dict(type="mx", format="fp8_e4m3", group_size=32, axes=-1, zero_point=False),
    "nvfp4": dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False),
    "int8_tensor": dict(type="int", format="int8", group_size=0, axes=-1, zero_point=False),
    "int8_chan": dict(type="int", format="int8", group_size=-2, axes=-2, zero_point=False),
}
dev = torch.device("cuda:0")
names = sys.argv[1:] or list(CASES)
g = torch.Generator(device=dev).manual_seed(0)
x = (0.02 * torch.randn(8 * 8192, 3072, generator=g, device=dev)).to(torch.bfloat16)  # 403 MB, > L2
res = {}
for name in names:
    cfg = dict(CASES[name], is_profile=False)
    q = lc.FakeQuantizer.build(cfg).to(dev)
    q.check_nan = False
    for _ in range(3):
        y = q(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        y = q(x)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    res[name] = {"ms": ms, "GBs_algorithmic_4B_per_elem": x.numel() * 4 / ms / 1e6}
print(json.dumps(res))

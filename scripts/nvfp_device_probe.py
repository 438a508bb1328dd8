"""Is the reference's NVFP quantizer device dependent?  Runs the reference's own NVFPQuantizer (oracle/_ref) on CPU and on
CUDA on the same bf16 tensor and tests the hypothesis that CUDA eager rounds the 0-dim fp32 scale s32 to bf16 first."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import torch
import ref_shim
cfg = dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False, is_profile=False)
g = torch.Generator().manual_seed(1)
x = (0.02 * torch.randn(256, 1024, generator=g)).to(torch.bfloat16)
q = ref_shim.build_quantizer(cfg)
y_cpu = q(x)
y_gpu = ref_shim.build_quantizer(cfg).to("cuda")(x.cuda()).cpu()
import llm_compressor_b200 as lc
y_ours = lc.FakeQuantizer.build(cfg).to("cuda")(x.cuda()).cpu()
print("reference cpu vs reference cuda: differing fraction", float((y_cpu != y_gpu).float().mean()))
print("ours vs reference cpu:", float((y_ours != y_cpu).float().mean()), " ours vs reference cuda:", float((y_ours != y_gpu).float().mean()))
a = torch.tensor(3.0, dtype=torch.bfloat16); s = torch.tensor(1.0 / 3.0, dtype=torch.float32)
v = torch.full((8,), 1.2345, dtype=torch.bfloat16)
sc = torch.tensor(0.0123456789, dtype=torch.float32)
print("bf16 tensor / fp32 0-dim: cpu", (v / sc)[0].item(), " cuda", (v.cuda() / sc.cuda())[0].item(),
      " pre-rounded scalar", (v / sc.to(torch.bfloat16))[0].item())

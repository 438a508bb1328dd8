"""One GPTQ update_weight (+ factorize) on a Llama-3.2-3B shape, for ncu launch lists. Development aid.
usage: update_once.py N K [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import llm_compressor_b200 as lc
from llm_compressor_b200 import ops, solvers

N, K = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(2 * K, K, generator=g, device=dev).to(torch.bfloat16)
H = torch.zeros(K, K, device=dev)
ops.hessian_add(H, X, 2.0 / X.shape[0], 0.0)
W = (0.02 * torch.randn(N, K, generator=g, device=dev)).to(torch.bfloat16)
cfg = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)


class Lin(torch.nn.Module):
    pass


for _ in range(reps):
    fac = solvers.factorize(H.clone(), 128, True, 0.01)
    lin = Lin()
    lin.weight = torch.nn.Parameter(W.clone(), requires_grad=False)
    lin.weight_quantizer = lc.FakeQuantizer.build(cfg).to(dev)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    solvers.update_weight(lin, dev, actorder=True, factor=fac)
    e.record()
    torch.cuda.synchronize()
    print("update ms", s.elapsed_time(e))
print("ok", float(lin.weight.float().abs().mean()))

"""A few Hessian accumulations (symmetric-half raw sums, as the bench does) for ncu. usage: hess_once.py K [reps] [defer]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_compressor_b200 import ops
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
defer = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(reps, 2048, K, generator=g, device=dev).to(torch.bfloat16)
H = torch.zeros(K, K, device=dev)
n = 0
acc = ops.HessianAccumulator(H, defer)
for j in range(reps):
    acc.add(X[j])
n = acc.flush()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for j in range(reps):
    acc.add(X[j])
n = acc.flush()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
print("K=%d defer=%d upper-half raw: %.4f ms/sample  algorithmic %.0f TFLOP/s" % (K, defer, ms, 2 * 2048 * K * K / ms / 1e9))
ops.hessian_finalize(H, 2.0 / n, True)
ref = 2.0 / n * 2 * (X[:, :, 128:384].double().reshape(-1, 256).T @ X[:, :, 1000:1300].double().reshape(-1, 300))
got = H[128:384, 1000:1300].double()
print("relF vs fp64 (256 x 300 block): %.2e" % float((got - ref).norm() / ref.norm()))

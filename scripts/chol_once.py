import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_compressor_b200 import ops
K = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
g = torch.Generator(device="cuda:0").manual_seed(0)
X = torch.randn(2 * K, K, generator=g, device="cuda:0").to(torch.bfloat16).float()
H = (X.T @ X / K).contiguous()
U = torch.empty_like(H)
for _ in range(2):
    U = ops.chol_inv_upper(H, percdamp=0.01, out=U)
torch.cuda.synchronize()
t0 = time.perf_counter()
U = ops.chol_inv_upper(H, percdamp=0.01, out=U)   # includes the status read-back (host sync)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("chol K=%d: host+device %.3f ms" % (K, (t1 - t0) * 1e3))
print("ok", float(U[0, 0]))

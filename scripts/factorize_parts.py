"""Where the `factorize` stage goes: solvers.factorize (dead fix, act-order sort, Cholesky-inverse) against its parts,
CUDA events, per call.  Development aid."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_compressor_b200 import ops, solvers

dev = torch.device("cuda:0")


def timeit(fn, n=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


out = {}
for K in (3072, 8192):
    g = torch.Generator(device=dev).manual_seed(K)
    X = torch.randn(2 * K, K, generator=g, device=dev).to(torch.bfloat16).float()
    H = ((1.0 / K) * X.T @ X).contiguous()
    del X
    r = {}

    def fac():
        f = solvers.factorize(H, 128, actorder=True, percdamp=0.01)
        f.resolve()
        return f

    def fac_noresolve():
        return solvers.factorize(H, 128, actorder=True, percdamp=0.01)

    f = fac()
    r["factorize_ms"] = timeit(fac)
    r["factorize_deferred_ms"] = timeit(fac_noresolve)
    r["chol_inv_upper_perm_ms"] = timeit(lambda: ops.chol_inv_upper(H, perm=f.col_perm, percdamp=0.01, defer=True))
    r["chol_inv_upper_noperm_ms"] = timeit(lambda: ops.chol_inv_upper(H, percdamp=0.01, defer=True))
    r["dead_fix_ms"] = timeit(lambda: ops.dead_fix(H))

    def sort_part():
        diag = torch.diag(H)
        perm = torch.argsort(diag.reshape(-1, 128).sum(-1), descending=True)
        col_perm = (perm.unsqueeze(1) * 128 + torch.arange(128, device=perm.device)).reshape(-1)
        return torch.argsort(perm), col_perm
    r["act_order_sort_ms"] = timeit(sort_part)
    r["zeros_KxK_ms"] = timeit(lambda: torch.zeros(K, K, device=dev))
    out["K%d" % K] = r
print(json.dumps(out, indent=1))

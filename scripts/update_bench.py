"""GPTQ update stage per Linear shape (CUDA events): the block loop alone (lcb_gptq_update) and the whole update_weight
(gather, find_params, loop, scatter).  Development aid."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import llm_compressor_b200 as lc
from llm_compressor_b200 import ops, solvers

dev = torch.device("cuda:0")
cfg = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)


def timeit(fn, n=3, warm=1):
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


class Lin(torch.nn.Module):
    pass


out = {}
for N, K in ((5120, 3072), (3072, 3072), (16384, 3072), (3072, 8192)):
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(2 * K, K, generator=g, device=dev).to(torch.bfloat16)
    H = torch.zeros(K, K, device=dev)
    ops.hessian_add(H, X, 2.0 / X.shape[0], 0.0)
    del X
    W = (0.02 * torch.randn(N, K, generator=g, device=dev)).to(torch.bfloat16)
    fac = solvers.factorize(H.clone(), 128, True, 0.01)
    fac.resolve()
    q = lc.FakeQuantizer.build(cfg).to(dev)
    Wp, keep = ops.gptq_gather(W, fac.col_perm, fac.dead)
    s, z = q.find_params(Wp)
    s2, z2 = s.float().reshape(N, K // 128).contiguous(), z.float().reshape(N, K // 128).contiguous()
    res = {}
    res["block_loop_ms"] = timeit(lambda: ops.gptq_block_update(q._cfg(), Wp.clone(), fac.U, s2, z2, keep, 128))
    res["clone_ms"] = timeit(lambda: Wp.clone())

    def full():
        lin = Lin()
        lin.weight = torch.nn.Parameter(W.clone(), requires_grad=False)
        lin.weight_quantizer = lc.FakeQuantizer.build(cfg).to(dev)
        solvers.update_weight(lin, dev, actorder=True, factor=fac)
    res["update_weight_ms"] = timeit(full)
    out["%dx%d" % (N, K)] = res
    del H, W, Wp, keep, fac
    torch.cuda.empty_cache()
print(json.dumps(out))

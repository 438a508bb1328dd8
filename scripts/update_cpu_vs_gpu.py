"""Is the GPTQ block loop bound by the host (launch + tensor-map encoding) or by the device?  For each Linear shape:
host time to ENQUEUE one lcb_gptq_update (no synchronisation inside) next to the device time between CUDA events.
Development aid."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import llm_compressor_b200 as lc
from llm_compressor_b200 import ops, solvers

dev = torch.device("cuda:0")
cfg = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)
out = {}
for N, K in ((640, 3072), (5120, 3072), (3072, 3072), (16384, 3072), (384, 8192), (3072, 8192)):
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
    Wc = [Wp.clone() for _ in range(4)]
    ops.gptq_block_update(q._cfg(), Wc[0], fac.U, s2, z2, keep, 128)
    torch.cuda.synchronize()
    host, devt = [], []
    for i in range(1, 4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        t0 = time.perf_counter()
        ops.gptq_block_update(q._cfg(), Wc[i], fac.U, s2, z2, keep, 128)
        t1 = time.perf_counter()
        b.record()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3)
        devt.append(a.elapsed_time(b))
    out["%dx%d" % (N, K)] = {"host_enqueue_ms": min(host), "device_ms": min(devt), "blocks": K // 128,
                             "host_us_per_block": min(host) * 1e3 / (K // 128), "device_us_per_block": min(devt) * 1e3 / (K // 128)}
    del H, W, Wp, keep, fac, Wc
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))

"""Per-stage device timings on Llama-3.2-3B shapes (CUDA events, warm). Development aid."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import llm_compressor_b200 as lc
from llm_compressor_b200 import ops, solvers

dev = torch.device("cuda:0")


def timeit(fn, n=5, warm=2):
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


res = {}
g = torch.Generator(device=dev).manual_seed(0)
for K in (3072, 8192):
    X = torch.randn(16, 2048, K, generator=g, device=dev).to(torch.bfloat16)
    H = torch.zeros(K, K, device=dev)
    st = {"n": 0}

    def hs():
        for j in range(16):
            st["n"] = ops.hessian_accum(H, X[j], st["n"])
    ms = timeit(hs, n=3, warm=1) / 16
    res[f"hessian_K{K}_ms_per_sample"] = ms
    res[f"hessian_K{K}_TFLOPs"] = 2 * 2048 * K * K / ms / 1e9
    Hc = H.clone()
    res[f"chol_K{K}_ms"] = timeit(lambda: ops.chol_inv_upper(Hc, percdamp=0.01), n=3, warm=1)
    fac = solvers.factorize(H.clone(), 128, True, 0.01)
    for N in ((3072, 1024, 8192) if K == 3072 else (3072,)):
        W = (0.02 * torch.randn(N, K, generator=g, device=dev)).to(torch.bfloat16)

        class Lin(torch.nn.Module):
            pass

        def upd():
            lin = Lin()
            lin.weight = torch.nn.Parameter(W.clone(), requires_grad=False)
            lin.weight_quantizer = lc.FakeQuantizer.build(dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)).to(dev)
            solvers.update_weight(lin, dev, actorder=True, factor=fac)
        res[f"update_N{N}_K{K}_ms"] = timeit(upd, n=3, warm=1)

# fake-quant bandwidth
for name, cfg in [("int4_g128_zp", dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True)),
                  ("int4_g128", dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False)),
                  ("int8_tok", dict(type="int", format="int8", group_size=-1, axes=-1, zero_point=False)),
                  ("fp8_tok", dict(type="fp", format="fp8_e4m3", group_size=-1, axes=-1, zero_point=False)),
                  ("mxfp4", dict(type="mx", format="fp4_e2m1", group_size=32, axes=-1, zero_point=False)),
                  ("mxfp8", dict(type="mx", format="fp8_e4m3", group_size=32, axes=-1, zero_point=False)),
                  ("nvfp4", dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False))]:
    cfg["is_profile"] = False
    x = (0.02 * torch.randn(8 * 8192, 3072, generator=g, device=dev)).to(torch.bfloat16)  # 402 MB > L2
    q = lc.FakeQuantizer.build(cfg).to(dev)
    q.check_nan = False
    ms = timeit(lambda: q(x), n=5, warm=2)
    res[f"qdq_{name}_GBs"] = x.numel() * 4 / ms / 1e6
print(json.dumps(res, indent=1))

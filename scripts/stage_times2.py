"""Device timings of the non-headline stages at Llama-3.2-3B shapes (CUDA events, warm). Development aid."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import llm_compressor_b200 as lc
from llm_compressor_b200 import ops, solvers

dev = torch.device("cuda:0")


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


res = {}
g = torch.Generator(device=dev).manual_seed(0)
only = sys.argv[1:]
for K in (3072, 8192):
    X = torch.randn(4, 2048, K, generator=g, device=dev).to(torch.bfloat16)
    Xfp = (X.float() + 0.05 * torch.randn(4, 2048, K, generator=g, device=dev)).to(torch.bfloat16)
    H = torch.zeros(K, K, device=dev)
    D = torch.zeros(K, K, device=dev)
    n = 0
    for j in range(4):
        n = ops.hessian_accum_raw(H, X[j], n, dxxt=D, x_fp=Xfp[j])
    res[f"hessian_dxxt_K{K}_ms_per_sample"] = timeit(lambda: ops.hessian_accum_raw(H, X[0], 0, dxxt=D, x_fp=Xfp[0]), n=5)
    res[f"hessian_only_K{K}_ms_per_sample"] = timeit(lambda: ops.hessian_accum_raw(H, X[0], 0), n=5)
    s_row = torch.zeros(K, device=dev)
    res[f"rownorm_K{K}_ms_per_sample"] = timeit(lambda: ops.rownorm_accum(s_row, X[0], 0), n=5)
    ops.hessian_finalize(H, 2.0 / n, True)
    ops.hessian_finalize(D, 2.0 / n, False)
    N = 3072
    W = (0.02 * torch.randn(N, K, generator=g, device=dev)).to(torch.bfloat16)

    def sparse():
        lay = solvers.Wrapper(torch.nn.Linear(K, N, bias=False, dtype=torch.bfloat16, device=dev), dev)
        lay.module.weight.data = W.clone()
        lay.H = H.clone()
        solvers.prune_weight(lay, dev, 0.5)
    res[f"sparsegpt_N{N}_K{K}_ms (incl. chol)"] = timeit(sparse, n=2)
    U = ops.chol_inv_upper(H.clone(), percdamp=0.01)
    Wf = W.float()
    res[f"sparsegpt_blockloop_N{N}_K{K}_ms"] = timeit(lambda: ops.sparsegpt_update(Wf.clone(), U, 0.5), n=2)

    def gptq(cfg, fac, alpha=None):
        def f():
            lin = Lin()
            lin.weight = torch.nn.Parameter(W.clone(), requires_grad=False)
            lin.weight_quantizer = lc.FakeQuantizer.build(dict(cfg, is_profile=False)).to(dev)
            if alpha is None:
                solvers.update_weight(lin, dev, actorder=True, factor=fac)
            else:
                solvers.gptaq_update_weight(lin, dev, actorder=True, alpha=alpha, factor=fac)
        return f
    fac128 = solvers.factorize(H.clone(), 128, True, 0.01)
    res[f"gptq_int4_g128_N{N}_K{K}_ms"] = timeit(gptq(dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False), fac128), n=2)
    fac32 = solvers.factorize(H.clone(), 32, True, 0.01)
    res[f"gptq_mxfp4_g32_N{N}_K{K}_ms"] = timeit(gptq(dict(type="mx", format="fp4_e2m1", group_size=32, axes=-1, zero_point=False), fac32), n=2)
    fac16 = solvers.factorize(H.clone(), 16, True, 0.01)
    res[f"gptq_nvfp4_g16_N{N}_K{K}_ms"] = timeit(gptq(dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False), fac16), n=2)
    facrow = solvers.factorize(H.clone(), -1, True, 0.01)
    res[f"gptq_int4_row_N{N}_K{K}_ms"] = timeit(gptq(dict(type="int", format="int4", group_size=-1, axes=-1, zero_point=True), facrow), n=2)
    res[f"gptaq_factorize_K{K}_ms"] = timeit(lambda: solvers.factorize(H.clone(), 128, True, 0.01, dXXT=D.clone(), alpha=0.25), n=2)
    facaq = solvers.factorize(H.clone(), 128, True, 0.01, dXXT=D.clone(), alpha=0.25)
    res[f"gptaq_int4_g128_N{N}_K{K}_ms"] = timeit(gptq(dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False), facaq, alpha=0.25), n=2)
    del fac128, fac32, fac16, facrow, facaq
    if K == 3072:
        W2 = (0.02 * torch.randn(8192, K, generator=g, device=dev)).to(torch.bfloat16)
        sr = torch.rand(K, device=dev) + 0.1
        res["mask_wanda_8192x3072_ms"] = timeit(lambda: ops.mask_wanda(W2, sr, 0.5), n=3)
        res["mask_ria_8192x3072_ms"] = timeit(lambda: ops.mask_ria(W2, sr, 0.5, 0.5), n=3)
        res["mask_magnitude_8192x3072_ms"] = timeit(lambda: ops.mask_magnitude(W2, 0.5), n=3)
    print(json.dumps(res, indent=1), flush=True)

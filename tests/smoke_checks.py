"""TEST INFRASTRUCTURE (called by __graft_entry__.smoke()): one small invocation of each stage of the hot path on the
GPU, checked against the CPU oracle.  Lives outside the product package because it imports oracle/."""
import numpy as np
import torch


def run(dev):
    import oracle as orc
    from llm_compressor_b200 import FakeQuantizer, ops, solvers

    g = torch.Generator().manual_seed(0)
    K, N, T = 256, 64, 256
    X = (torch.randn(2, T, K, generator=g) * torch.exp(0.5 * torch.randn(K, generator=g))).to(torch.bfloat16)
    W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16)
    cfg = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)

    # (a) Hessian on the tensor cores, hook semantics
    lin = torch.nn.Linear(K, N, bias=False, dtype=torch.bfloat16, device=dev)
    lin.weight.data = W.to(dev)
    lin.weight_quantizer = FakeQuantizer.build(cfg).to(dev)
    lin.weight_quantizer.nsamples = 0
    lin.weight_quantizer.H = torch.zeros(K, K, device=dev)
    Href = np.zeros((K, K), np.float32)
    n = 0
    for j in range(2):
        solvers.cache_hessian_weight(lin, (X[j].to(dev).unsqueeze(0),), None)
        n = orc.hessian_accum(Href, X[j].float().numpy(), n)
    H = solvers.finalize_hessian(lin.weight_quantizer).clone()
    rel = float(np.linalg.norm(H.cpu().numpy() - Href) / np.linalg.norm(Href))
    assert rel < 1e-5, "Hessian differs from the oracle: %g" % rel

    # (b) GPTQ solve
    ref = orc.gptq_update(W.float().numpy(), Href.copy(), cfg)
    solvers.update_weight(lin, dev, actorder=True)
    got = lin.weight.data.float().cpu().numpy()
    assert float(np.mean(got != ref)) < 2e-2, "GPTQ result differs from the oracle"

    # (d) Wanda mask
    s = torch.rand(K, generator=g) + 0.1
    m = ops.mask_wanda(W.to(dev), s.to(dev), 0.5).cpu().numpy()
    assert np.array_equal(m, orc.mask_wanda(W.float().numpy(), s.numpy(), 0.5)), "Wanda mask differs"

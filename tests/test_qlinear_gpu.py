"""f3: fused activation fake-quant + GEMM (lcb_qlinear_fwd) against the two-kernel form the reference runs,
F.linear(input_quantizer(x), W, b) (ref: modules/qlinear.py:86-88), with the activation quantised by this package's
quantizer kernel (bit-exact vs the reference, test_qdq_gpu.py).  Stated tolerance: identical up to the fp32 accumulation
order of the GEMM -- every output within one bf16 ulp of the cuBLAS result (two ulp where the bias is added), relative
Frobenius error < 2e-3; and the operand really is the quantised activation: an int8 vs int4 quantiser changes the
output by the quantisation error, a plain GEMM does not match."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _q(fmt, gs, zp):
    import llm_compressor_b200 as lc
    return lc.FakeQuantizer.build(dict(type="int", format=fmt, group_size=gs, axes=-1, zero_point=zp, is_profile=False)).to(DEV)


CASES = [("int8", -1, False), ("int8", 128, False), ("int4", -1, True), ("int4", 128, True), ("int8", 64, True), ("int4", 256, False)]


@pytest.mark.parametrize("fmt,gs,zp", CASES, ids=["%s-g%d%s" % (f, g, "-zp" if z else "") for f, g, z in CASES])
@pytest.mark.parametrize("shape", [(2048, 3072, 3072), (300, 1024, 512), (2048, 8192, 3072), (129, 384, 200)],
                         ids=["T2048-N3072-K3072", "T300-N1024-K512", "T2048-N8192-K3072", "T129-N384-K200rag"])
def test_fused_qlinear_matches_two_kernel_form(fmt, gs, zp, shape):
    from llm_compressor_b200 import ops
    T, N, K = shape
    if K % 64 or (gs > 0 and K % gs):
        pytest.skip("geometry outside the fused path (falls back to quantizer + F.linear)")
    g = torch.Generator().manual_seed(T + N)
    x = (torch.randn(1, T, K, generator=g) * torch.exp(0.5 * torch.randn(K, generator=g))).to(torch.bfloat16).to(DEV)
    W = (0.03 * torch.randn(N, K, generator=g)).to(torch.bfloat16).to(DEV)
    b = (0.1 * torch.randn(N, generator=g)).to(torch.bfloat16).to(DEV)
    q = _q(fmt, gs, zp)
    assert ops.qlinear_fusable(x, W, q)
    xq = _q(fmt, gs, zp)(x)
    for bias in (None, b):
        ref = F.linear(xq, W, bias)
        got = ops.qlinear_forward(x, W, bias, q)
        assert got.shape == ref.shape and got.dtype == torch.bfloat16
        r32, g32 = ref.float(), got.float()
        ulp = r32.abs().clamp_min(1e-3) * 2.0 ** -7
        worst = float(((r32 - g32).abs() / ulp).max())
        rel = float((r32 - g32).norm() / r32.norm())
        same = float((ref == got).float().mean())
        print(f"{fmt} g{gs} zp={zp} {shape} bias={bias is not None}: identical {same:.4f}, worst {worst:.2f} bf16 ulp, relF {rel:.2e}")
        assert worst <= (2.0 if bias is not None else 1.0) + 1e-3 and rel < 2e-3
        # against an fp64 product of the SAME quantised operand, ours is as close as cuBLAS is
        ex = (xq.double().reshape(-1, K) @ W.double().T + (0 if bias is None else bias.double())).reshape(ref.shape)
        assert float((g32.double() - ex).norm()) <= 1.05 * float((r32.double() - ex).norm()) + 1e-6
    plain = F.linear(x, W)
    assert float((ops.qlinear_forward(x, W, None, None).float() - plain.float()).norm() / plain.float().norm()) < 2e-3
    assert float((got.float() - F.linear(x, W, b).float()).norm() / plain.float().norm()) > 1e-3   # the QDQ really happened


def test_qlinear_module_uses_the_fused_kernel_and_matches():
    from llm_compressor_b200 import _lib, modules
    cfg = dict(act_in=dict(type="int", format="int8", group_size=-1, axes=-1, zero_point=False, is_profile=False),
               weight=dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False),
               act_out=dict(type=None, format=None, group_size=None, axes=None, zero_point=None, is_profile=False))
    lin = torch.nn.Linear(1024, 768, bias=True)
    x = torch.randn(1, 512, 1024).to(torch.bfloat16).to(DEV)
    m = modules.QLinear(lin, cfg, torch.bfloat16, op_name="t").to(DEV)
    L = _lib.lib()
    prev = modules.FUSED_ACT_QDQ
    modules.FUSED_ACT_QDQ = True
    try:
        n0 = L.lcb_launch_count()
        y = m(x)
        launches = L.lcb_launch_count() - n0
    finally:
        modules.FUSED_ACT_QDQ = prev
    modules.FUSED_ACT_QDQ = False
    try:
        y2 = m(x)
    finally:
        modules.FUSED_ACT_QDQ = prev
    assert launches == 2                     # find-only pass + the fused GEMM
    assert float((y.float() - y2.float()).abs().max()) <= float(y2.float().abs().max()) * 2.0 ** -7

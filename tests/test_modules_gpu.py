"""c6: the operator wrappers.  QLinear / QMatmul of this package against the REFERENCE's own modules
(modules/qlinear.py, modules/qmatmul.py, imported from oracle/_ref or /root/reference) running their own quantizers in
PyTorch eager on the same GPU, at attention shapes: Q / K^T grouped along the last dim, V in groups of 128 DOWN the rows
(`axes = -2`), per-token (-1 -> -2 flip) and block-scaled formats; plus the reference's classes bound to this package's
FakeQuantizer (modules.bind_reference).  Outputs must be bit-identical: the quantised operands are (test_qdq_gpu.py)
and the matmul is the same torch call."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shim.available(), reason="no reference copy (oracle/_ref) on this box")]
DEV = "cuda:0"


def _ref_modules():
    ref_shim.install()
    import llm_compressor.quantization.calibrations.gptq.core  # noqa: F401  (sys.path hack of the reference)
    import llm_compressor.modules.qlinear as RL
    import llm_compressor.modules.qmatmul as RM
    from llm_compressor.utils.parser import QuantConfigParser
    return RL, RM, QuantConfigParser


def _cfgs(QuantConfigParser, act):
    qc = QuantConfigParser().build_cfg("int4-g[128]-rw", act, None, None)
    return ref_shim.EasyDict(qc.linear), ref_shim.EasyDict(qc.matmul)


ACTS = ["int8-g[128]-rw", "int8-g[-1]-rw", "int4-g[-1]-zp-rw", "mxfp8_e4m3-g[32]-rw", "nvfp4_e2m1-g[16]-rw", "fp8_e4m3-g[-1]-rw"]


@pytest.mark.parametrize("act", ACTS)
@pytest.mark.parametrize("axes", [-1, -2])
def test_qmatmul_matches_reference_module(act, axes):
    from llm_compressor_b200 import modules
    RL, RM, P = _ref_modules()
    _, mm_cfg = _cfgs(P, act)
    g = torch.Generator().manual_seed(5)
    h, T, d = 4, 256, 128
    a = torch.randn(1, h, T, d if axes == -1 else T, generator=g).to(torch.bfloat16).to(DEV)
    b = torch.randn(1, h, d if axes == -1 else T, T if axes == -1 else d, generator=g).to(torch.bfloat16).to(DEV)
    if axes == -2:
        a = torch.softmax(a.float(), -1).to(torch.bfloat16)            # attention probabilities @ V
    ref = RM.QMatmul(ref_shim.EasyDict({k: dict(v) if isinstance(v, dict) else v for k, v in mm_cfg.items()}), axes=axes,
                     op_name="t").to(DEV)
    ours = modules.QMatmul(mm_cfg, axes=axes, op_name="t").to(DEV)
    prev = modules.bind_reference(RM)
    try:
        bound = RM.QMatmul(ref_shim.EasyDict({k: dict(v) if isinstance(v, dict) else v for k, v in mm_cfg.items()}), axes=axes,
                           op_name="t").to(DEV)
    finally:
        RM.FakeQuantizer = prev[0]
    assert type(bound.input2_quantizer).__module__.startswith("llm_compressor_b200")
    assert ours.input2_quantizer.axes == ref.input2_quantizer.axes and ours.input2_quantizer.group_size == ref.input2_quantizer.group_size
    y_ours, y_bound = ours(a, b), bound(a, b)
    assert torch.equal(y_ours, y_bound)
    # operand level against the reference module on the CPU (the pinned semantics: for NVFP the reference itself is
    # device dependent, see tests/test_reference_dropin_gpu.py): both quantised operands bit for bit ...
    ref = ref.cpu()
    qa_ref, qb_ref = ref.input1_quantizer(a.cpu()), ref.input2_quantizer(b.cpu())
    assert torch.equal(qa_ref, ours.input1_quantizer(a).cpu())
    assert torch.equal(qb_ref, ours.input2_quantizer(b).cpu())
    # ... and the product is the same torch.matmul of them
    assert torch.equal(y_ours, torch.matmul(qa_ref.to(DEV), qb_ref.to(DEV)))
    if not act.startswith("nvfp"):
        assert torch.equal(ref.to(DEV)(a, b), y_ours)     # the reference module end to end in CUDA eager


@pytest.mark.parametrize("act", [None] + ACTS[:4])
def test_qlinear_matches_reference_module(act):
    from llm_compressor_b200 import modules
    RL, RM, P = _ref_modules()
    lin_cfg, _ = _cfgs(P, act)
    g = torch.Generator().manual_seed(9)
    lin = torch.nn.Linear(512, 384, bias=True)
    lin.weight.data = 0.05 * torch.randn(384, 512, generator=g)
    x = torch.randn(1, 300, 512, generator=g).to(torch.bfloat16).to(DEV)
    ref = RL.QLinear(linear=lin, quant_config=lin_cfg, dtype=torch.bfloat16, op_name="t").to(DEV)
    ours = modules.QLinear(lin, lin_cfg, torch.bfloat16, op_name="t").to(DEV)
    assert torch.equal(ref.weight.data, ours.weight.data) and torch.equal(ref.bias.data, ours.bias.data)
    assert torch.equal(ref(x), ours(x))
    assert torch.equal(ref.weight_quantizer(ref.weight.data), ours.weight_quantizer(ours.weight.data))
    with pytest.raises(NotImplementedError):
        ours(x, R1=torch.eye(512, device=DEV))

"""Fake-quant operator wrappers around the CUDA quantizers.

The reference's wrappers (modules/qlinear.py, modules/qmatmul.py) are torch glue that stays the reference's own
code in a drop-in deployment: `bind_reference()` points THEIR modules at this package's `FakeQuantizer` factory and
nothing else changes (tests/test_reference_dropin_gpu.py runs the reference's QLinear that way).  For use without the
reference tree, `QLinear` / `QMatmul` below offer the same constructor arguments and attribute names
(`input_quantizer`, `weight_quantizer`, `output_quantizer`; `input1_quantizer`, `input2_quantizer`) on a small
slot-table design; the SpinQuant-training online rotation kwargs of the reference's forward are out of scope here
(offline rotation: llm_compressor_b200.hadamard.rotate_model).
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .quantizers import FakeQuantizer

# True: INT per-token / per-group activation quantisers in front of a bf16 Linear run as ONE kernel (ops.qlinear_forward,
# bit-identical output).  Default False: measured on B200 (bench.py `fused_act_qdq_linear`) the single-CTA tcgen05 GEMM of
# lcb_qlinear_fwd reaches 0.35-0.59 PFLOP/s while quantizer kernel + cuBLAS (2-CTA tiles) runs the same call at the
# equivalent of 0.9-1.4 PFLOP/s -- at calibration shapes the activation fits L2, so the saved HBM round trip does not pay
# for the slower GEMM yet.
FUSED_ACT_QDQ = False


def bind_reference(*reference_modules):
    """Swap this package's quantizer factory into reference modules that did `from quantization.quant import
    FakeQuantizer` (modules/qlinear.py:13, modules/qmatmul.py:13); returns the previous bindings for undoing."""
    previous = [getattr(m, "FakeQuantizer", None) for m in reference_modules]
    for m in reference_modules:
        m.FakeQuantizer = FakeQuantizer
    return previous


def _cfg(quant_config, key):
    return quant_config[key] if isinstance(quant_config, dict) else getattr(quant_config, key)


def _slots(owner, quant_config, table, kwargs):
    """table: {attribute name: (config key, op-name suffix, overrides or None)} -> quantizer sub-modules on `owner`."""
    op, path = kwargs.get("op_name"), kwargs.get("save_path", "./")
    for attr, (key, suffix, override) in table.items():
        cfg = _cfg(quant_config, key)
        if override:
            cfg = dict(cfg, **override)
        setattr(owner, attr, FakeQuantizer.build(cfg, op_name=f"{op}.{suffix}", save_path=path))


class QLinear(nn.Linear):
    """y = Q_out(linear(Q_in(x), W, b)); `weight_quantizer` is applied by the calibration drivers, not in forward
    (ref: modules/qlinear.py:86-88)."""

    def __init__(self, linear, quant_config, dtype, **kwargs):
        super().__init__(linear.in_features, linear.out_features, linear.bias is not None, linear.weight.device, dtype)
        self.train(linear.training)
        self.load_state_dict({k: v.to(dtype) for k, v in linear.state_dict().items()})
        _slots(self, quant_config, {"input_quantizer": ("act_in", "input", None), "weight_quantizer": ("weight", "weight", None),
                                    "output_quantizer": ("act_out", "output", None)}, kwargs)

    def forward(self, inputs, **kwargs):
        if kwargs.get("R1") is not None:
            raise NotImplementedError("online R1 / R2 rotation belongs to the SpinQuant training model (out of scope); "
                                      "rotate offline with llm_compressor_b200.hadamard.rotate_model")
        if FUSED_ACT_QDQ and ops.qlinear_fusable(inputs, self.weight, self.input_quantizer):
            # activation fake-quant in the operand prologue of a tcgen05 GEMM: QDQ(x) is never written to HBM
            return self.output_quantizer(ops.qlinear_forward(inputs, self.weight, self.bias, self.input_quantizer))
        return self.output_quantizer(F.linear(self.input_quantizer(inputs), self.weight, self.bias))


class QMatmul(nn.Module):
    """Q_out(Q_1(a) @ Q_2(b)).  The second operand is grouped along `axes`: -1 for K^T in Q K^T (reduction dim last), -2 for
    V in S V (groups run down the rows); a per-row / per-column group size follows the axis (ref: modules/qmatmul.py:34-46)."""

    def __init__(self, quant_config, axes=-1, **kwargs):
        super().__init__()
        gs = dict(_cfg(quant_config, "act_in")).get("group_size")
        follow = {(-1, -2): -1, (-2, -1): -2}.get((axes, gs))
        second = {"axes": axes} if follow is None else {"axes": axes, "group_size": follow}
        _slots(self, quant_config, {"input1_quantizer": ("act_in", "input1", None), "input2_quantizer": ("act_in", "input2", second),
                                    "output_quantizer": ("act_out", "output", None)}, kwargs)

    def forward(self, inputs1, inputs2):
        return self.output_quantizer(torch.matmul(self.input1_quantizer(inputs1).to(inputs2), self.input2_quantizer(inputs2)))

"""Fake-quant wrapper modules with the reference's interface.

ref: llm_compressor/modules/qlinear.py:16-88 (QLinear), llm_compressor/modules/qmatmul.py:16-65
(QMatmul).  The wrappers stay PyTorch modules (north_star); their quantizers are the CUDA-backed
ones of llm_compressor_b200.quantizers.
"""
from copy import deepcopy

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .quantizers import FakeQuantizer


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class QLinear(nn.Linear):
    def __init__(self, linear: nn.Linear, quant_config, dtype, **kwargs):
        super().__init__(linear.in_features, linear.out_features, linear.bias is not None, linear.weight.device, dtype)
        op_name = kwargs.get("op_name", None)
        save_path = kwargs.get("save_path", "./")
        self.train(linear.training)
        with torch.no_grad():
            self.weight.copy_(linear.weight)
            if self.bias is not None:
                self.bias.copy_(linear.bias)
        self.input_quantizer = FakeQuantizer.build(_get(quant_config, "act_in"), op_name=f"{op_name}.input", save_path=save_path)
        self.weight_quantizer = FakeQuantizer.build(_get(quant_config, "weight"), op_name=f"{op_name}.weight", save_path=save_path)
        self.output_quantizer = FakeQuantizer.build(_get(quant_config, "act_out"), op_name=f"{op_name}.output", save_path=save_path)

    def forward(self, inputs: Tensor, **kwargs) -> Tensor:
        R1 = kwargs.get("R1", None)
        if R1 is not None:  # online rotation branch used by the SpinQuant training model (qlinear.py:59-84)
            dtype = self.weight.dtype
            transpose = kwargs.get("transpose", False)
            if not transpose:
                weight = self.weight.to(torch.float64) @ R1.to(torch.float64)
            else:
                weight = R1.T.to(torch.float64) @ self.weight.to(torch.float64)
            R2 = kwargs.get("R2", None)
            if R2 is not None:
                had_dim = R2.shape[0]
                if transpose:
                    init_shape = weight.shape
                    temp = weight.reshape(-1, init_shape[-1] // had_dim, had_dim)
                    weight = (temp.to(torch.float64) @ R2.to(torch.float64)).reshape(init_shape)
                else:
                    W_ = weight.t()
                    tshape = W_.shape
                    temp = W_.reshape(-1, tshape[-1] // had_dim, had_dim)
                    weight = (temp.to(torch.float64) @ R2.to(torch.float64)).reshape(tshape).t()
            self.weight.data = self.weight_quantizer(weight.data.to(dtype))
        return self.output_quantizer(F.linear(self.input_quantizer(inputs), self.weight, self.bias))

    def extra_repr(self):
        return f"(in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None})"


class QMatmul(nn.Module):
    def __init__(self, quant_config, axes=-1, **kwargs):
        super().__init__()
        op_name = kwargs.get("op_name", None)
        save_path = kwargs.get("save_path", "./")
        act_in = dict(_get(quant_config, "act_in"))
        act_in2 = deepcopy(act_in)
        self.input1_quantizer = FakeQuantizer.build(act_in, op_name=f"{op_name}.input1", save_path=save_path)
        act_in2["axes"] = axes
        if (axes == -1) and (act_in.get("group_size") == -2):
            act_in2["group_size"] = -1
        if (axes == -2) and (act_in.get("group_size") == -1):
            act_in2["group_size"] = -2
        self.input2_quantizer = FakeQuantizer.build(act_in2, op_name=f"{op_name}.input2", save_path=save_path)
        self.output_quantizer = FakeQuantizer.build(_get(quant_config, "act_out"), op_name=f"{op_name}.output", save_path=save_path)

    def forward(self, inputs1: Tensor, inputs2: Tensor) -> Tensor:
        return self.output_quantizer(
            torch.matmul(self.input1_quantizer(inputs1).to(inputs2), self.input2_quantizer(inputs2)))

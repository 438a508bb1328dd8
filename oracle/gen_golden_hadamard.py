"""TEST INFRASTRUCTURE ONLY -- golden vectors for the Hadamard / rotation stage.

Runs the UNMODIFIED reference (spinquant/hadamard_utils.py, rotation_utils.py, fuse_norm_utils.py through
oracle/ref_shim.py) on CPU and stores: the sign patterns of its fixed H_K tables (packed bits), matmul_hadU outputs
on seeded fp64 inputs, seeded random_hadamard_matrix results, and a fuse_layer_norms + rotate_model run on the tiny
Llama of gen_golden_drivers.py.

    python oracle/gen_golden_hadamard.py        (build container only)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import llm_compressor.quantization.calibrations.gptq.core  # noqa: F401,E402  (installs the reference's sys.path hack)
from llm_compressor.quantization.calibrations.spinquant import fuse_norm_utils as FN  # noqa: E402
from llm_compressor.quantization.calibrations.spinquant import hadamard_utils as HU  # noqa: E402
from llm_compressor.quantization.calibrations.spinquant import rotation_utils as RU  # noqa: E402

import gen_golden_drivers as GD  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    out = {}
    for K in (12, 20, 28, 36, 40, 44, 60, 52, 108, 140, 156, 172):
        H = getattr(HU, "get_had%d" % K)().numpy()
        assert set(np.unique(H)) == {-1.0, 1.0}
        out["had%d" % K] = np.packbits(H < 0)
    sizes = [16, 64, 96, 160, 224, 288, 352, 480, 2048, 2560, 3072, 208, 216, 280, 312, 344, 11008]
    for n in sizes:
        g = torch.Generator().manual_seed(n)
        X = torch.randn(3, n, generator=g, dtype=torch.float64)
        out["hadU_in_%d" % n] = X.numpy()
        out["hadU_out_%d" % n] = HU.matmul_hadU(X).numpy()
        out["hadUt_out_%d" % n] = HU.matmul_hadUt(X).numpy()
    for n in (64, 96, 160):
        torch.manual_seed(n)
        out["rand_had_%d" % n] = RU.random_hadamard_matrix(n, "cpu").numpy()
    # fuse_layer_norms + rotate_model on the tiny Llama (head_dim 32), CPU, seed 7
    m = GD.tiny_llama()
    for p in m.parameters():   # non-trivial norm weights so that the fusion is visible
        if p.dim() == 1:
            p.data = (1.0 + 0.1 * torch.randn(p.shape, generator=torch.Generator().manual_seed(p.numel()))).to(p.dtype)
    for k, v in m.state_dict().items():
        out["rot_init/" + k] = GD.bits(v) if v.dtype == torch.bfloat16 else v.numpy()
    m.get_layers = types.MethodType(lambda self: self.model.layers, m)
    FN.fuse_layer_norms(m)
    torch.manual_seed(7)
    HU_cuda = HU.apply_exact_had_to_linear

    def apply_cpu(module, had_dim=-1, output=False, R2=None):
        # the reference hard-codes .cuda() here (hadamard_utils.py:146,155); run the same arithmetic on the CPU
        W_ = module.weight.data
        dtype = W_.dtype
        W_ = W_.float()
        hadK = R2.to(torch.float64)
        if output:
            W_ = W_.t()
            shp = W_.shape
            W_ = (W_.reshape(-1, shp[-1] // had_dim, had_dim).to(torch.float64) @ hadK).reshape(shp).t()
        else:
            shp = W_.shape
            W_ = (W_.reshape(-1, shp[-1] // had_dim, had_dim).to(torch.float64) @ hadK).reshape(shp)
        module.weight.data = W_.to(dtype=dtype)

    RU.apply_exact_had_to_linear = apply_cpu
    RU.rotate_model(m, "hadamard", "cpu")
    RU.apply_exact_had_to_linear = HU_cuda
    for k, v in m.state_dict().items():
        out["rot_out/" + k] = GD.bits(v) if v.dtype == torch.bfloat16 else v.numpy()
    out["__meta__"] = np.array(repr(dict(sizes=sizes, rand=[64, 96, 160], ks=[12, 20, 28, 36, 40, 44, 60, 52, 108, 140, 156, 172], rot_seed=7)))
    np.savez_compressed(os.path.join(GOLD, "hadamard.npz"), **out)
    print("wrote hadamard.npz", os.path.getsize(os.path.join(GOLD, "hadamard.npz")))


if __name__ == "__main__":
    main()

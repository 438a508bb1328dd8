"""tcgen05 3xTF32 GEMM (csrc/tgemm.cu) against fp64: the building block of the lazy-batch update
(ref: gptq/core.py:265) and of the Cholesky / GPTAQ-P contractions on the tensor-core path."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from llm_compressor_b200 import ops
    return ops


@pytest.mark.parametrize("M,N,Kd", [(128, 256, 32), (128, 256, 128), (256, 512, 128), (3072, 2944, 128),
                                    (1000, 776, 100), (64, 40, 8), (1024, 7168, 1024), (130, 260, 36)])
def test_tgemm_store_vs_fp64(M, N, Kd):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(M + N + Kd)
    A = torch.randn(M, Kd, generator=g, device=DEV) * torch.exp(torch.randn(M, 1, generator=g, device=DEV))
    B = torch.randn(N, Kd, generator=g, device=DEV)
    C = ops.tgemm_nt(A, B, alpha=-0.5)
    ref = -0.5 * (A.double() @ B.double().T)
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    relf = float((C.double() - ref).norm() / ref.norm())
    ref32 = -0.5 * (A @ B.T)  # cuBLAS fp32 (may itself use a split path) for scale
    relf32 = float((ref32.double() - ref).norm() / ref.norm())
    print(f"{M}x{N}x{Kd}: max-rel={err:.2e} relF={relf:.2e} (torch fp32 relF={relf32:.2e})")
    # 3xTF32 split error ~2^-21 per product; the tensor core's fp32 accumulator truncates, which adds a
    # bias growing with the MMA chain length (3 * Kd / 8 instructions).  Single-pass TF32 would be ~5e-4.
    assert relf < 2e-6 * max(1.0, Kd / 128)
    assert err < 1e-5 * max(1.0, Kd / 128)


def test_tgemm_accumulate_strided():
    """C += alpha A B^T on sub-blocks of larger matrices (leading dimensions, reduce-add epilogue)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    big_a = torch.randn(700, 1024, generator=g, device=DEV)
    big_b = torch.randn(900, 1024, generator=g, device=DEV)
    big_c = torch.randn(700, 2048, generator=g, device=DEV)
    A = big_a[:, 128:256]
    B = big_b[4:, 512:640]
    Cv = big_c[:, 1024:1024 + 896]
    ref_all = big_c.clone().double()
    ref_all[:, 1024:1024 + 896] += 2.0 * (A.double() @ B.double().T)
    ops.tgemm_nt(A, B, Cv, alpha=2.0, accumulate=True)
    assert float((big_c.double() - ref_all).abs().max()) < 3e-6 * float(ref_all.abs().max())
    # untouched columns are bit-identical
    assert torch.equal(big_c[:, :1024].double(), ref_all[:, :1024])
    assert torch.equal(big_c[:, 1024 + 896:].double(), ref_all[:, 1024 + 896:])


@pytest.mark.parametrize("accumulate", [False, True])
def test_tgemm_chained_accumulation(accumulate):
    """kchain = 256: chains of 96 MMAs combined by fp32 reduce-adds -> error independent of Kd."""
    ops = _ops()
    M, N, Kd = 1024, 1536, 2048
    g = torch.Generator(device=DEV).manual_seed(11)
    A = torch.randn(M, Kd, generator=g, device=DEV).abs()   # same-sign products: worst case for truncation bias
    B = torch.randn(N, Kd, generator=g, device=DEV).abs()
    C0 = torch.randn(M, N, generator=g, device=DEV)
    ref = A.double() @ B.double().T + (C0.double() if accumulate else 0.0)
    C1 = C0.clone()
    ops.tgemm_nt(A, B, C1, accumulate=accumulate, kchain=0)
    C2 = C0.clone()
    ops.tgemm_nt(A, B, C2, accumulate=accumulate, kchain=256)
    e1 = float((C1.double() - ref).norm() / ref.norm())
    e2 = float((C2.double() - ref).norm() / ref.norm())
    print(f"one chain relF={e1:.2e}, 256-column chains relF={e2:.2e}")
    assert e2 < 3e-6 and e2 <= e1

"""Hadamard / rotation stage: generated H_K blocks and the CPU oracle against the reference's own
tables / outputs (tests/golden/hadamard.npz, oracle/gen_golden_hadamard.py); the CUDA transform (`lcb_hadamard_rows`
through llm_compressor_b200.hadamard) against the golden vectors, the oracle and the reference's dense fp64 GEMM form.

Tolerances: fp64 outputs 1e-13 relative (the reference sums n products in GEMM order, the transform in butterfly order);
bf16 / fp32 outputs equal to the fp64 result rounded once (checked exactly, up to ties that move under 1e-16 noise)."""
import numpy as np
import pytest
import torch

import golden_io as gio
import oracle as orc
from util import t_from_bits

Z, META = gio.load("hadamard")
DEV = "cuda:0"


@pytest.mark.parametrize("K", [12, 20, 28, 36, 40, 44, 60, 52, 108, 140, 156, 172])
def test_generated_hadK_equals_reference_table(K):
    from llm_compressor_b200 import hadamard as H
    neg = np.unpackbits(Z["had%d" % K])[: K * K].reshape(K, K).astype(bool)
    ref = np.where(neg, -1.0, 1.0).astype(np.float32)
    got = H._had(K).numpy()
    assert np.array_equal(got, ref)
    assert np.array_equal(got @ got.T, K * np.eye(K, dtype=np.float32))


def test_get_hadK_precedence_and_unsupported():
    from llm_compressor_b200 import hadamard as H
    assert H.get_hadK(2560)[1] == 40 and H.get_hadK(3072)[1] == 12 and H.get_hadK(10240)[1] == 40
    assert H.get_hadK(8192) == (None, 1) and H.get_hadK(14336)[1] == 28
    assert H.get_hadK(11008)[1] == 172 and H.get_hadK(13824)[1] == 108 and H.get_hadK(5120)[1] == 40   # Llama-2 7B / 13B ffn
    assert H.get_hadK(6656)[1] == 52 and H.get_hadK(17920)[1] == 140 and H.get_hadK(19968)[1] == 156   # Llama-1 30B
    t = H.get_hadK(96, transpose=True)[0]
    assert torch.equal(t, H.get_hadK(96)[0].T)


@pytest.mark.parametrize("n", META["sizes"])
def test_oracle_matmul_hadU_matches_reference(n):
    x = Z["hadU_in_%d" % n]
    for tr, key in ((False, "hadU_out_%d"), (True, "hadUt_out_%d")):
        ref = Z[key % n]
        got = orc.matmul_hadU(x, transpose=tr)
        assert np.max(np.abs(got - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.gpu
@pytest.mark.parametrize("n", META["sizes"])
def test_cuda_matmul_hadU_matches_reference_golden(n):
    from llm_compressor_b200 import hadamard as H
    x = torch.from_numpy(Z["hadU_in_%d" % n]).to(DEV)
    for tr, key in ((False, "hadU_out_%d"), (True, "hadUt_out_%d")):
        ref = Z[key % n]
        got = H.matmul_hadU(x, transpose=tr).cpu().numpy()
        assert got.dtype == np.float64
        assert np.max(np.abs(got - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref))), (n, tr)


@pytest.mark.gpu
@pytest.mark.parametrize("n", META["rand"])
def test_random_hadamard_matrix_same_seed_same_matrix(n):
    from llm_compressor_b200 import hadamard as H
    torch.manual_seed(n)
    got = H.random_hadamard_matrix(n, DEV).cpu().numpy()
    ref = Z["rand_had_%d" % n]
    assert np.max(np.abs(got - ref)) <= 1e-15
    assert np.max(np.abs(got @ got.T - np.eye(n))) < 1e-6  # the reference's fp32 sqrt(n) leaves ~3e-8


@pytest.mark.gpu
@pytest.mark.parametrize("rows,n", [(512, 2048), (300, 2560), (256, 3072), (64, 8192), (32, 10240), (1000, 128), (77, 64),
                                    (5, 12), (3, 16384), (64, 11008), (16, 13824), (9, 6656)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_cuda_rotation_vs_oracle_and_dense_gemm(rows, n, dtype):
    """W @ R (R = diag(s) matmul_hadU(I)) at model widths: fast transform == oracle == the reference's dense fp64 GEMM."""
    from llm_compressor_b200 import hadamard as H
    g = torch.Generator().manual_seed(rows + n)
    W = (0.02 * torch.randn(rows, n, generator=g)).to(dtype)
    s = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).double()
    rot = H.HadamardRotation(s, DEV)
    got = rot.right(W.to(DEV))
    assert got.dtype == dtype
    exact = orc.matmul_hadU(W.double().numpy(), signs=s.numpy())
    ref = torch.from_numpy(exact).to(dtype)
    neq = int((got.cpu() != ref).sum())
    assert neq <= max(1, rows * n // 100000), neq     # a tie may fall differently under 1e-16 noise; nothing else may
    if n <= 3072:  # the reference's own form: dense fp64 R and a GEMM (rotation_utils.py:57-63)
        R = rot.dense()
        dense = torch.matmul(W.to(DEV).double(), R).to(dtype)
        assert int((got != dense).sum()) <= max(1, rows * n // 100000)
        # R.T @ W down the columns
        gotl = rot.left_t(W.t().contiguous().to(DEV))
        densel = torch.matmul(R.T, W.t().to(DEV).double()).to(dtype)
        assert int((gotl != densel).sum()) <= max(1, rows * n // 100000)
    # fp32 accumulation variant stays within fp32 rounding of the exact result
    fast = H.hadamard_rows(W.to(DEV), rot.signs, out_dtype=torch.float32, acc64=False).cpu().double().numpy()
    assert np.max(np.abs(fast - exact)) <= 2e-6 * np.max(np.abs(exact))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,n", [(37, 1024), (300, 2048), (257, 3072), (1030, 3072), (131, 4096), (67, 8192), (600, 8192), (45, 2560), (777, 2560), (1000, 128), (77, 64), (50, 256), (33, 512)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("with_signs", [True, False])
def test_register_kernel_fp32_accumulation(rows, n, out_dtype, with_signs):
    """The register-resident kernel (bf16 rows of the common widths, acc64=False; hadamard.cu hadamard_reg_kernel):
    every output within fp32 accumulation error of the exact transform (2e-6 of the row scale for fp32 output, one bf16
    rounding for bf16 output), ragged row counts, in place, and nearly everywhere the same bf16 value as the fp64 path."""
    from llm_compressor_b200 import hadamard as H
    g = torch.Generator().manual_seed(rows * 7 + n)
    W = (0.02 * torch.randn(rows, n, generator=g) * torch.exp(torch.randn(rows, 1, generator=g))).to(torch.bfloat16)
    s = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).double() if with_signs else None
    exact = orc.matmul_hadU(W.double().numpy(), signs=None if s is None else s.numpy())
    sd = None if s is None else s.to(DEV)
    got = H.hadamard_rows(W.to(DEV), sd, out_dtype=out_dtype, acc64=False)
    assert got.dtype == out_dtype
    err = np.abs(got.double().cpu().numpy() - exact)
    scale = np.abs(exact).max(axis=1, keepdims=True) + 1e-30
    tol = 2e-6 if out_dtype == torch.float32 else 2.0 ** -8
    assert float((err / scale).max()) <= tol, float((err / scale).max())
    if out_dtype == torch.bfloat16:
        ref64 = H.hadamard_rows(W.to(DEV), sd, acc64=True)
        assert float((got != ref64).float().mean()) < 1e-3      # fp32 sums differ from fp64 sums only at rounding ties
        z = W.to(DEV).clone()
        H.hadamard_rows(z, sd, acc64=False, out=z)              # in place
        assert torch.equal(z, got)


@pytest.mark.gpu
def test_transform_is_orthogonal_and_in_place_at_full_size():
    from llm_compressor_b200 import hadamard as H
    for n in (2560, 3072, 8192):
        x = torch.randn(4096, n, device=DEV, dtype=torch.float32)
        y = H.matmul_hadU(x)
        back = H.matmul_hadUt(y)
        assert float((back - x).abs().max()) < 1e-5
        assert abs(float(y.double().pow(2).sum() / x.double().pow(2).sum()) - 1.0) < 1e-6
        z = x.clone()
        H.hadamard_rows(z, out=z)
        assert torch.equal(z, y)


@pytest.mark.gpu
def test_fuse_layer_norms_and_rotate_model_match_reference_on_tiny_llama():
    """fuse_layer_norms + rotate_model (R1 over the hidden size, one R2 per layer over head_dim) with the reference's RNG
    stream: every rotated weight equals the reference's fp64-GEMM result bit for bit (bf16)."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from llm_compressor_b200 import adapters
    from llm_compressor_b200 import hadamard as H
    DZ, DM = gio.load("drivers")
    cfg = LlamaConfig(vocab_size=DM["vocab"], hidden_size=DM["d"], intermediate_size=DM["ffn"], num_hidden_layers=DM["layers"],
                      num_attention_heads=DM["heads"], num_key_value_heads=DM["kv"], max_position_embeddings=DM["seqlen"],
                      tie_word_embeddings=False, attn_implementation="eager")
    m = LlamaForCausalLM(cfg).to(torch.bfloat16)
    sd = {k[len("rot_init/"):]: t_from_bits(Z[k]) for k in Z.files if k.startswith("rot_init/")}
    m.load_state_dict(sd, strict=False)
    adapters.attach_duck_type(m)
    H.fuse_layer_norms(m)
    torch.manual_seed(META["rot_seed"])
    H.rotate_model(m, "hadamard", DEV)
    bad = {}
    for k, v in m.state_dict().items():
        ref = t_from_bits(Z["rot_out/" + k]) if Z["rot_out/" + k].dtype == np.uint16 else torch.from_numpy(Z["rot_out/" + k])
        # bit-exact, except where the exact result is 0 (sums of +-w that cancel): the reference's fp64 GEMM leaves
        # ~1e-17 of rounding noise there, the transform returns the exact 0 (measured: 2 such elements of 330k)
        d = int(((v.cpu().float() != ref.float()) & ((v.cpu().float() - ref.float()).abs() > 1e-12)).sum())
        if d:
            bad[k] = d
    assert not bad, bad


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fwht_all_power_of_two_dims_like_the_reference_fwht_test(dtype):
    """Mirror of the only asserting test in the reference tree (third_party/fast-hadamard-transform/tests/
    test_fast_hadamard_transform.py:12-47: dims 1 .. 32768, batch 15): the transform against the exact fp64 result;
    tolerance = the output dtype's rounding (fp32: 2e-6 relative to the row scale, bf16: one bf16 ulp)."""
    from llm_compressor_b200 import hadamard as H
    for logn in range(0, 16):
        n = 1 << logn
        g = torch.Generator().manual_seed(n)
        x = torch.randn(15, n, generator=g).to(dtype)
        exact = orc.matmul_hadU(x.double().numpy())
        got = H.hadamard_rows(x.to(DEV), acc64=(n <= 16384)).double().cpu().numpy()
        scale = np.abs(exact).max() + 1e-30
        tol = 2e-6 if dtype == torch.float32 else 2.0 ** -8
        assert np.max(np.abs(got - exact)) <= tol * scale, (n, float(np.max(np.abs(got - exact)) / scale))

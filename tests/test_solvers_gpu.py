"""GPU parity of the calibration statistics (a), layer solvers (b) and masks (d) against the
reference's outputs (tests/golden) and the CPU oracle.

Contract (BASELINE.json north_star): masks bit-exact; Hessians / factors / updated weights within
the tolerances written next to each assertion; layer-output SQNR within 0.1 dB of the reference.
"""
import numpy as np
import pytest
import torch

import golden_io as gio
import oracle as orc
from util import t_from_bits, to_f32_np

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SOLV, SMETA = gio.load("solvers")


def _ops():
    from llm_compressor_b200 import ops
    return ops


@pytest.fixture(params=["tcgen05-3xtf32", "exact-fp32"])
def gemm_mode(request):
    """Solver GEMM precision (lcb_set_gemm_mode): the tensor-core path is the product default, the
    exact FFMA path is the anchor that reproduces the reference's fp32 results bit for bit."""
    ops = _ops()
    prev = ops.set_gemm_mode(1 if request.param.startswith("tcgen05") else 0)
    yield request.param
    ops.set_gemm_mode(prev)


def relf(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _accumulate(Xb, Xfpb=None, lazy=False):
    ops = _ops()
    X = t_from_bits(Xb, DEV)
    K = X.shape[-1]
    H = torch.zeros(K, K, device=DEV)
    D = torch.zeros(K, K, device=DEV) if Xfpb is not None else None
    Xfp = t_from_bits(Xfpb, DEV) if Xfpb is not None else None
    n = 0
    fn = ops.hessian_accum_raw if lazy else ops.hessian_accum
    for j in range(X.shape[0]):
        n = fn(H, X[j].unsqueeze(0), n, dxxt=D, x_fp=None if Xfp is None else Xfp[j])
    if lazy:
        ops.hessian_finalize(H, 2.0 / n, True)
        if D is not None:
            ops.hessian_finalize(D, 2.0 / n, False)
    return H, D


@pytest.mark.parametrize("lazy", [False, True], ids=["running-mean", "raw-sums-upper"])
def test_hessian_matches_reference_golden(lazy):
    H, D = _accumulate(SOLV["X"], SOLV["Xfp"], lazy)
    # tcgen05 accumulates the exact bf16 products in fp32 with truncation inside one MMA chain
    # (measured bias ~ -4e-6 relative on 2048-token sums); tolerance: relF 1e-5, max-abs 1e-5 of max
    assert relf(to_f32_np(H), SOLV["H"]) < 1e-5
    assert np.abs(to_f32_np(H) - SOLV["H"]).max() <= 1e-5 * np.abs(SOLV["H"]).max()
    assert relf(to_f32_np(D), SOLV["dXXT"]) < 5e-5       # dX carried as bf16 hi+lo (16 mantissa bits)
    assert np.allclose(to_f32_np(H), to_f32_np(H).T, rtol=0, atol=1e-6 * np.abs(SOLV["H"]).max())


@pytest.mark.parametrize("T,K", [(2048, 3072), (100, 328), (65, 8), (3000, 776), (2048, 8192)])
def test_hessian_upper_raw_vs_fp64(T, K):
    """raw-sum / symmetric-half form (stream-K split tiles, TMA reduce-add) against fp64."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(T * 7 + K)
    X = (torch.randn(T, K, generator=g, device=DEV) * torch.exp(0.5 * torch.randn(K, generator=g, device=DEV))).to(torch.bfloat16)
    H = torch.zeros(K, K, device=DEV)
    n = 0
    for _ in range(2):
        n = ops.hessian_accum_raw(H, X.unsqueeze(0), n)
    ops.hessian_finalize(H, 2.0 / n, True)
    ref = (2.0 * (X.double().T @ X.double())).float() if K <= 4096 else None
    if ref is None:
        sl = slice(5000, 5300)
        ref = 2.0 * (X[:, sl].double().T @ X.double())
        got = H[sl].double()
        assert float((got - ref).norm() / ref.norm()) < 1e-5
    else:
        assert float((H.double() - ref.double()).norm() / ref.double().norm()) < 1e-5
    assert torch.equal(H, H.T)


@pytest.mark.parametrize("T,K", [(2048, 2048), (2048, 3072), (100, 328), (4096, 1024), (65, 8), (3000, 776)])
def test_hessian_shapes_vs_oracle(T, K):
    ops = _ops()
    g = torch.Generator().manual_seed(T + K)
    X = (torch.randn(T, K, generator=g) * torch.exp(0.5 * torch.randn(K, generator=g))).to(torch.bfloat16)
    H = torch.full((K, K), 0.25, device=DEV)
    Href = np.full((K, K), 0.25, np.float32)
    n = ops.hessian_accum(H, X.to(DEV).unsqueeze(0), 3)
    assert n == 4
    orc.hessian_accum(Href, X.float().numpy(), 3)
    # truncating fp32 accumulation inside the tensor core: bias grows with the chain length T/16
    assert relf(to_f32_np(H), Href) < 1e-5 * max(1.0, T / 2048)


@pytest.mark.parametrize("K", [200, 1000, 3072, 8192])
def test_hessian_packed_upper_matches_finalize(K):
    """The token-sharded ranks' packed all-reduce form (lcb_hessian_pack_upper -> lcb_hessian_finalize_packed):
    bit for bit lcb_hessian_finalize(symmetric_from_upper) of the same raw sums, garbage below the diagonal tiles never
    read, layout as written in include/lcb200.h (checked against the numpy stand-in of the gloo tests)."""
    from test_parallel_cpu import _NumpyPacked
    from llm_compressor_b200 import parallel
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(K)
    X = torch.randn(512, K, generator=g).to(torch.bfloat16).to(DEV)
    H = torch.tril(torch.full((K, K), 9.0, device=DEV), -1)   # junk everywhere below the diagonal: must never be read
    ops.hessian_add(H, X, 1.0, 1.0, upper_only=True)
    packed = ops.hessian_pack_upper(H)
    nb = -(-K // 32)
    assert packed.numel() == nb * (nb + 1) // 2 * 1024
    if K <= 1000:
        assert np.array_equal(packed.cpu().numpy(), _NumpyPacked().hessian_pack_upper(H.cpu()).numpy())
    out = torch.full((K, K), 123.0, device=DEV)
    ops.hessian_finalize_packed(packed, out, 0.25)
    ref = ops.hessian_finalize(H.clone(), 0.25, True)
    assert torch.equal(out, ref) and torch.equal(out, out.t())
    # single process: parallel.reduce_finalize_hessian_ is the plain finalize
    H1 = H.clone()
    assert parallel.reduce_finalize_hessian_(H1, 8) == 8 and torch.equal(H1, ops.hessian_finalize(H.clone(), 2.0 / 8, True))


def test_hessian_linearity_full_size():
    """K = 8192 (down_proj of Llama-3.2-3B), 2 samples: H(X1) + H(X2) == H([X1; X2]) scaled."""
    ops = _ops()
    K, T = 8192, 2048
    g = torch.Generator(device=DEV).manual_seed(1)
    X = torch.randn(2 * T, K, generator=g, device=DEV).to(torch.bfloat16)
    Ha = torch.zeros(K, K, device=DEV)
    ops.hessian_add(Ha, X[:T].contiguous(), 1.0, 0.0)
    ops.hessian_add(Ha, X[T:].contiguous(), 1.0, 1.0)
    Hb = torch.zeros(K, K, device=DEV)
    ops.hessian_add(Hb, X, 1.0, 1.0, upper_only=True)
    ops.hessian_finalize(Hb, 1.0, True)
    assert torch.equal(Hb, Hb.T)
    assert float((Ha - Hb).norm() / Hb.norm()) < 5e-6
    # spot check one 128 x 256 tile against fp64
    ref = X[:, 4096:4224].double().T @ X[:, 512:768].double()
    assert float((Hb[4096:4224, 512:768].double() - ref).norm() / ref.norm()) < 1e-5


def test_rownorm_vs_oracle():
    ops = _ops()
    z, _ = gio.load("masks")
    X = t_from_bits(z["X"], DEV)
    s = torch.zeros(X.shape[-1], device=DEV)
    n = 0
    for j in range(X.shape[0]):
        n = ops.rownorm_accum(s, X[j].unsqueeze(0), n)
    np.testing.assert_allclose(to_f32_np(s), z["scaler_row"], rtol=3e-6)


def _spd(K, seed, T=None):
    g = torch.Generator().manual_seed(seed)
    T = T or 2 * K
    X = (torch.randn(T, K, generator=g) * torch.exp(0.7 * torch.randn(K, generator=g))).to(torch.bfloat16).float()
    return ((2.0 / T) * X.T @ X).contiguous()


@pytest.mark.parametrize("K", [128, 256, 384, 1000, 2048, 3072])
def test_chol_inv_upper(K, gemm_mode):
    ops = _ops()
    H = _spd(K, K)
    Hd = H.double() + 0.01 * torch.diag(H).double().mean() * torch.eye(K, dtype=torch.float64)
    U = ops.chol_inv_upper(H.to(DEV), percdamp=0.01).cpu().double()
    assert float(torch.tril(U, -1).abs().max()) == 0.0
    # U^T U (H + damp) == I
    resid = (U.T @ U @ Hd - torch.eye(K, dtype=torch.float64)).norm() / K ** 0.5
    Uref = torch.linalg.cholesky(torch.linalg.inv(Hd), upper=True)
    rel = float((U - Uref).norm() / Uref.norm())
    # reference chain in fp32 for scale: how far is IT from fp64?
    Uo = orc.damp_and_factor(H.numpy().copy(), 0.01).astype(np.float64)
    rel_ref = float(np.linalg.norm(Uo - Uref.numpy()) / np.linalg.norm(Uref.numpy()))
    print(f"K={K} resid={resid:.2e} rel={rel:.2e} reference-chain rel={rel_ref:.2e}")
    assert resid < 1e-3
    assert rel < max(5e-4, 3 * rel_ref)


def test_chol_inv_upper_headline_k8192(gemm_mode):
    """K = 8192 (down_proj of Llama-3.2-3B, 1 of the 4 factorisations per layer of the timed step): residual and relF
    against fp64 (computed on the GPU) next to the reference's own fp32 chain potrf -> potri -> potrf in PyTorch eager
    (ref: gptq/core.py:213-224)."""
    ops = _ops()
    K = 8192
    H = _spd(K, K).to(DEV)
    Hd = H.double() + 0.01 * torch.diag(H).double().mean() * torch.eye(K, dtype=torch.float64, device=DEV)
    U = ops.chol_inv_upper(H, percdamp=0.01).double()
    assert float(torch.tril(U, -1).abs().max()) == 0.0
    resid = float((U.T @ U @ Hd - torch.eye(K, dtype=torch.float64, device=DEV)).norm() / K ** 0.5)
    Uref = torch.linalg.cholesky(torch.linalg.inv(Hd), upper=True)
    rel = float((U - Uref).norm() / Uref.norm())
    Hf = H.clone()
    Hf.diagonal().add_(0.01 * torch.mean(torch.diag(H)))
    Ur = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hf)), upper=True).double()
    rel_ref = float((Ur - Uref).norm() / Uref.norm())
    print(f"K=8192 {gemm_mode}: resid={resid:.2e} relF={rel:.2e}  reference fp32 chain relF={rel_ref:.2e}")
    # measured on B200: tensor-core tile kernel 2.3e-6, exact path 5.5e-7, reference chain 8.7e-7
    assert resid < 2e-5
    assert rel < (5e-6 if gemm_mode.startswith("tcgen05") else 2e-6)


def test_chol_perm_and_not_spd_retry(gemm_mode):
    ops = _ops()
    K = 256
    H = _spd(K, 7)
    perm = torch.randperm(K, generator=torch.Generator().manual_seed(0))
    U = ops.chol_inv_upper(H.to(DEV), perm=perm.to(DEV), percdamp=0.01).cpu().double()
    Hp = H[perm][:, perm].double()
    Hp += 0.01 * torch.diag(H).double().mean() * torch.eye(K, dtype=torch.float64)
    assert float((U.T @ U @ Hp - torch.eye(K, dtype=torch.float64)).norm() / K ** 0.5) < 1e-3
    bad = -torch.eye(K)  # negative definite: never factorises -> error like torch.linalg.cholesky
    with pytest.raises(RuntimeError):
        ops.chol_inv_upper(bad.to(DEV), percdamp=0.01)


class _Lin(torch.nn.Linear):
    pass


def _layer(W, cfg):
    import llm_compressor_b200 as lc
    N, K = W.shape
    lin = _Lin(K, N, bias=False, dtype=W.dtype, device=DEV)
    lin.weight.data = W.clone().to(DEV)
    lin.weight_quantizer = lc.FakeQuantizer.build(cfg).to(DEV)
    return lin


def _sqnr_db(X, W, Wq):
    ref = X @ W.T
    return float(10 * np.log10((ref ** 2).sum() / max(((ref - X @ Wq.T) ** 2).sum(), 1e-30)))


@pytest.mark.parametrize("case", SMETA, ids=[m[0] for m in SMETA])
def test_solvers_vs_reference_golden(case, gemm_mode):
    from llm_compressor_b200 import solvers
    name, cfg, kind = case
    W = t_from_bits(SOLV[name + "/W"])
    ref = gio.bits_to_f32(SOLV[name + "/Wnew"])
    H = torch.from_numpy(SOLV["H"].copy()).to(DEV)
    if kind == "sparsegpt":
        lay = solvers.Wrapper(_Lin(W.shape[1], W.shape[0], bias=False, dtype=W.dtype, device=DEV), DEV)
        lay.module.weight.data = W.clone().to(DEV)
        lay.H = H
        solvers.prune_weight(lay, DEV, 0.5)
        got = to_f32_np(lay.module.weight.data)
        mask_diff = float(np.mean((got == 0) != (ref == 0)))
        print(name, "mask mismatch fraction", mask_diff)
        assert mask_diff <= (0.0 if gemm_mode == "exact-fp32" else 2e-3)
    else:
        lin = _layer(W, cfg)
        lin.weight_quantizer.H = H
        if kind == "gptaq":
            lin.weight_quantizer.dXXT = torch.from_numpy(SOLV["dXXT"].copy()).to(DEV)
            solvers.gptaq_update_weight(lin, DEV, actorder=True, alpha=0.25)
        else:
            solvers.update_weight(lin, DEV, actorder=True)
        got = to_f32_np(lin.weight.data)
    rel = relf(got, ref)
    frac = float(np.mean(got != ref))
    X = gio.bits_to_f32(SOLV["X"]).reshape(-1, W.shape[1]).astype(np.float64)
    W64 = W.float().numpy().astype(np.float64)
    d_sqnr = abs(_sqnr_db(X, W64, got.astype(np.float64)) - _sqnr_db(X, W64, ref.astype(np.float64)))
    print(f"{name}: relF={rel:.3e} changed={frac:.3e} max-abs={np.abs(got - ref).max():.3e} dSQNR={d_sqnr:.4f} dB")
    assert d_sqnr < 0.1            # layer-output SQNR within 0.1 dB of the reference
    if gemm_mode == "exact-fp32" and kind != "sparsegpt":
        assert frac == 0.0         # exact fp32 contractions: bit-identical to the reference's output
    elif gemm_mode == "exact-fp32":
        assert rel < 1e-4          # SparseGPT: identical mask (asserted above), values at fp32 rounding level
    else:
        # 3xTF32 contractions + two-level lazy batch: a handful of one-step code flips (~1e-5
        # at 2048^2; these cases have 1.3e5..2.6e5 elements)
        assert frac < 2e-3 and rel < 2e-2


@pytest.mark.parametrize("N,K,cfgname", [(1024, 2048, "int4_g128"), (512, 3072, "int4_row"), (768, 1024, "mxfp4_g32"),
                                          (256, 1024, "nvfp4_g16"), (512, 1024, "gptaq_int4_g128")])
def test_gptq_vs_oracle_seeded(N, K, cfgname, gemm_mode):
    from llm_compressor_b200 import solvers
    cfgs = {
        "int4_g128": dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False),
        "int4_row": dict(type="int", format="int4", group_size=-1, axes=-1, zero_point=True, is_profile=False),
        "mxfp4_g32": dict(type="mx", format="fp4_e2m1", group_size=32, axes=-1, zero_point=False, is_profile=False),
        "nvfp4_g16": dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False, is_profile=False),
    }
    gptaq = cfgname.startswith("gptaq_")
    cfg = cfgs[cfgname[6:] if gptaq else cfgname]
    g = torch.Generator().manual_seed(N + K)
    W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16)
    T = 1024
    chan = torch.exp(0.8 * torch.randn(K, generator=g))
    X = (torch.randn(T, K, generator=g) * chan).to(torch.bfloat16)
    Xfp = (X.float() + 0.05 * torch.randn(T, K, generator=g) * chan).to(torch.bfloat16)
    H = np.zeros((K, K), np.float32); D = np.zeros((K, K), np.float32)
    orc.hessian_accum(H, X.float().numpy(), 0, dXXT=D, x_fp=Xfp.float().numpy())
    ref = orc.gptq_update(W.float().numpy(), H.copy(), cfg, dXXT=D.copy() if gptaq else None)
    lin = _layer(W, cfg)
    lin.weight_quantizer.H = torch.from_numpy(H.copy()).to(DEV)
    if gptaq:
        lin.weight_quantizer.dXXT = torch.from_numpy(D.copy()).to(DEV)
        solvers.gptaq_update_weight(lin, DEV, actorder=True, alpha=0.25)
    else:
        solvers.update_weight(lin, DEV, actorder=True)
    got = to_f32_np(lin.weight.data)
    X64 = X.float().numpy().astype(np.float64); W64 = W.float().numpy().astype(np.float64)
    d_sqnr = abs(_sqnr_db(X64, W64, got.astype(np.float64)) - _sqnr_db(X64, W64, ref.astype(np.float64)))
    frac = float(np.mean(got != ref))
    print(f"{cfgname} {N}x{K}: relF={relf(got, ref):.3e} changed={frac:.3e} dSQNR={d_sqnr:.4f} dB")
    assert d_sqnr < 0.1
    assert frac < (5e-3 if gemm_mode == "exact-fp32" else 1e-2)


_HEADLINE_ORACLE = {}


def _headline_case(N, K):
    """Seeded problem at a shape the benchmark times: bf16 weights, H = (2/T) X^T X of 2 K bf16 tokens (fp32 GEMM on the
    GPU; H is an INPUT to both sides), oracle result cached across the two GEMM modes."""
    key = (N, K)
    if key not in _HEADLINE_ORACLE:
        cfg = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)
        g = torch.Generator().manual_seed(N * 7 + K)
        W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16)
        T = 2 * K
        chan = torch.exp(0.8 * torch.randn(K, generator=g))
        X = (torch.randn(T, K, generator=g) * chan).to(torch.bfloat16)
        Xd = X.to(DEV).float()
        H = ((2.0 / T) * (Xd.T @ Xd)).contiguous()
        Hn = H.cpu().numpy()
        ref = orc.gptq_update(W.float().numpy(), Hn.copy(), cfg)
        _HEADLINE_ORACLE[key] = (cfg, W, X[:4096].float().numpy().astype(np.float64), Hn, ref)
    return _HEADLINE_ORACLE[key]


@pytest.mark.parametrize("N,K", [(3072, 8192), (8192, 3072), (5120, 3072), (16384, 3072)],
                         ids=["down_proj_3072x8192", "gate_proj_8192x3072", "qkv_stacked_5120x3072", "gate_up_stacked_16384x3072"])
def test_gptq_headline_shapes_vs_oracle(N, K, gemm_mode):
    """The solves one timed step of bench.py consists of (Llama-3.2-3B, int4-g[128]-rw, act-order) against the CPU oracle
    (= the reference, test_oracle_golden.py): max-abs, relative Frobenius error, changed-weight fraction and layer-output
    SQNR, as north_star words the contract.  K = 8192 takes the 64-column-tile Cholesky, the 1024-wide super-block lazy
    updates; the stacked shapes are what solvers.update_weights_shared hands to one solve."""
    from llm_compressor_b200 import solvers
    if gemm_mode == "exact-fp32" and (N, K) != (3072, 8192):
        pytest.skip("exact-fp32 anchor mode is run at the K = 8192 shape only (CPU oracle time)")
    cfg, W, X64, Hn, ref = _headline_case(N, K)
    lin = _layer(W, cfg)
    lin.weight_quantizer.H = torch.from_numpy(Hn.copy()).to(DEV)
    solvers.update_weight(lin, DEV, actorder=True)
    got = to_f32_np(lin.weight.data)
    W64 = W.float().numpy().astype(np.float64)
    s_ref, s_got = _sqnr_db(X64, W64, ref.astype(np.float64)), _sqnr_db(X64, W64, got.astype(np.float64))
    frac = float(np.mean(got != ref))
    mabs = float(np.abs(got - ref).max())
    step = float(np.abs(W.float().numpy()).max()) / 7.0
    print(f"GPTQ int4-g128 {N}x{K} {gemm_mode}: max-abs={mabs:.3e} (one step <= {step:.3e}) relF={relf(got, ref):.3e} "
          f"changed={frac:.3e} SQNR ref {s_ref:.3f} dB, ours {s_got:.3f} dB")
    assert abs(s_ref - s_got) < 0.1          # north_star: layer-output SQNR within 0.1 dB
    assert mabs <= 2.0 * step                # differing weights differ by a quantisation step, never more
    assert frac < (2e-3 if gemm_mode == "exact-fp32" else 5e-3)
    assert relf(got, ref) < 3e-2


def test_shared_factor_with_mixed_precision_group(gemm_mode):
    """Linears that share a Hessian may carry different quantizers (per-layer mixed precision: int4 q_proj,
    int8 k_proj, int4 v_proj at the same group size, ref: utils/parser.py register_4_to_8bit_config).  The shared-factor
    path must quantise every Linear with ITS quantizer: identical to solving them one by one."""
    from llm_compressor_b200 import solvers
    K, T = 1024, 2048
    g = torch.Generator().manual_seed(5)
    X = (torch.randn(T, K, generator=g) * torch.exp(0.8 * torch.randn(K, generator=g))).to(torch.bfloat16).to(DEV).float()
    H = ((2.0 / T) * X.T @ X).contiguous()
    fmts = [("int4", 384), ("int8", 128), ("int4", 256)]
    Ws = [(0.02 * torch.randn(n, K, generator=g)).to(torch.bfloat16) for _, n in fmts]

    def layers():
        return [_layer(W, dict(type="int", format=f, group_size=128, axes=-1, zero_point=False, is_profile=False))
                for (f, _), W in zip(fmts, Ws)]

    one_by_one = layers()
    for lin in one_by_one:
        lin.weight_quantizer.H = H.clone()
        solvers.update_weight(lin, DEV, actorder=True)
    shared = layers()
    fac = solvers.factorize(H.clone(), 128, actorder=True, percdamp=0.01)
    by_q = {}
    for lin in shared:
        by_q.setdefault(solvers.quantizer_key(lin.weight_quantizer), []).append(lin)
    assert len(by_q) == 2
    for same in by_q.values():
        solvers.update_weights_shared(same, DEV, fac)
    solvers.update_weights_shared(layers(), DEV, fac)   # a mixed list must not be stacked either (falls back per layer)
    for a, b, (f, _) in zip(one_by_one, shared, fmts):
        same = float((a.weight.data == b.weight.data).float().mean())
        nvals = int(torch.unique((a.weight.data.float() / a.weight.data.float().abs().amax(1, keepdim=True))).numel())
        print(f"{f}: identical to the separate solve {same:.6f}, distinct normalised levels {nvals}")
        assert same > (0.999 if gemm_mode.startswith("tcgen05") else 0.99999)   # tensor-core mode: run-to-run last-bit order
    # the int8 Linear really was quantised with 8 bits (an int4 solve has at most 15 levels per group)
    w8 = shared[1].weight.data.float()[:, :128]
    assert int(torch.unique(w8[0]).numel()) > 16


def test_sparsegpt_vs_oracle_seeded(gemm_mode):
    from llm_compressor_b200 import solvers
    N, K, T = 768, 1024, 1024
    g = torch.Generator().manual_seed(77)
    W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16)
    X = (torch.randn(T, K, generator=g) * torch.exp(0.8 * torch.randn(K, generator=g))).to(torch.bfloat16)
    H = np.zeros((K, K), np.float32)
    orc.hessian_accum(H, X.float().numpy(), 0)
    ref = orc.sparsegpt_prune(W.float().numpy(), H.copy(), 0.5)
    lay = solvers.Wrapper(_Lin(K, N, bias=False, dtype=W.dtype, device=DEV), DEV)
    lay.module.weight.data = W.clone().to(DEV)
    lay.H = torch.from_numpy(H.copy()).to(DEV)
    solvers.prune_weight(lay, DEV, 0.5)
    got = to_f32_np(lay.module.weight.data)
    sp = float(np.mean(got == 0))
    mm = float(np.mean((got == 0) != (ref == 0)))
    print(f"sparsity={sp:.4f} mask mismatch={mm:.3e} relF={relf(got, ref):.3e}")
    assert abs(sp - 0.5) < 2e-3
    assert mm < 5e-3


def test_masks_bit_exact_vs_reference_golden():
    ops = _ops()
    z, _ = gio.load("masks")
    W = t_from_bits(z["W"], DEV)
    s = torch.from_numpy(z["scaler_row"].copy()).to(DEV)
    for tag, ratio in (("30", 0.3), ("50", 0.5)):
        assert np.array_equal(ops.mask_wanda(W, s, ratio).cpu().numpy(), z["wanda_" + tag])
        assert np.array_equal(ops.mask_magnitude(W, ratio).cpu().numpy(), z["magnitude_" + tag])
        for alpha in (0.5, 1.0):
            got = ops.mask_ria(W, s, ratio, alpha).cpu().numpy()
            assert np.array_equal(got, z[f"ria_{tag}_{alpha}"]), (tag, alpha, int((got != z[f"ria_{tag}_{alpha}"]).sum()))


@pytest.mark.parametrize("N,K", [(512, 3072), (300, 8192), (64, 1000)])
def test_masks_vs_oracle_seeded(N, K):
    ops = _ops()
    g = torch.Generator().manual_seed(N * 3 + K)
    W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16)
    W[:, 3] = 0
    s = (torch.rand(K, generator=g) * 4 + 0.01)
    for ratio in (0.5, 0.3):
        ref = orc.mask_wanda(W.float().numpy(), s.numpy(), ratio)
        got = ops.mask_wanda(W.to(DEV), s.to(DEV), ratio).cpu().numpy()
        assert np.array_equal(got, ref)
        assert np.all(got.sum(1) == int(K * ratio))       # exactly k per row, ties by index
        ref, _ = orc.mask_magnitude(W.float().numpy(), ratio)
        assert np.array_equal(ops.mask_magnitude(W.to(DEV), ratio).cpu().numpy(), ref)
    Wd = W.clone().to(DEV)
    m = ops.mask_wanda(Wd, s.to(DEV), 0.5)
    ops.apply_mask(Wd, m)
    assert float((Wd == 0).float().mean()) >= 0.5


def test_stacked_solve_equals_per_linear(gemm_mode):
    """q/k/v share H: solving them as one stacked [sum N, K] problem (solvers.update_weights_shared) gives
    the same weights as the reference's one-Linear-at-a-time loop (ref: gptq/core.py:129-137)."""
    from llm_compressor_b200 import solvers
    K, Ns, T = 1024, (512, 128, 128), 1024
    cfg = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)
    g = torch.Generator().manual_seed(9)
    X = (torch.randn(T, K, generator=g) * torch.exp(0.8 * torch.randn(K, generator=g))).to(torch.bfloat16)
    H = torch.zeros(K, K, device=DEV)
    _ops().hessian_add(H, X.to(DEV), 2.0 / 1, 0.0)
    Ws = [(0.02 * torch.randn(n, K, generator=g)).to(torch.bfloat16) for n in Ns]
    fac = solvers.factorize(H.clone(), 128, True, 0.01)
    sep = []
    for W in Ws:
        lin = _layer(W, cfg)
        solvers.update_weight(lin, DEV, actorder=True, factor=fac)
        sep.append(lin.weight.data.clone())
    lins = [_layer(W, cfg) for W in Ws]
    solvers.update_weights_shared(lins, DEV, fac)
    for a, l in zip(sep, lins):
        assert torch.equal(a, l.weight.data)


# ------------------------------------------------------------------ row-sharded thresholds
def _two_shard_kth(ops, parts, kth):
    """lcb_select_* phases with the all-reduce of the two 'ranks' done by hand on one GPU."""
    states = [ops.select_state(DEV) for _ in parts]
    th = [torch.zeros(1, device=DEV) for _ in parts]
    for st, _ in states:
        ops.select_init(st, kth)
    for p in range(4):
        for (st, _), sc in zip(states, parts):
            ops.select_hist(st, sc.reshape(-1), p)
        total = sum(h for _, h in states)
        for (st, h), t in zip(states, th):
            h.copy_(total)
            ops.select_scan(st, p, t)
    assert all(torch.equal(th[0], t) for t in th)
    return th[0]


@pytest.mark.parametrize("N,K,split", [(512, 3072, 256), (300, 1000, 7), (64, 128, 64)])
def test_phase_api_row_shards_equal_unsharded_masks(N, K, split):
    ops = _ops()
    g = torch.Generator().manual_seed(N + K)
    W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16).to(DEV)
    s = (torch.rand(K, generator=g) * 4 + 0.01).to(DEV)
    shards = [W[:split].contiguous(), W[split:].contiguous()]
    for ratio in (0.5, 0.3):
        kth = min(int(N * K * ratio), N * K - 1)
        # magnitude: bit-exact
        scores = [ops.metric_magnitude(w) for w in shards]
        th = _two_shard_kth(ops, scores, kth)
        got = torch.cat([ops.mask_le(sc, th) for sc in scores], 0)
        assert torch.equal(got, ops.mask_magnitude(W, ratio))
        assert float(th) == float(np.sort(np.abs(W.float().cpu().numpy()).reshape(-1))[kth])
        # RIA: column sums re-associated (partial sums per shard) -> identical unless a bf16 rounding boundary moves
        sums = [ops.ria_sums(w) for w in shards]
        col = sums[0][0] + sums[1][0]
        scores = [ops.ria_metric(w, col, rs, s, 0.5) for w, (_, rs) in zip(shards, sums)]
        th = _two_shard_kth(ops, scores, kth)
        got = torch.cat([ops.mask_le(sc, th) for sc in scores], 0)
        agree = float((got == ops.mask_ria(W, s, ratio, 0.5)).float().mean())
        assert agree > 0.9995, agree


def test_sparsegpt_sharded_entry_world1_is_identical():
    """lcb_sparsegpt_update_sharded with an identity reduce and n_total == n must reproduce lcb_sparsegpt_update."""
    ops = _ops()
    N, K = 256, 512
    g = torch.Generator().manual_seed(5)
    W = (0.02 * torch.randn(N, K, generator=g)).to(DEV)
    X = torch.randn(2048, K, generator=g).to(DEV)
    H = (X.T @ X) / 1024
    U = ops.chol_inv_upper(H, percdamp=0.01)
    a = ops.sparsegpt_update(W.clone(), U, 0.5)
    calls = []
    b = ops.sparsegpt_update(W.clone(), U, 0.5, n_total=N, reduce=lambda h: calls.append(int(h.sum())))
    assert torch.equal(a, b)
    assert len(calls) == 4 * (K // 128) and calls[0] == N * 128
    with pytest.raises(RuntimeError):
        ops.sparsegpt_update(W.clone(), U, 0.5, n_total=N, reduce=lambda h: (_ for _ in ()).throw(RuntimeError("boom")))


def test_multi_gpu_sharded_pruning_script():
    """2-GPU NCCL run of tests/mgpu_sharded.py (row-sharded SparseGPT / RIA / magnitude == unsharded); needs 2 GPUs."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(here, "mgpu_sharded.py")],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "mgpu sharded ok" in r.stdout


@pytest.mark.parametrize("defer", [1, 2, 4, 8])
def test_hessian_accumulator_deferred_launches(defer):
    """ops.HessianAccumulator: `defer` hook inputs per launch give the same H as one launch per call (5 samples with
    different token counts: a shape change and a partial last batch force early launches)."""
    ops = _ops()
    K = 1024
    g = torch.Generator(device=DEV).manual_seed(3)
    Ts = [512, 512, 768, 512, 100]
    Xs = [torch.randn(1, t, K, generator=g, device=DEV).to(torch.bfloat16) for t in Ts]
    H = torch.zeros(K, K, device=DEV)
    acc = ops.HessianAccumulator(H, defer)
    for x in Xs:
        acc.add(x)
    n = acc.flush()
    assert n == len(Ts) and acc.flush() == n
    ops.hessian_finalize(H, 2.0 / n, True)
    Xall = torch.cat([x[0] for x in Xs], 0).double()
    ref = (2.0 / n) * (Xall.T @ Xall)
    assert float((H.double() - ref).norm() / ref.norm()) < 1e-5
    assert torch.equal(H, H.T)


@pytest.mark.parametrize("K,defer", [(3072, 1), (3072, 4), (8192, 4)])
def test_hessian_truncation_error_is_a_uniform_scale(K, defer):
    """What the truncating fp32 accumulation of the tensor core does to H (DESIGN.md deviations 1 / 9): most of the error is
    ONE common factor (every element loses a similar relative amount per accumulated token), and GPTQ / GPTAQ / SparseGPT
    are invariant to a scale of H (U scales by 1 / sqrt(a), the error term (w - q) / U_ii * U_i: does not).  Measured on
    B200: one 2048-token chain relF 4.7e-6, of which 1.5e-6 is left after removing the best-fit factor (-4.5e-6); four
    deferred inputs (8192-token chains) 2.9e-5 -> 1.5e-5 (K = 3072), 2.9e-5 -> 1.7e-5 (K = 8192)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(K + defer)
    Xs = [(torch.randn(2048, K, generator=g, device=DEV) * torch.exp(0.5 * torch.randn(K, generator=g, device=DEV))).to(torch.bfloat16)
          for _ in range(4)]
    H = torch.zeros(K, K, device=DEV)
    acc = ops.HessianAccumulator(H, defer)
    for x in Xs:
        acc.add(x.unsqueeze(0))
    n = acc.flush()
    ops.hessian_finalize(H, 2.0 / n, True)
    rows = slice(1024, 1536)                              # a row band is enough (and cheap in fp64 at K = 8192)
    X = torch.cat(Xs, 0).double()
    ref = (2.0 / n) * (X[:, rows].T @ X)
    got = H[rows].double()
    a = float((got * ref).sum() / (ref * ref).sum())      # best-fit common factor
    raw = float((got - ref).norm() / ref.norm())
    res = float((got / a - ref).norm() / ref.norm())
    print("K=%d defer=%d: relF %.2e, common factor 1 %+.2e, relF after removing it %.2e" % (K, defer, raw, a - 1.0, res))
    assert raw < 1e-5 * max(1, defer) and abs(a - 1.0) < 1e-5 * max(1, defer)
    assert a < 1.0 and res < 0.7 * raw


@pytest.mark.parametrize("K", [2048, 8192])
def test_hessian_multi_sample_chain_accuracy(K):
    """4 x 2048 tokens in one launch (one tensor-core accumulation chain of 8192 tokens per tile) against four launches.
    Same-sign products are the worst case for the truncation bias of the tensor core's fp32 accumulator: measured relF
    1.06e-5 (one 8192-token chain) vs 1.6e-6 (2048-token chains) at K = 2048.  The product default defers 4 hook inputs;
    test_deferred_hessian_chains_do_not_move_gptq_results measures what that does to the GPTQ result (nothing visible)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(4)
    Xs = [torch.randn(2048, K, generator=g, device=DEV).abs().to(torch.bfloat16) for _ in range(4)]
    H1 = torch.zeros(K, K, device=DEV)
    acc = ops.HessianAccumulator(H1, 4)
    for x in Xs:
        acc.add(x)
    acc.flush()
    H4 = torch.zeros(K, K, device=DEV)
    for x in Xs:
        ops.hessian_add(H4, x, 1.0, 1.0, upper_only=True)
    X = torch.cat(Xs, 0)
    ref = torch.triu(X[:, :256].double().T @ X.double())
    e1 = float((torch.triu(H1)[:256].double() - ref).norm() / ref.norm())
    e4 = float((torch.triu(H4)[:256].double() - ref).norm() / ref.norm())
    print(f"K={K}: one launch relF={e1:.2e}, four launches relF={e4:.2e}")
    assert e1 < 5e-5 and e4 < 1e-5  # measured: K=2048 1.06e-5 / 1.6e-6, K=8192 (CTA-pair kernel) 3.5e-5 / 5.6e-6


@pytest.mark.parametrize("K", [2048, 8192])
def test_deferred_hessian_chains_do_not_move_gptq_results(K):
    """End-to-end effect of the longer tensor-core accumulation chains of deferred launches: GPTQ int4-g128 with H from
    8 x 2048 calibration tokens accumulated 1 / 4 / 8 hook inputs per launch (2048 / 8192 / 16384-token chains) -- the
    layer-output SQNR must agree within 0.01 dB (contract: 0.1 dB) and the changed weights stay at the level every other
    tolerance-level perturbation of the solver produces."""
    from llm_compressor_b200 import solvers
    ops = _ops()
    N = 512
    cfg = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)
    g = torch.Generator().manual_seed(K)
    W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16)
    chan = torch.exp(0.8 * torch.randn(K, generator=g))
    Xs = [(torch.randn(2048, K, generator=g) * chan).to(torch.bfloat16).to(DEV) for _ in range(8)]
    X64 = torch.cat(Xs[:2], 0).double()
    outs, hs = {}, {}
    for defer in (1, 4, 8):
        lin = _layer(W, cfg)
        lin.weight_quantizer.H = torch.zeros(K, K, device=DEV)
        lin.weight_quantizer.nsamples = 0
        acc = ops.HessianAccumulator(lin.weight_quantizer.H, defer)
        for x in Xs:
            acc.add(x.unsqueeze(0))
        lin.weight_quantizer.nsamples = acc.flush()
        lin.weight_quantizer._h_raw = True
        hs[defer] = solvers.finalize_hessian(lin.weight_quantizer).clone()
        solvers.update_weight(lin, DEV, actorder=True)
        outs[defer] = lin.weight.data.double()
    Wd = W.to(DEV).double()
    ref_out = X64 @ Wd.T

    def sqnr(q):
        return float(10 * torch.log10(ref_out.pow(2).sum() / (ref_out - X64 @ q.T).pow(2).sum()))

    for defer in (4, 8):
        relh = float((hs[defer] - hs[1]).norm() / hs[1].norm())
        frac = float((outs[defer] != outs[1]).double().mean())
        ds = abs(sqnr(outs[defer]) - sqnr(outs[1]))
        print(f"K={K} defer={defer}: relF(H)={relh:.2e} changed weights={frac:.2e} dSQNR={ds:.4f} dB")
        # measured on B200: K=2048 defer 4 / 8: relF(H) 7.8e-6 / 2.0e-5, changed 9.5e-6 / 1.1e-5, dSQNR 0.0000 dB;
        #                   K=8192 defer 4 / 8: relF(H) 2.1e-5 / 5.4e-5, changed 2.3e-3 / 6.0e-3, dSQNR 0.0006 / 0.0004 dB
        assert relh < 1e-4 and ds < 0.01 and frac < 1e-2

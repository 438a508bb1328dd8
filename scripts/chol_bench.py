"""Cholesky-inverse stage on one B200: time + accuracy of lcb_chol_inv_upper next to the reference chain in PyTorch
eager (cuSOLVER potrf -> potri -> potrf, ref: gptq/core.py:213-224).  Development aid / source of profiles/*chol*.json.
    python scripts/chol_bench.py [K ...]      (LCB_CHOL_TILES=0 selects the old panel-launch chain)"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_compressor_b200 import _lib, ops

dev = torch.device("cuda:0")


def spd(K, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    T = 2 * K
    X = (torch.randn(T, K, generator=g, device=dev) * torch.exp(0.7 * torch.randn(K, generator=g, device=dev)))
    X = X.to(torch.bfloat16).float()
    return ((2.0 / T) * X.T @ X).contiguous()


def timeit(fn, n=5, warm=2):
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


def raw_call(H, U, ws, status, perm=None):
    L = _lib.lib()
    k = H.shape[0]
    rc = L.lcb_chol_inv_upper(H.data_ptr(), U.data_ptr(), k, None if perm is None else perm.data_ptr(), 0.01, ws.data_ptr(),
                              ws.numel(), status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.lib().lcb_last_error()


out = {"tiles_mode": os.environ.get("LCB_CHOL_TILES", "1")}
Ks = [int(a) for a in sys.argv[1:]] or [1000, 2048, 3072, 4096, 8192]
for K in Ks:
    H = spd(K, K)
    U = torch.empty_like(H)
    ws = torch.empty(_lib.lib().lcb_chol_ws_bytes(K), dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    r = {}
    for mode in (1, 0):
        ops.set_gemm_mode(mode)
        raw_call(H, U, ws, status)
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        Hd = H.double() + 0.01 * torch.diag(H).double().mean() * torch.eye(K, dtype=torch.float64, device=dev)
        Uref = torch.linalg.cholesky(torch.linalg.inv(Hd), upper=True)
        Ud = U.double()
        r["relF_vs_fp64_mode%d" % mode] = float((Ud - Uref).norm() / Uref.norm())
        r["resid_mode%d" % mode] = float((Ud.T @ Ud @ Hd - torch.eye(K, dtype=torch.float64, device=dev)).norm() / K ** 0.5)
        r["lower_zero_mode%d" % mode] = float(torch.tril(U, -1).abs().max())
        r["ms_mode%d" % mode] = timeit(lambda: raw_call(H, U, ws, status))
        # run-to-run reproducibility
        U2 = torch.empty_like(U)
        raw_call(H, U2, ws, status)
        r["bitrepro_mode%d" % mode] = bool(torch.equal(U, U2))
    ops.set_gemm_mode(1)

    def ref_chain():
        Hf = H.clone()
        d = torch.arange(K, device=dev)
        Hf[d, d] += 0.01 * torch.mean(torch.diag(Hf))
        L1 = torch.linalg.cholesky(Hf)
        Hi = torch.cholesky_inverse(L1)
        return torch.linalg.cholesky(Hi, upper=True)
    Ur = ref_chain().double()
    r["relF_reference_chain_vs_fp64"] = float((Ur - Uref).norm() / Uref.norm())
    r["ms_reference_eager_cusolver_chain"] = timeit(ref_chain, n=3, warm=1)
    out["K%d" % K] = r
    del H, U, ws, Hd, Uref, Ud, Ur
    torch.cuda.empty_cache()
print(json.dumps(out))

"""world_size-2 gloo tests (CPU) of the multi-GPU partitioning logic (llm_compressor_b200/parallel.py).

The numeric stages need a GPU, so here the CPU oracle plays their role: what is checked is that
sample-sharded raw Hessian sums + all-reduce + one finalize equal the reference's running mean, and
that row-sharded solves + all-gather equal the unsharded solve bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = globals()[fn](rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _inputs():
    import oracle as orc
    rng = np.random.default_rng(3)
    n_samples, T, K, N = 5, 96, 256, 96     # odd sample count: unbalanced shards
    X = orc.bf16_round(rng.standard_normal((n_samples, T, K), dtype=np.float32) * np.exp(0.5 * rng.standard_normal(K)).astype(np.float32))
    W = orc.bf16_round(0.02 * rng.standard_normal((N, K), dtype=np.float32))
    return orc, X, W


def _hessian_case(rank, world):
    from llm_compressor_b200 import parallel
    orc, X, W = _inputs()
    n_samples, T, K = X.shape
    mine = parallel.sample_shard(n_samples)
    S = np.zeros((K, K), np.float64)
    for j in mine:  # raw sums, as the lazy hook accumulates them
        S += X[j].astype(np.float64).T @ X[j].astype(np.float64)
    H = torch.from_numpy(S.astype(np.float32))
    H2 = H.clone()
    n_tot = parallel.reduce_hessian_(H, len(mine))
    # a statically split calibration set: the total is known, no count all-reduce / read-back; same sums
    assert parallel.reduce_hessian_(H2, len(mine), n_total=n_samples) == n_tot and torch.equal(H, H2)
    H = H.numpy() * np.float32(2.0 / n_tot)
    Href = np.zeros((K, K), np.float32)
    n = 0
    for j in range(n_samples):  # the reference's running mean over ALL samples
        n = orc.hessian_accum(Href, X[j], n)
    rel = float(np.linalg.norm(H - Href) / np.linalg.norm(Href))
    return dict(n_tot=n_tot, mine=mine, rel=rel, H=H)


class _NumpyPacked:
    """CPU stand-in for lcb_hessian_pack_upper / lcb_hessian_finalize_packed / lcb_hessian_finalize (csrc/hessian.cu; layout in
    include/lcb200.h), so that parallel.reduce_finalize_hessian_ runs under gloo without a GPU.  TEST CODE ONLY; the same class
    is the checker of the CUDA kernels in tests/test_solvers_gpu.py::test_hessian_packed_upper_matches_finalize."""

    @staticmethod
    def offsets(k):
        nb = -(-k // 32)
        return nb, {(bi, bj): 1024 * (bi * nb - bi * (bi - 1) // 2 + bj - bi) for bi in range(nb) for bj in range(bi, nb)}

    def hessian_pack_upper(self, H):
        k = H.shape[0]
        nb, off = self.offsets(k)
        Hp = np.zeros((nb * 32, nb * 32), np.float32)
        Hp[:k, :k] = H.numpy()
        out = np.zeros(nb * (nb + 1) // 2 * 1024, np.float32)
        for (bi, bj), o in off.items():
            out[o:o + 1024] = Hp[bi * 32:(bi + 1) * 32, bj * 32:(bj + 1) * 32].reshape(-1)
        return torch.from_numpy(out)

    def hessian_finalize_packed(self, packed, H, scale):
        k = H.shape[0]
        nb, off = self.offsets(k)
        Hp = np.zeros((nb * 32, nb * 32), np.float32)
        for (bi, bj), o in off.items():
            Hp[bi * 32:(bi + 1) * 32, bj * 32:(bj + 1) * 32] = packed.numpy()[o:o + 1024].reshape(32, 32)
        U = np.triu(Hp[:k, :k]) * np.float32(scale)
        H.copy_(torch.from_numpy(U + np.triu(U, 1).T))
        return H

    def hessian_finalize(self, H, scale, symmetric):
        U = np.triu(H.numpy()) * np.float32(scale)
        H.copy_(torch.from_numpy(U + np.triu(U, 1).T))
        return H


def _packed_hessian_case(rank, world):
    """sample-sharded raw sums of the UPPER tiles only (what the lazy hooks leave), all-reduced as packed upper blocks:
    equal on every rank, equal to all-reduce of the whole matrix + finalize, and the reference's running mean."""
    from llm_compressor_b200 import parallel
    orc, X, W = _inputs()
    X = X[:, :, :200]                      # K = 200: a ragged last 32-block (200 = 6 * 32 + 8)
    n_samples, T, K = X.shape
    mine = parallel.sample_shard(n_samples)
    S = np.zeros((K, K), np.float64)
    for j in mine:
        S += X[j].astype(np.float64).T @ X[j].astype(np.float64)
    S = S.astype(np.float32)
    tile = np.add.outer(np.arange(K) // 32, -(np.arange(K) // 32)) <= 0     # 32-blocks touching the upper triangle
    H = torch.from_numpy(np.where(tile, S, np.float32(-7.0)))               # garbage below: must not travel
    full = torch.from_numpy(S.copy())
    be = _NumpyPacked()
    n_tot = parallel.reduce_finalize_hessian_(H, len(mine), backend=be)
    H2 = torch.from_numpy(np.where(tile, S, np.float32(3.0)))
    assert parallel.reduce_finalize_hessian_(H2, len(mine), n_total=n_samples, backend=be) == n_tot
    parallel.reduce_hessian_(full, len(mine))
    be.hessian_finalize(full, 2.0 / n_tot, True)
    Href = np.zeros((K, K), np.float32)
    n = 0
    for j in range(n_samples):
        n = orc.hessian_accum(Href, X[j], n)
    rel = float(np.linalg.norm(H.numpy() - Href) / np.linalg.norm(Href))
    nb = -(-K // 32)
    # packed vs whole-matrix all-reduce: the same sums, but a ring adds the ranks in an order that depends on the element's
    # offset in the buffer, so with 3 ranks the last bit may differ -- equal to ~1 ulp, not bit for bit
    close = float((H - full).norm() / full.norm()) < 2e-7
    return dict(n_tot=n_tot, rel=rel, same=bool(close and torch.equal(H, H2) and torch.equal(H, H.t())), H=H.numpy(),
                packed_floats=int(be.hessian_pack_upper(full).numel()), expect_floats=nb * (nb + 1) // 2 * 1024)


def _finalize_case(rank, world):
    """solvers.finalize_hessian(all_reduce=True) end to end on CPU: the numpy stand-ins are patched over the three ops it
    reaches (pack / finalize_packed / finalize), H travels packed, GPTAQ's non-symmetric dXXT travels whole."""
    from llm_compressor_b200 import ops, parallel, solvers
    orc, X, W = _inputs()
    X = X[:, :, :96]
    n_samples, T, K = X.shape
    be = _NumpyPacked()
    ops.hessian_pack_upper, ops.hessian_finalize_packed = be.hessian_pack_upper, be.hessian_finalize_packed

    def finalize(H, scale, symmetric):
        return be.hessian_finalize(H, scale, True) if symmetric else H.mul_(scale)
    ops.hessian_finalize = finalize
    mine = parallel.sample_shard(n_samples)
    S = np.zeros((K, K), np.float64)
    D = np.zeros((K, K), np.float64)
    for j in mine:
        x = X[j].astype(np.float64)
        S += x.T @ x
        D += (0.5 * x[::-1]).T @ x          # any non-symmetric per-sample product
    holder = type("Holder", (), {})()
    holder.H, holder.dXXT = torch.from_numpy(np.triu(S).astype(np.float32)), torch.from_numpy(D.astype(np.float32))
    holder.nsamples, holder._h_raw = len(mine), True
    H = solvers.finalize_hessian(holder, all_reduce=True, n_total=n_samples)
    Sall = sum(X[j].astype(np.float64).T @ X[j].astype(np.float64) for j in range(n_samples)) * (2.0 / n_samples)
    Dall = sum((0.5 * X[j].astype(np.float64)[::-1]).T @ X[j].astype(np.float64) for j in range(n_samples)) * (2.0 / n_samples)
    again = solvers.finalize_hessian(holder, all_reduce=True, n_total=n_samples)      # idempotent: no second all-reduce
    return dict(n=holder.nsamples, relH=float(np.linalg.norm(H.numpy() - Sall) / np.linalg.norm(Sall)),
                relD=float(np.linalg.norm(holder.dXXT.numpy() - Dall) / np.linalg.norm(Dall)),
                sym=bool(torch.equal(H, H.t())), same=again is H)


def _rows_case(rank, world):
    from llm_compressor_b200 import parallel
    orc, X, W = _inputs()
    n_samples, T, K = X.shape
    N = W.shape[0]
    H = np.zeros((K, K), np.float32)
    n = 0
    for j in range(n_samples):
        n = orc.hessian_accum(H, X[j], n)
    cfg = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)
    full = orc.gptq_update(W, H.copy(), cfg)
    rows = parallel.row_shard(N)
    part = orc.gptq_update(W[rows], H.copy(), cfg)     # rows are independent given H
    got = parallel.gather_rows(torch.from_numpy(part), N).numpy()
    # NVFP: the per-matrix amax must be reduced over the row shards
    amax = torch.tensor([float(np.abs(W[rows]).max())])
    parallel.allreduce_max_(amax)
    return dict(equal=bool(np.array_equal(got, full)), rows=(rows.start, rows.stop),
                amax_ok=bool(float(amax) == float(np.abs(W).max())))


def _ragged_case(rank, world):
    from llm_compressor_b200 import parallel
    N, K = 7, 12   # 7 rows over 3 ranks: 3 + 3 + 1
    full = torch.arange(N * K, dtype=torch.float32).reshape(N, K)
    rows = parallel.row_shard(N)
    got = parallel.gather_rows(full[rows].clone(), N)
    return bool(torch.equal(got, full))


class _NumpyPhases:
    """CPU stand-in for the phase kernels of csrc/masks.cu (lcb_select_* / lcb_metric_* / lcb_ria_* / lcb_mask_le),
    so that the collective protocol of parallel.py can be exercised under gloo without a GPU.  TEST CODE ONLY."""

    @staticmethod
    def _keys(v):  # order-preserving uint32 keys of fp32 values (NaN last), like f2key
        b = np.ascontiguousarray(v, np.float32).view(np.uint32)
        return np.where(b >> 31, ~b, b | np.uint32(0x80000000)).astype(np.uint32)

    def select_state(self, device):
        st = torch.zeros(4 + 256, dtype=torch.int32)
        return st, st[4:]

    def select_init(self, st, kth):
        st.zero_()
        self.prefix, self.krem = 0, int(kth)

    def select_hist(self, st, scores, p):
        k = self._keys(scores.numpy())
        shift = 24 - 8 * p
        himask = 0 if p == 0 else (0xFFFFFFFF << (shift + 8)) & 0xFFFFFFFF
        sel = (k & np.uint32(himask)) == np.uint32(self.prefix & himask)
        st[4:] += torch.from_numpy(np.bincount((k[sel] >> np.uint32(shift)) & np.uint32(0xFF), minlength=256).astype(np.int32))

    def select_scan(self, st, p, thresh):
        h = st[4:].numpy().astype(np.int64)
        cum = np.cumsum(h)
        b = int(np.searchsorted(cum, self.krem, side="right"))
        b = min(b, 255)
        self.krem -= int(cum[b - 1]) if b > 0 else 0
        self.prefix |= b << (24 - 8 * p)
        st[4:] = 0
        if p == 3:
            key = np.uint32(self.prefix)
            bits = key & np.uint32(0x7FFFFFFF) if key >> 31 else ~key
            thresh[0] = float(np.array([bits], np.uint32).view(np.float32)[0])

    def metric_magnitude(self, W):
        return W.float().abs()

    def ria_sums(self, W):
        a = W.float().abs()
        return a.sum(0), a.sum(1).to(W.dtype).float()

    def ria_metric(self, W, colsum, rowsum, srow, alpha):
        a = W.abs()
        base = a / colsum.to(W.dtype)[None, :] + a / rowsum.to(W.dtype)[:, None]     # W-dtype ops, one rounding each
        return base.float() * torch.sqrt(srow.float())[None, :] ** alpha

    def mask_le(self, scores, thresh):
        return scores <= thresh[0]


def _threshold_case(rank, world):
    from llm_compressor_b200 import parallel
    orc, X, W = _inputs()
    be = _NumpyPhases()
    Wt = torch.from_numpy(W).to(torch.bfloat16)
    N = W.shape[0]
    rows = parallel.row_shard(N)
    srow = torch.from_numpy((X.astype(np.float32) ** 2).sum((0, 1)) / X.shape[0])
    out = {}
    # exact global k-th of sharded scores == sort(all)[kth]
    scores = torch.from_numpy(np.abs(W)).float()
    for kth in (0, 1, 5000, W.size // 2, W.size - 1):
        th = parallel.select_kth_sharded(scores[rows].contiguous(), kth, backend=be)
        out["kth%d" % kth] = bool(float(th) == float(np.sort(np.abs(W).reshape(-1))[kth]))
    m = parallel.mask_magnitude_sharded(Wt[rows].contiguous(), 0.5, backend=be)
    full, _ = orc.mask_magnitude(W, 0.5)
    out["magnitude"] = bool(np.array_equal(m.numpy(), full[rows]))
    m = parallel.mask_ria_sharded(Wt[rows].contiguous(), srow, 0.5, 0.5, backend=be)
    full, _ = orc.mask_ria(W, srow.numpy(), 0.5, 0.5, orc.BF16)
    out["ria_agree"] = float((m.numpy() == full[rows]).mean())
    return out


def test_row_sharded_global_thresholds_world2_and_3():
    """Distributed radix select (4 histogram all-reduces) gives the exact global order statistic; the magnitude mask
    of row shards equals the oracle's unsharded mask; RIA agrees up to the bf16 rounding of the re-associated column sums."""
    for world in (2, 3):
        for r in _run("_threshold_case", world):
            assert all(v for k, v in r.items() if k.startswith("kth")), r
            assert r["magnitude"], r
            assert r["ria_agree"] > 0.999, r


def test_sample_sharded_hessian_allreduce_matches_running_mean():
    res = _run("_hessian_case", 2)
    assert res[0]["n_tot"] == res[1]["n_tot"] == 5
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(5))
    assert res[0]["rel"] < 1e-6 and res[1]["rel"] < 1e-6
    assert np.array_equal(res[0]["H"], res[1]["H"])


def test_sample_sharded_hessian_packed_upper_allreduce():
    for world in (2, 3):
        res = _run("_packed_hessian_case", world)
        assert all(r["n_tot"] == 5 and r["same"] and r["rel"] < 1e-6 for r in res), [(r["n_tot"], r["same"], r["rel"]) for r in res]
        assert all(np.array_equal(res[0]["H"], r["H"]) for r in res)
        assert res[0]["packed_floats"] == res[0]["expect_floats"] == 7 * 8 // 2 * 1024     # about half of 224 * 224


def test_finalize_hessian_all_reduce_host_logic():
    for r in _run("_finalize_case", 2):
        assert r["n"] == 5 and r["relH"] < 1e-6 and r["relD"] < 1e-6 and r["sym"] and r["same"], r


def test_row_sharded_solve_allgather_is_bit_identical():
    res = _run("_rows_case", 2)
    assert all(r["equal"] and r["amax_ok"] for r in res)
    assert res[0]["rows"] == (0, 48) and res[1]["rows"] == (48, 96)


def test_ragged_row_shards_world3():
    assert all(_run("_ragged_case", 3))


def test_shard_index_math_single_process():
    from llm_compressor_b200 import parallel
    assert parallel.world() == (0, 1)
    assert parallel.sample_shard(10, 1, 4) == [1, 5, 9]
    sl = [parallel.row_shard(1024, r, 8) for r in range(8)]
    assert [s.start for s in sl] == list(range(0, 1024, 128)) and sl[-1].stop == 1024
    sl = [parallel.row_shard(5, r, 4) for r in range(4)]
    assert [(s.start, s.stop) for s in sl] == [(0, 2), (2, 4), (4, 5), (5, 5)]

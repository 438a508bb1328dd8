"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed.

The reference is single-device; these are the two natural shardings of its layer loop:

  * calibration samples are sharded over ranks (`sample_shard`).  Each rank's hooks accumulate RAW sums
    S_r = sum_j X_j^T X_j (solvers.HESSIAN_MODE == "lazy"); `reduce_hessian_` all-reduces them and the
    single finalize applies 2 / n_total -- algebraically the running mean of ref: gptq/core.py:113-119.
  * Linear output rows are sharded for the solve (`row_shard`): rows are independent given U, the
    act-order permutation and per-row / per-group parameters (ref: gptq/core.py:226-265 touches rows
    only through W[:, cols]); `gather_rows` all-gathers the updated rows so every rank holds the full
    layer for the next calibration forward.  Statistics that span rows are reduced explicitly:
    NVFP's per-matrix amax (ref: nvfp_quant.py:87) with `allreduce_max_`; SparseGPT's per-block threshold
    and the RIA / magnitude global thresholds with an exact distributed radix select (`select_kth_sharded`:
    4 all-reduces of 256 counters, no gather of the scores); RIA's column sums with one all-reduce of [K].

Only collectives and index arithmetic live here (they work with NCCL on CUDA tensors and with gloo on
CPU tensors, which is how tests/test_parallel_cpu.py covers the N > 1 logic without a GPU); every
numeric stage stays in liblcb200.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group; (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def sample_shard(n_samples, rank=None, world_size=None):
    """Calibration samples of this rank: rank, rank + world, ... (round-robin keeps the shards balanced
    when n_samples is not a multiple of world)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, n_samples, world_size))


def row_shard(n_rows, rank=None, world_size=None):
    """Contiguous slice of Linear output rows owned by this rank (ceil-divided; trailing ranks may be short
    or empty)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    per = (n_rows + world_size - 1) // world_size
    lo = min(n_rows, rank * per)
    return slice(lo, min(n_rows, lo + per))


def reduce_hessian_(H, nsamples, dxxt=None, group=None, n_total=None):
    """all-reduce(SUM) of the raw per-rank sums (and of the per-rank sample count).  Returns n_total; the
    caller finalises ONCE with scale 2 / n_total (ops.hessian_finalize).
    n_total: the number of samples over all ranks when the caller knows it (a static `sample_shard` split: the size of
    the calibration set) -- skips the all-reduce of the counts and, with it, the host read-back that drains the stream
    once per Hessian."""
    r, w = world()
    if w == 1:
        return nsamples
    dist.all_reduce(H, op=dist.ReduceOp.SUM, group=group)
    if dxxt is not None:
        dist.all_reduce(dxxt, op=dist.ReduceOp.SUM, group=group)
    if n_total is not None:
        return int(n_total)
    n = torch.tensor([float(nsamples)], dtype=torch.float64, device=H.device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return int(round(float(n.item())))


def reduce_finalize_hessian_(H, nsamples, n_total=None, group=None, backend=None):
    """`reduce_hessian_` + `ops.hessian_finalize(H, 2 / n_total, True)` for the symmetric H with HALF the all-reduce bytes:
    the lazy hooks fill only the tiles touching the upper triangle, so the ranks pack the upper 32 x 32 blocks into one
    contiguous buffer (`lcb_hessian_pack_upper`), all-reduce that, and the finalize reads the packed sums
    (`lcb_hessian_finalize_packed`: scale + mirror, what it does anyway).  Llama-3.2-3B: 5.4 instead of 10.7 GB per model
    over NVLink.  Returns n_total.  `backend` provides hessian_pack_upper / hessian_finalize_packed / hessian_finalize
    (default: llm_compressor_b200.ops, i.e. the CUDA kernels)."""
    if backend is None:
        from . import ops as backend
    r, w = world()
    if w == 1:
        n = int(nsamples)
        backend.hessian_finalize(H, 2.0 / max(n, 1), True)
        return n
    packed = backend.hessian_pack_upper(H)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    if n_total is None:
        cnt = torch.tensor([float(nsamples)], dtype=torch.float64, device=H.device)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
        n_total = int(round(float(cnt.item())))
    backend.hessian_finalize_packed(packed, H, 2.0 / max(int(n_total), 1))
    return int(n_total)


def reduce_rownorm_(scaler_raw, nsamples, group=None):
    """Wanda / RIA row norms (ref: wanda/core.py:92-105) kept as raw sums sum_t X[t,:]^2 per rank:
    all-reduce(SUM), then the caller divides by n_total."""
    r, w = world()
    if w == 1:
        return nsamples
    dist.all_reduce(scaler_raw, op=dist.ReduceOp.SUM, group=group)
    n = torch.tensor([float(nsamples)], dtype=torch.float64, device=scaler_raw.device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return int(round(float(n.item())))


def allreduce_max_(t, group=None):
    """all-reduce(MAX), e.g. of lcb_nvfp_global_amax's result when a matrix is row-sharded."""
    r, w = world()
    if w > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def gather_rows(local_rows, n_rows, group=None):
    """Reassemble [n_rows, k] from the per-rank row slices of `row_shard` (all ranks get the result)."""
    r, w = world()
    if w == 1:
        return local_rows
    per = (n_rows + w - 1) // w
    k = local_rows.shape[1]
    padded = local_rows
    if local_rows.shape[0] != per:  # short / empty trailing shard: pad so that all_gather sees equal shapes
        padded = torch.zeros((per, k), dtype=local_rows.dtype, device=local_rows.device)
        padded[: local_rows.shape[0]] = local_rows
    out = torch.empty((w * per, k), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return out[:n_rows]


# ------------------------------------------------------------------ global thresholds over row shards
def select_kth_sharded(scores_local, kth, group=None, backend=None):
    """Exact k-th smallest (0-based) of fp32 scores spread over the ranks, without gathering them: MSB-first
    radix select, per pass the local 256-bin histograms are summed with one all-reduce and every rank takes the
    same branch (ref semantics: `sort(flatten)[kth]` of ria/core.py:124, magnitude/core.py:41).
    Returns a [1] fp32 tensor, identical on every rank.  `backend` provides select_state / select_init /
    select_hist / select_scan (default: llm_compressor_b200.ops, i.e. the CUDA kernels)."""
    if backend is None:
        from . import ops as backend
    r, w = world()
    state, hist = backend.select_state(scores_local.device)
    thresh = torch.zeros(1, dtype=torch.float32, device=scores_local.device)
    backend.select_init(state, kth)
    flat = scores_local.reshape(-1)
    for p in range(4):
        backend.select_hist(state, flat, p)
        if w > 1:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        backend.select_scan(state, p, thresh)
    return thresh


def _total(n_local, device, group=None):
    r, w = world()
    if w == 1:
        return int(n_local)
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def mask_magnitude_sharded(W_local, ratio, group=None, backend=None):
    """magnitude mask of a row-sharded weight (ref: magnitude/core.py:38-43): global threshold, local rows."""
    if backend is None:
        from . import ops as backend
    numel = _total(W_local.numel(), W_local.device, group)
    scores = backend.metric_magnitude(W_local)
    kth = min(int(numel * ratio), numel - 1)
    return backend.mask_le(scores, select_kth_sharded(scores, kth, group, backend))


def mask_ria_sharded(W_local, scaler_row, ratio, alpha, group=None, backend=None):
    """RIA mask of a row-sharded weight (ref: ria/core.py:118-126): column sums all-reduced in fp32 before the
    rounding to W's dtype, row sums local, global threshold by `select_kth_sharded`."""
    if backend is None:
        from . import ops as backend
    r, w = world()
    numel = _total(W_local.numel(), W_local.device, group)
    colsum, rowsum = backend.ria_sums(W_local)
    if w > 1:
        dist.all_reduce(colsum, op=dist.ReduceOp.SUM, group=group)
    scores = backend.ria_metric(W_local, colsum, rowsum, scaler_row, alpha)
    kth = min(int(numel * ratio), numel - 1)
    return backend.mask_le(scores, select_kth_sharded(scores, kth, group, backend))


def sparsegpt_update_sharded(W_local, U, sparsity, block=128, group=None):
    """SparseGPT block loop on this rank's rows with the exact global per-block threshold
    (ref: sparsegpt/core.py:201-203; the histogram all-reduce runs inside the loop, 4 per block)."""
    from . import ops
    r, w = world()
    if w == 1:
        return ops.sparsegpt_update(W_local, U, sparsity, block)
    n_total = _total(W_local.shape[0], W_local.device, group)
    if W_local.shape[0] == 0:
        # an empty row shard (ceil-divided shards: trailing ranks can be empty) still takes part in every collective of
        # the loop -- 4 histogram all-reduces per 128-column block -- with zero counts, or the other ranks would block
        zero = torch.zeros(256, dtype=torch.int32, device=U.device)
        for _ in range(4 * ((U.shape[0] + block - 1) // block)):
            zero.zero_()
            dist.all_reduce(zero, op=dist.ReduceOp.SUM, group=group)
        return W_local
    return ops.sparsegpt_update(W_local, U, sparsity, block, n_total=n_total,
                                reduce=lambda hist: dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group))


def nvfp_quantize_rows_sharded(quantizer, W_local, group=None):
    """NVFP fake-quant of this rank's rows with the per-MATRIX amax of the reference (nvfp_quant.py:87): local
    lcb_nvfp_global_amax, all-reduce(MAX), lcb_qdq with nv_amax -- bit-identical to quantising the unsharded matrix."""
    amax = quantizer.global_amax(W_local)
    allreduce_max_(amax, group)
    return quantizer(W_local, nv_amax=amax)

"""Per-layer numerics with the reference's function signatures (drop-in for L1 of SURVEY section 1).

  cache_hessian_weight / update_weight            ref: quantization/calibrations/gptq/core.py:103-119,163-281
  cache_hessian_dxxt_weight / gptaq_update_weight ref: quantization/calibrations/gptaq/core.py:116-141,198-335
  Wrapper / prune_weight                          ref: pruning/sparsegpt/core.py:78-101,160-228
  cache_scalar_row / wanda_mask / ria_mask / magnitude_mask
                                                  ref: pruning/wanda/core.py:92-126, ria/core.py:118-126,
                                                       magnitude/core.py:38-43

The host code below is glue only (tensor bookkeeping, permutation indices, dtype casts); every
arithmetic stage runs in liblcb200.so through llm_compressor_b200.ops.
"""
import torch
import torch.nn as nn

from . import _lib, ops, parallel
from .quantizers import DeferredStatus


def _is_conv1d(layer):
    return type(layer).__name__ == "Conv1D"


# ----------------------------------------------------------------------------------- hooks
# "lazy": the hooks accumulate raw sums S += X^T X (symmetric half only) and the running-mean factor
#         2/n is applied once, by finalize_hessian(), right before the solver reads H -- algebraically the
#         reference's H (gptq/core.py:113-119 telescopes to (2/n) * sum_j X_j^T X_j).
# "exact": every hook call leaves H equal to the reference's running mean (beta = n/(n+1), alpha = 2/(n+1)).
HESSIAN_MODE = "lazy"
# lazy mode: hook inputs accumulated per kernel launch (ops.HessianAccumulator keeps references to the deferred inputs
# until the launch -- they must not be modified in place meanwhile); 1 = one launch per hook call
HESSIAN_DEFER = 4


def _accumulate(holder, x, dxxt=None, x_fp=None):
    if HESSIAN_MODE == "exact":
        holder.nsamples = ops.hessian_accum(holder.H, x, holder.nsamples, dxxt=dxxt, x_fp=x_fp)
    elif dxxt is not None or HESSIAN_DEFER <= 1:
        holder.nsamples = ops.hessian_accum_raw(holder.H, x, holder.nsamples, dxxt=dxxt, x_fp=x_fp)
        holder._h_raw = True
    else:
        acc = getattr(holder, "_h_acc", None)
        if acc is None or acc.H is not holder.H:
            acc = holder._h_acc = ops.HessianAccumulator(holder.H, HESSIAN_DEFER)
            acc.n = holder.nsamples
        holder.nsamples = acc.add(x)
        holder._h_raw = True


def finalize_hessian(holder, all_reduce=False, n_total=None):
    """Bring holder.H (and holder.dXXT) to the reference's running-mean value. Idempotent.
    all_reduce (sample-sharded calibration, one process per GPU): the raw per-rank sums and sample counts are summed over
    the ranks first (parallel.reduce_finalize_hessian_), so every rank ends with the Hessian of ALL samples; n_total = the known size
    of the whole calibration set lets that skip the count all-reduce and its host read-back."""
    if getattr(holder, "_h_raw", False):
        acc = getattr(holder, "_h_acc", None)
        if acc is not None:
            acc.flush()
            del holder._h_acc
        dxxt = getattr(holder, "dXXT", None)
        if all_reduce:
            # the symmetric H travels as its packed upper blocks (half the bytes); dXXT is not symmetric: whole matrix
            holder.nsamples = parallel.reduce_finalize_hessian_(holder.H, holder.nsamples, n_total=n_total)
            if dxxt is not None:
                parallel.reduce_hessian_(dxxt, holder.nsamples, n_total=holder.nsamples)
        else:
            ops.hessian_finalize(holder.H, 2.0 / max(holder.nsamples, 1), True)
        if dxxt is not None:
            ops.hessian_finalize(dxxt, 2.0 / max(holder.nsamples, 1), False)
        holder._h_raw = False
    return holder.H


def cache_hessian_weight(m, x, y):
    """forward hook: m.weight_quantizer.H / .nsamples (ref: gptq/core.py:103-119)"""
    _accumulate(m.weight_quantizer, x[0].detach())


def cache_hessian_dxxt_weight(m, x, y):
    """forward hook incl. the asymmetric-calibration term (ref: gptaq/core.py:116-141)"""
    q = m.weight_quantizer
    _accumulate(q, x[0].detach(), dxxt=q.dXXT, x_fp=m.fp_inp[0])
    del m.fp_inp[0]


def cache_scalar_row(m, x, y):
    """forward hook: running mean of squared input-channel norms (ref: wanda/core.py:92-105)"""
    m.nsamples = ops.rownorm_accum(m.scaler_row, x[0].detach(), m.nsamples)


# ----------------------------------------------------------------------------------- GPTQ / GPTAQ
class Factor:
    """Everything update_weight derives from the Hessian alone: dead columns, act-order permutation
    and U = chol(H^-1, upper).  Linears fed by the same activations (q/k/v, gate/up) have identical
    H, so a driver may compute one Factor per group and reuse it (the reference recomputes it per
    Linear from identical copies, ref: gptq/core.py:121-137; gptaq shares H the same way, :143-159)."""

    def __init__(self, dead, perm, invperm, col_perm, U, P, group_size, pending=None, p_args=None):
        self.dead, self.perm, self.invperm, self.col_perm = dead, perm, invperm, col_perm
        self.U, self.P, self.group_size = U, P, group_size
        self._pending, self._p_args = pending, p_args

    def resolve(self):
        """Deferred outcome of the factorisation (ops.PendingFactor): True when U had to be recomputed with the
        reference's 10x damping retry -- whatever was solved with the first U must be solved again."""
        if self._pending is None:
            return False
        redone = self._pending.resolve()
        self._pending = None
        if redone and self._p_args is not None:
            self.P = ops.gptaq_p(self._p_args[0], self.U, self._p_args[1])
        self._p_args = None
        return redone


def factorize(H, group_size, actorder=True, percdamp=0.01, dXXT=None, alpha=0.25):
    """Dead fix (in place on H, idempotent), act-order permutation, damped Cholesky-inverse; with
    dXXT also GPTAQ's P.  group_size is the quantizer's value BEFORE find_params (so -1 / 0 select
    the per-column branch, ref: gptq/core.py:171,181)."""
    K = H.shape[0]
    dead = ops.dead_fix(H)
    if dXXT is not None:
        dXXT.masked_fill_(dead.unsqueeze(0), 0)
    per_col = group_size in (0, -1)
    perm = invperm = col_perm = None
    if actorder:
        diag = torch.diag(H)
        if per_col:
            perm = torch.argsort(diag, descending=True)
            col_perm = perm
        else:
            perm = torch.argsort(diag.reshape(-1, group_size).sum(-1), descending=True)
            col_perm = (perm.unsqueeze(1) * group_size + torch.arange(group_size, device=perm.device)).reshape(-1)
        invperm = torch.argsort(perm)
    U, pending = ops.chol_inv_upper(H, perm=col_perm, percdamp=percdamp, defer=True)
    P = p_args = None
    if dXXT is not None:
        if col_perm is not None:
            dXXT = dXXT[col_perm][:, col_perm].contiguous()
        P = ops.gptaq_p(dXXT, U, alpha)
        p_args = (dXXT, alpha)
    return Factor(dead, perm, invperm, col_perm, U, P, group_size, pending, p_args)


def _solve(q, W, factor, block_size, shard_rows=False):
    """GPTQ / GPTAQ solve of one weight matrix W [N, K] (rows = outputs; bf16 or fp32, not modified) with quantizer
    q and a ready Factor; returns the dequantised result [N, K] in W's dtype and the original column order.
    shard_rows (under torch.distributed, one process per GPU): every rank solves its slice of the output rows -- rows are
    independent given U, the permutation and row-wise parameters -- and the slices are all-gathered."""
    if shard_rows and parallel.world()[1] > 1 and type(q).__name__ != "NVFPQuantizer" and q.group_size != 0:
        n_rows = W.shape[0]
        mine = W[parallel.row_shard(n_rows)].contiguous()
        if mine.shape[0] == 0:      # empty trailing shard: nothing to solve, but the collectives below still run
            factor.resolve()
            part = mine
        else:
            part = _solve(q, mine, factor, block_size)
        return parallel.gather_rows(part, n_rows)
    N, K = W.shape
    group_size = factor.group_size
    per_col = group_size in (0, -1)
    # W.float(), MASK = W != 0, W[:, dead] = 0 and the act-order gather in one pass (ref: gptq/core.py:164-201)
    Wp, keep = ops.gptq_gather(W.contiguous(), factor.col_perm, factor.dead)
    # per-column branch: the reference takes the per-row parameters before the permutation (ref :179-185); row
    # max / min do not depend on the column order, so the permuted matrix gives the same values.  Grouped branch:
    # static groups of the re-ordered W (ref :198).
    nan_st = DeferredStatus(Wp.device) if getattr(q, "check_nan", True) else None
    scales, zeros = q.find_params(Wp, status=nan_st) if nan_st is not None else q.find_params(Wp)
    if per_col:
        s2 = scales.float().reshape(-1, 1).expand(N, 1).contiguous()
        z2 = zeros.float().reshape(-1, 1).expand(N, 1).contiguous()
        grp = -1
    else:
        s2 = scales.float().reshape(N, K // group_size).contiguous()
        z2 = zeros.float().reshape(N, K // group_size).contiguous()
        grp = group_size
    Q = ops.gptq_block_update(q._cfg(), Wp, factor.U, s2, z2, keep, grp, P=factor.P, block=block_size)
    # inverse permutation + cast back to the weight dtype (ref :267-278)
    out = ops.gptq_scatter(Q, factor.col_perm, W.dtype)
    # the two error conditions the reference reads synchronously, looked at AFTER the solve has been enqueued
    if nan_st is not None:
        assert not (nan_st.value() & _lib.ST_NAN_SCALE), "NaN in quantization scales"
    if factor.resolve():   # not positive definite at 1 % damping: U was redone with the 10x retry (ref: gptq/core.py:216-221)
        return _solve(q, W, factor, block_size)
    return out


def _layer_w(layer):
    W = layer.weight.data
    if _is_conv1d(layer):
        W = W.t()
    return W.contiguous()


def _store_w(layer, Q):
    if _is_conv1d(layer):
        Q = Q.t()
    layer.weight.data = Q.reshape(layer.weight.shape).to(layer.weight.data.dtype).contiguous()


def _update_weight(layer, device, block_size, percdamp, actorder, alpha=None, factor=None, shard_rows=False):
    q = layer.weight_quantizer
    W = _layer_w(layer)
    group_size = q.group_size  # read before find_params mutates -1 (ref: gptq/core.py:171)
    if factor is None:
        H = finalize_hessian(q)
        dXXT = q.dXXT if alpha is not None else None
        factor = factorize(H, group_size, actorder, percdamp, dXXT, alpha if alpha is not None else 0.25)
    for attr in ("H", "dXXT"):
        if hasattr(q, attr):
            delattr(q, attr)
    _store_w(layer, _solve(q, W, factor, block_size, shard_rows))


def quantizer_key(q):
    """Everything of a weight quantizer that changes the solve's result (used to decide what may share a stacked solve)."""
    return (type(q).__name__, getattr(q.format, "name", str(q.format)), q.group_size, q.axes, bool(q.zero_point),
            int(getattr(q, "scale_ebits", 8)), bool(getattr(q, "mse", False)))


def update_weights_shared(layers, device, factor, block_size=128, shard_rows=False):
    """Solve several Linears that share one Factor (q/k/v, gate/up: same input, same H) as ONE stacked
    [sum N_i, K] problem.  Rows are independent given U, the permutation and row-wise quantiser
    parameters, so the result equals calling update_weight on each layer (ref: gptq/core.py:129-137 loops
    over them one by one); it just runs K/128 block steps once instead of once per Linear.
    Quantisers with a per-matrix statistic (NVFP's global amax, per-tensor scales) are not stacked."""
    layers = list(layers)
    q0 = layers[0].weight_quantizer
    # one stacked solve runs q0's find_params / format for every row: only Linears whose quantizers agree in EVERYTHING
    # that shapes the result may be stacked (the reference supports per-layer mixed precision, e.g. int4 q_proj next to
    # int8 k_proj at the same group size: parser.py register_4_to_8bit_config); the others are solved one by one with
    # the shared Factor, each with its own quantizer as the reference does (gptq/core.py:129-137).
    stackable = len(layers) > 1 and type(q0).__name__ != "NVFPQuantizer" and q0.group_size != 0 and q0.axes == -1 \
        and not any(_is_conv1d(l) for l in layers) and all(quantizer_key(l.weight_quantizer) == quantizer_key(q0) for l in layers)
    if not stackable:
        for l in layers:
            _update_weight(l, device, block_size, 0.01, True, factor=factor, shard_rows=shard_rows)
        return
    sizes = [l.weight.shape[0] for l in layers]
    W = torch.cat([l.weight.data for l in layers], 0).contiguous()
    for l in layers:
        for attr in ("H", "dXXT"):
            if hasattr(l.weight_quantizer, attr):
                delattr(l.weight_quantizer, attr)
    Q = _solve(q0, W, factor, block_size, shard_rows)
    o = 0
    for l, n in zip(layers, sizes):
        _store_w(l, Q[o:o + n])
        o += n


def update_weight(layer, device, block_size=128, percdamp=0.01, actorder=False, factor=None):
    """ref: quantization/calibrations/gptq/core.py:163-281"""
    _update_weight(layer, device, block_size, percdamp, actorder, alpha=None, factor=factor)


def gptaq_update_weight(layer, device, block_size=128, percdamp=0.01, actorder=False, alpha=0.25, factor=None):
    """ref: quantization/calibrations/gptaq/core.py:198-335"""
    _update_weight(layer, device, block_size, percdamp, actorder, alpha=alpha, factor=factor)


# ----------------------------------------------------------------------------------- SparseGPT
class Wrapper:
    """ref: pruning/sparsegpt/core.py:78-101"""

    def __init__(self, module, device):
        self.module = module
        columns = module.weight.shape[1]
        self.nsamples = 0
        self.H = torch.zeros((columns, columns), device=device)

    def cache_hessian_weight(self, x, y):
        _accumulate(self, x[0].detach())


def prune_weight(layer, device, sparsity_ratio, block_size=128, percdamp=0.01, shard_rows=False, all_reduce=False,
                 n_total=None):
    """ref: pruning/sparsegpt/core.py:160-228.  `shard_rows=True` under torch.distributed (one process per GPU):
    every rank factors the Hessian of ALL samples, solves its slice of the output rows with the exact global per-block
    threshold (parallel.sparsegpt_update_sharded) and the rows are all-gathered.  `all_reduce=True` when the
    ranks' hooks saw different calibration samples (parallel.sample_shard): the raw sums are all-reduced first
    (finalize_hessian; n_total = the size of the whole calibration set, if known) -- leave it False when every rank ran the
    whole calibration set or already finalised with all_reduce."""
    W = layer.module.weight.data.clone()
    W = W.float().contiguous()
    H = finalize_hessian(layer, all_reduce=all_reduce, n_total=n_total)
    del layer.H
    dead = ops.dead_fix(H)
    W.masked_fill_(dead.unsqueeze(0), 0)
    U = ops.chol_inv_upper(H, perm=None, percdamp=percdamp)
    del H
    if shard_rows and parallel.world()[1] > 1:
        n_rows = W.shape[0]
        part = parallel.sparsegpt_update_sharded(W[parallel.row_shard(n_rows)].contiguous(), U, sparsity_ratio, block_size)
        W = parallel.gather_rows(part, n_rows)
    else:
        ops.sparsegpt_update(W, U, sparsity_ratio, block=block_size)
    layer.module.weight.data = W.reshape(layer.module.weight.shape).to(layer.module.weight.data.dtype)


# ----------------------------------------------------------------------------------- masks
def wanda_prune_(module, sparsity_ratio):
    """W[mask] = 0 with the Wanda mask (ref: wanda/core.py:116-126)"""
    W = module.weight.data
    mask = ops.mask_wanda(W.contiguous(), module.scaler_row, sparsity_ratio)
    if W.is_contiguous():
        ops.apply_mask(W, mask)
    else:
        W[mask] = 0
    return mask


def ria_prune_(module, sparsity_ratio, alpha):
    """ref: ria/core.py:118-126"""
    W = module.weight.data
    mask = ops.mask_ria(W.contiguous(), module.scaler_row, sparsity_ratio, alpha)
    if W.is_contiguous():
        ops.apply_mask(W, mask)
    else:
        W[mask] = 0
    return mask


def magnitude_prune_(module, sparsity_ratio):
    """ref: magnitude/core.py:38-43"""
    W = module.weight.data
    mask = ops.mask_magnitude(W.contiguous(), sparsity_ratio)
    if W.is_contiguous():
        ops.apply_mask(W, mask)
    else:
        W[mask] = 0
    return mask


def find_layers(module, layers=(nn.Conv2d, nn.Linear), name=""):
    """ref: utils/module.py:54-64"""
    if isinstance(module, layers):
        return {name: module}
    res = {}
    for name1, child in module.named_children():
        res.update(find_layers(child, layers=layers, name=name + "." + name1 if name != "" else name1))
    return res

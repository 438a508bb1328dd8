"""Layer-wise calibration loops (L2 of SURVEY section 1) with the reference's signatures.

They stay Python / PyTorch on purpose (north_star): walk `model.get_layers()`, run the
calibration forwards with hooks, and call the per-layer numerics of llm_compressor_b200.solvers.
  rtn        ref: quantization/calibrations/rtn/core.py:17-60
  gptq       ref: quantization/calibrations/gptq/core.py:23-160
  gptaq      ref: quantization/calibrations/gptaq/core.py:24-195
  sparsegpt  ref: pruning/sparsegpt/core.py:23-157
  wanda      ref: pruning/wanda/core.py:22-145
  ria        ref: pruning/ria/core.py:22-145
  magnitude  ref: pruning/magnitude/core.py:14-52
The reference loads wikitext2 through `get_loaders` (network); here the calibration batches are
passed in (`dataloader`, a list of `(input_ids[1, seq_len], _)`), synthetic by default.
The model is duck-typed exactly like the reference's wrappers: `config.use_cache`,
`get_layers()`, `get_sequential(mode)`, `move_embed(device)`, `lm_head`.
"""
import functools

import torch
from torch import nn

from . import parallel, ops, solvers


def synthetic_loader(vocab_size, nsamples=128, seqlen=2048, seed=0):
    """Seeded random token batches shaped like get_loaders' output (ref: utils/dataset.py:99)."""
    return [
        (torch.randint(0, vocab_size, (1, seqlen), generator=torch.Generator().manual_seed(seed + i)), None)
        for i in range(nsamples)
    ]


class _Catcher(nn.Module):
    def __init__(self, module, inps, layer_kwargs):
        super().__init__()
        self.module = module
        self._inps = inps
        self._kw = layer_kwargs

    def forward(self, inp, **kwargs):
        self._inps.append(inp)
        self._kw.update(kwargs)
        raise ValueError  # early exit, like the reference's Catcher hack


def _catch_inputs(model, device, dataloader):
    layers = model.get_layers()
    inps, layer_kwargs = [], {}
    dt = next(layers[0].parameters()).dtype
    if dt != torch.bfloat16:
        # the calibration kernels take bf16 activations (exact bf16 x bf16 products on the tensor cores); the reference's
        # hooks upcast whatever comes with x.float().  Fail before any calibration work rather than in the first hook.
        raise NotImplementedError("llm_compressor_b200 drivers calibrate bfloat16 models only (got %s): load the checkpoint "
                                  "with torch_dtype=torch.bfloat16" % dt)
    layers[0] = layers[0].to(device)
    model.move_embed(device)
    layers[0] = _Catcher(layers[0], inps, layer_kwargs)
    for batch in dataloader:
        try:
            model(batch[0].to(device))
        except ValueError:
            pass
    layers[0] = layers[0].module
    layers[0] = layers[0].cpu()
    model.move_embed("cpu")
    inps = torch.cat(inps, dim=0)
    return layers, inps, torch.zeros_like(inps), layer_kwargs


def _first(out):
    """layer(x)[0] as the reference writes it (gptq/core.py:140): the hidden states of a tuple, or -- transformers 5.x
    layers return a tensor -- its single batch entry."""
    return out[0]


def _default_loader(model, n_samples, seq_len, dataloader):
    if dataloader is not None:
        return dataloader
    return synthetic_loader(model.config.vocab_size, n_samples, seq_len)


@torch.no_grad()
def rtn(model, device, mse=False, verbose=True):
    use_cache = model.config.use_cache
    model.config.use_cache = False
    model.eval()
    layers = model.get_layers()
    for i in range(len(layers)):
        layer = layers[i].to(device)
        subset = solvers.find_layers(layer)
        for name in subset:
            subset[name].weight_quantizer.mse = mse
            W = subset[name].weight.data
            out = subset[name].weight_quantizer(W)
            ops.apply_mask(out, W == 0)  # `* MASK` of the reference (rtn/core.py:39-41)
            subset[name].weight.data = out
            del subset[name].weight_quantizer
        layers[i] = layer.cpu()
        del layer
    model.lm_head.to(device)
    model.lm_head.weight_quantizer.mse = mse
    model.lm_head.weight.data = model.lm_head.weight_quantizer(model.lm_head.weight.data)
    del model.lm_head.weight_quantizer
    model.lm_head.cpu()
    model.config.use_cache = use_cache


class FPInputsCache:
    """ref: quantization/calibrations/gptaq/model_utils.py:5-41"""

    def __init__(self, sequential):
        self.fp_cache = {}
        self.names = []
        for seq in sequential:
            self.names += seq
        for name in self.names:
            self.fp_cache[name] = []
        self.handles = []

    def cache_fp_input(self, m, inp, out, name):
        inp = inp[0].detach()
        if len(inp.shape) == 3:
            inp = inp.reshape((-1, inp.shape[-1]))
        self.fp_cache[name] += [inp]  # token-major [T, K] (the kernels read X un-transposed)

    def add_hook(self, full):
        for name in self.names:
            self.handles.append(full[name].register_forward_hook(functools.partial(self.cache_fp_input, name=name)))

    def clear_hook(self):
        for h in self.handles:
            h.remove()
        self.handles = []

    def clear_cache(self):
        for name in self.names:
            self.fp_cache[name] = []


def _quantize_head(model, device, mse):
    model.lm_head.to(device)
    model.lm_head.weight_quantizer.mse = mse
    model.lm_head.weight.data = model.lm_head.weight_quantizer(model.lm_head.weight.data)
    del model.lm_head.weight_quantizer
    model.lm_head.cpu()


@torch.no_grad()
def gptq(model, device, n_samples=512, seq_len=2048, mse=False, verbose=True, dataloader=None, distributed=False):
    """ref: quantization/calibrations/gptq/core.py:21-160.
    distributed=True under torch.distributed (one process per GPU): every rank runs the calibration
    forwards of ITS samples only (parallel.sample_shard), the raw Hessian sums are all-reduced, every rank factors the
    Hessian, solves its slice of the output rows and the slices are all-gathered -- all ranks end with the same model."""
    dist_on = bool(distributed) and parallel.world()[1] > 1
    use_cache = model.config.use_cache
    model.config.use_cache = False
    model.eval()
    sequential = model.get_sequential(mode="true")
    loader = _default_loader(model, n_samples, seq_len, dataloader)
    n_total = None
    if dist_on:
        n_total = sum(int(b[0].shape[0]) if hasattr(b[0], "shape") else 1 for b in loader)   # hooks count x.shape[0] per call
        loader = [loader[j] for j in parallel.sample_shard(len(loader))]     # this rank's calibration samples
    layers, inps, outs, layer_kwargs = _catch_inputs(model, device, loader)
    n_samples = inps.shape[0]
    for i in range(len(layers)):
        layer = layers[i].to(device)
        full = solvers.find_layers(layer)
        for names in sequential:
            subset = {n: full[n] for n in names}
            # The Linears of one sequential group see the same input, so their Hessians are identical
            # (the reference accumulates one copy per Linear, gptq/core.py:121-123): hook the first
            # module only and share H, the act-order permutation and the Cholesky factor.
            first = list(subset.keys())[0]
            columns = subset[first].weight.shape[1]
            fq = subset[first].weight_quantizer
            fq.nsamples = 0
            fq.H = torch.zeros((columns, columns), device=device)
            handle = subset[first].register_forward_hook(solvers.cache_hessian_weight)
            for j in range(n_samples):
                layer(inps[j].unsqueeze(0), **layer_kwargs)
            handle.remove()
            H = solvers.finalize_hessian(fq, all_reduce=dist_on, n_total=n_total)
            by_group = {}
            for name in subset:
                wq = subset[name].weight_quantizer
                wq.mse = mse
                # the Factor depends on the group size only (act-order granularity); within one Factor,
                # update_weights_shared stacks just the Linears whose quantizers agree completely (solvers.quantizer_key)
                key = wq.group_size if wq.group_size not in (0, -1) else -1
                by_group.setdefault(key, []).append(name)
            for key, members in by_group.items():
                gs = subset[members[0]].weight_quantizer.group_size
                factor = solvers.factorize(H, gs, actorder=True, percdamp=0.01)
                # one stacked [sum N, K] solve per set of identically configured Linears that share this factor
                by_q = {}
                for n in members:
                    by_q.setdefault(solvers.quantizer_key(subset[n].weight_quantizer), []).append(n)
                for same in by_q.values():
                    solvers.update_weights_shared([subset[n] for n in same], device, factor, block_size=128, shard_rows=dist_on)
                del factor
            for name in subset:
                del subset[name].weight_quantizer
            del H
        for j in range(n_samples):
            outs[j] = _first(layer(inps[j].unsqueeze(0), **layer_kwargs))
        layers[i] = layer.cpu()
        del layer
        inps, outs = outs, inps
    _quantize_head(model, device, mse)
    model.config.use_cache = use_cache


@torch.no_grad()
def gptaq(model, device, n_samples=512, seq_len=2048, mse=False, verbose=True, dataloader=None):
    use_cache = model.config.use_cache
    model.config.use_cache = False
    model.eval()
    sequential = model.get_sequential(mode="true")
    layers, inps, outs, layer_kwargs = _catch_inputs(model, device, _default_loader(model, n_samples, seq_len, dataloader))
    n_samples = inps.shape[0]
    fp_inputs_cache = FPInputsCache(sequential)
    fp_inps = inps.clone()
    for i in range(len(layers)):
        layer = layers[i].to(device)
        full = solvers.find_layers(layer)
        fp_inputs_cache.add_hook(full)
        for j in range(n_samples):
            fp_inps[j] = _first(layer(fp_inps[j].unsqueeze(0), **layer_kwargs))
        fp_inputs_cache.clear_hook()
        for names in sequential:
            subset = {n: full[n] for n in names}
            for name in subset:
                columns = subset[name].weight.shape[1]
                subset[name].weight_quantizer.mse = mse
                subset[name].weight_quantizer.nsamples = 0
                subset[name].weight_quantizer.H = torch.zeros((columns, columns), device=device)
                subset[name].weight_quantizer.dXXT = torch.zeros((columns, columns), device=device)
                subset[name].fp_inp = fp_inputs_cache.fp_cache[name]
            first = list(subset.keys())[0]
            handle = subset[first].register_forward_hook(solvers.cache_hessian_dxxt_weight)
            for j in range(n_samples):
                layer(inps[j].unsqueeze(0), **layer_kwargs)
            handle.remove()
            solvers.finalize_hessian(subset[first].weight_quantizer)
            for name in subset:  # H and dXXT are shared by the whole group (gptaq/core.py:149-159)
                if name != first:
                    subset[name].weight_quantizer.H = subset[first].weight_quantizer.H
                    subset[name].weight_quantizer.dXXT = subset[first].weight_quantizer.dXXT
            for name in subset:
                # H / dXXT stay shared: the solver only applies the (idempotent) dead-column fix in place
                solvers.gptaq_update_weight(layer=subset[name], device=device, block_size=128, percdamp=0.01,
                                            actorder=True, alpha=0.25)
                del subset[name].weight_quantizer
                del subset[name].fp_inp
        for j in range(n_samples):
            outs[j] = _first(layer(inps[j].unsqueeze(0), **layer_kwargs))
        fp_inputs_cache.clear_cache()
        layers[i] = layer.cpu()
        del layer
        inps, outs = outs, inps
    _quantize_head(model, device, mse)
    model.config.use_cache = use_cache


def _prune_loop(model, device, n_samples, seq_len, dataloader, per_layer):
    use_cache = model.config.use_cache
    model.config.use_cache = False
    model.eval()
    layers, inps, outs, layer_kwargs = _catch_inputs(model, device, _default_loader(model, n_samples, seq_len, dataloader))
    n_samples = inps.shape[0]
    for i in range(len(layers)):
        layer = layers[i].to(device)
        subset = solvers.find_layers(layer)

        def run_forwards():
            for j in range(n_samples):
                layer(inps[j].unsqueeze(0), **layer_kwargs)

        per_layer(subset, run_forwards)
        for j in range(n_samples):
            outs[j] = _first(layer(inps[j].unsqueeze(0), **layer_kwargs))
        layers[i] = layer.cpu()
        del layer
        inps, outs = outs, inps
    model.config.use_cache = use_cache


@torch.no_grad()
def sparsegpt(model, device, sparsity_ratio, n_samples=512, seq_len=2048, verbose=True, dataloader=None):
    def per_layer(subset, run_forwards):
        gpts = {name: solvers.Wrapper(subset[name], device) for name in subset}

        def add_batch(name):
            def tmp(_, inp, out):
                gpts[name].cache_hessian_weight(inp, out)
            return tmp

        handles = [subset[name].register_forward_hook(add_batch(name)) for name in subset]
        run_forwards()
        for h in handles:
            h.remove()
        for name in subset:
            solvers.prune_weight(layer=gpts[name], device=device, sparsity_ratio=sparsity_ratio, block_size=128,
                                 percdamp=0.01)
            subset[name].weight.data = gpts[name].module.weight.data
        gpts.clear()

    _prune_loop(model, device, n_samples, seq_len, dataloader, per_layer)


def _rownorm_layer(subset, run_forwards, device):
    for name in subset:
        columns = subset[name].weight.shape[1]
        subset[name].nsamples = 0
        subset[name].scaler_row = torch.zeros((columns), device=device)
    handles = [subset[name].register_forward_hook(solvers.cache_scalar_row) for name in subset]
    run_forwards()
    for h in handles:
        h.remove()


@torch.no_grad()
def wanda(model, device, sparsity_ratio, n_samples=512, seq_len=2048, verbose=True, dataloader=None):
    def per_layer(subset, run_forwards):
        _rownorm_layer(subset, run_forwards, device)
        for name in subset:
            solvers.wanda_prune_(subset[name], sparsity_ratio)
            del subset[name].nsamples, subset[name].scaler_row

    _prune_loop(model, device, n_samples, seq_len, dataloader, per_layer)


@torch.no_grad()
def ria(model, device, sparsity_ratio, alpha, n_samples=512, seq_len=2048, verbose=True, dataloader=None):
    def per_layer(subset, run_forwards):
        _rownorm_layer(subset, run_forwards, device)
        for name in subset:
            solvers.ria_prune_(subset[name], sparsity_ratio, alpha)
            del subset[name].nsamples, subset[name].scaler_row

    _prune_loop(model, device, n_samples, seq_len, dataloader, per_layer)


@torch.no_grad()
def magnitude(model, device, sparsity_ratio, verbose=True):
    use_cache = model.config.use_cache
    model.config.use_cache = False
    model.eval()
    layers = model.get_layers()
    for i in range(len(layers)):
        layer = layers[i].to(device)
        subset = solvers.find_layers(layer)
        for name in subset:
            solvers.magnitude_prune_(subset[name], sparsity_ratio)
        layers[i] = layer.cpu()
        del layer
    model.config.use_cache = use_cache

"""BASELINE INFRASTRUCTURE ONLY (used by bench.py's reference legs and by tests/test_reference_dropin_gpu.py).

Runs the UNMODIFIED reference driver `gptq()` (quantization/calibrations/gptq/core.py:21-160, imported through
oracle/ref_shim.py from /root/reference or oracle/_ref) on a one-block duck-typed model whose block holds ONE reference
`QLinear` of shape [N, K], with synthetic bf16 token activations, and measures the two reference functions of the hot
path where they run: the forward hook `cache_hessian_weight` (a closure inside gptq(), core.py:103-119 -- timed by
wrapping `nn.Module.register_forward_hook`) and `update_weight` (core.py:163-281 -- timed by wrapping the module
global the driver looks up).  No reference code is edited or re-typed; the harness only supplies what
`models/llama.py:232-258` supplies (get_layers / get_sequential / move_embed) and a synthetic `get_loaders`."""
import os
import sys
import time
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

VOCAB = 512


def available():
    return ref_shim.available()


class _Block(nn.Module):
    """A 'decoder layer' with a single Linear; returns its input so that the driver's outs[j] = layer(x)[0] type-checks."""

    def __init__(self, proj):
        super().__init__()
        self.proj = proj

    def forward(self, x, **kwargs):
        self.proj(x)
        return (x,)


def build_model(N, K, weight="int4-g[128]-rw", act_in=None, dtype=torch.bfloat16, seed=0, W=None, fake_quantizer=None):
    """`fake_quantizer`: a replacement for the reference's FakeQuantizer factory (quant.py:36-63) that the reference's own
    QLinear then calls (modules/qlinear.py:41-55) -- the first of the three swaps of INTEGRATION.md section 2."""
    ref_shim.install()
    import llm_compressor.quantization.calibrations.gptq.core as G  # noqa: F401  (installs the sys.path hack first)
    import llm_compressor.modules.qlinear as QL
    from llm_compressor.modules.qlinear import QLinear
    from llm_compressor.utils.parser import QuantConfigParser

    saved_fq = QL.FakeQuantizer
    if fake_quantizer is not None:
        QL.FakeQuantizer = fake_quantizer
    try:
        return _build_model(QLinear, QuantConfigParser, N, K, weight, act_in, dtype, seed, W)
    finally:
        QL.FakeQuantizer = saved_fq


def _build_model(QLinear, QuantConfigParser, N, K, weight, act_in, dtype, seed, W):

    qc = QuantConfigParser().build_cfg(weight, act_in, None, None)
    g = torch.Generator().manual_seed(seed)
    lin = nn.Linear(K, N, bias=False)
    lin.weight.data = (0.02 * torch.randn(N, K, generator=g)) if W is None else W.float()
    lin = lin.to(dtype)
    head = nn.Linear(8, 8, bias=False).to(dtype)
    m = nn.Module()
    m.embed = nn.Embedding(VOCAB, K)
    m.embed.weight.data = (torch.randn(VOCAB, K, generator=g) * torch.exp(0.7 * torch.randn(K, generator=g)))
    m.embed = m.embed.to(dtype)
    m.layers = nn.ModuleList([_Block(QLinear(linear=lin, quant_config=ref_shim.EasyDict(qc.linear), dtype=dtype, op_name="proj"))])
    m.lm_head = QLinear(linear=head, quant_config=ref_shim.EasyDict(qc.head), dtype=dtype, op_name="lm_head")
    m.config = ref_shim.EasyDict(use_cache=False, _name_or_path="duck")
    m.get_layers = types.MethodType(lambda self: self.layers, m)
    m.get_sequential = types.MethodType(lambda self, mode="true": [["proj"]], m)

    def move_embed(self, device):
        self.embed = self.embed.to(device)

    m.move_embed = types.MethodType(move_embed, m)
    m.forward = types.MethodType(lambda self, ids: self.layers[0](self.embed(ids)), m)
    return m


def run_rtn(N, K, device, weight="int4-g[128]-zp-rw", seed=0, W=None, fake_quantizer=None):
    """The reference's rtn() driver (rtn/core.py:16-61), unmodified, on the duck model; returns the model."""
    ref_shim.install()
    import llm_compressor.quantization.calibrations.gptq.core as G  # noqa: F401
    import llm_compressor.quantization.calibrations.rtn.core as RTN

    model = build_model(N, K, weight, seed=seed, W=W, fake_quantizer=fake_quantizer)
    with torch.no_grad():
        RTN.rtn(model, torch.device(device), False, False)
    return model


def run_gptq(N, K, n_samples, seq_len, device, weight="int4-g[128]-rw", seed=0, patch=None, W=None, fake_quantizer=None,
             hook_override=None):
    """One call of the reference's gptq() on the duck model.  Returns (model, hook seconds per call, update_weight seconds).
    The drop-in test swaps in this repository's objects, as INTEGRATION.md section 2 tells a maintainer to:
    `fake_quantizer` replaces the FakeQuantizer factory, `patch(G)` may replace reference globals (G.update_weight),
    `hook_override` stands in for the body of the forward-hook closure (gptq/core.py:103-119 is local to gptq())."""
    ref_shim.install()
    import llm_compressor.quantization.calibrations.gptq.core as G

    dev = torch.device(device)
    cuda = dev.type == "cuda"
    sync = torch.cuda.synchronize if cuda else (lambda: None)
    model = build_model(N, K, weight, seed=seed, W=W, fake_quantizer=fake_quantizer)
    loader = [(torch.randint(0, VOCAB, (1, seq_len), generator=torch.Generator().manual_seed(seed + i)), None)
              for i in range(n_samples)]
    saved = dict(get_loaders=G.get_loaders, update_weight=G.update_weight, reg=nn.Module.register_forward_hook)
    hook_s, upd_s = [], []
    G.get_loaders = lambda name, tokenizer_path, nsamples=128, seqlen=2048, seed=0: (loader, None)
    if patch is not None:
        patch(G)
    inner_update = G.update_weight

    def timed_update(*a, **k):
        sync()
        t0 = time.perf_counter()
        r = inner_update(*a, **k)
        sync()
        upd_s.append(time.perf_counter() - t0)
        return r

    def reg(self, hook, *a, **k):
        if hook_override is not None:
            hook = hook_override

        def timed_hook(mod, x, y):
            sync()
            t0 = time.perf_counter()
            r = hook(mod, x, y)
            sync()
            hook_s.append(time.perf_counter() - t0)
            return r
        return saved["reg"](self, timed_hook, *a, **k)

    G.update_weight = timed_update
    nn.Module.register_forward_hook = reg
    try:
        with torch.no_grad():
            G.gptq(model, dev, n_samples, seq_len, False, False)
    finally:
        nn.Module.register_forward_hook = saved["reg"]
        G.update_weight = saved["update_weight"]
        G.get_loaders = saved["get_loaders"]
    return model, hook_s, (upd_s[0] if upd_s else float("nan"))

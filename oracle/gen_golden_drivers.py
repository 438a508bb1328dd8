"""TEST INFRASTRUCTURE ONLY -- golden vectors for the layer-wise DRIVERS (L2 loops).

Runs the UNMODIFIED reference drivers (rtn / gptq / gptaq / sparsegpt / wanda / ria / magnitude, imported from
/root/reference through oracle/ref_shim.py) on a tiny random-init Llama (2 layers, d = 128) on CPU, with the
harness of SURVEY.md section 8c: plain HF `LlamaForCausalLM` + the reference's own `QLinear` swapped in
(restating `_prepare_qmodule`, models/llama.py:177-208) + `get_layers / get_sequential / move_embed`
(models/llama.py:232-258) + a synthetic `get_loaders`.  Stores the initial weights, the calibration tokens
and every compressed Linear weight (bf16 bit patterns) in tests/golden/drivers.npz.

    python oracle/gen_golden_drivers.py        (build container only)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

VOCAB, D, FFN, LAYERS, HEADS, KV = 256, 128, 256, 2, 4, 2
NSAMPLES, SEQLEN = 16, 64


def tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(vocab_size=VOCAB, hidden_size=D, intermediate_size=FFN, num_hidden_layers=LAYERS,
                      num_attention_heads=HEADS, num_key_value_heads=KV, max_position_embeddings=SEQLEN,
                      tie_word_embeddings=False, attn_implementation="eager")
    torch.manual_seed(0)
    m = LlamaForCausalLM(cfg).to(torch.bfloat16)
    # give the activations per-channel outliers so that act-order / Wanda scores are not degenerate
    g = torch.Generator().manual_seed(1)
    m.model.embed_tokens.weight.data *= torch.exp(0.7 * torch.randn(D, generator=g)).to(torch.bfloat16)
    return m


def loader(seed=0):
    return [(torch.randint(0, VOCAB, (1, SEQLEN), generator=torch.Generator().manual_seed(seed + i)), None)
            for i in range(NSAMPLES)]


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def wrap_reference(model, weight, act_in=None):
    """Linear -> the reference's QLinear; duck-type methods (ref: models/llama.py:177-258)."""
    import types

    from llm_compressor.modules.qlinear import QLinear
    from llm_compressor.utils.parser import QuantConfigParser
    from torch import nn

    qc = QuantConfigParser().build_cfg(weight, act_in, None, None)
    mods = dict(model.named_modules())
    for name, module in list(mods.items()):
        if isinstance(module, nn.Linear):
            cfg = ref_shim.EasyDict(qc.head if "lm_head" in name else qc.linear)
            q = QLinear(linear=module, quant_config=cfg, dtype=torch.bfloat16, op_name=name.replace("model.", ""))
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(mods[parent], child, q)
            else:
                setattr(model, name, q)
    model.get_layers = types.MethodType(lambda self: self.model.layers, model)
    model.get_sequential = types.MethodType(lambda self, mode="true": [
        ["self_attn.k_proj", "self_attn.v_proj", "self_attn.q_proj"], ["self_attn.o_proj"],
        ["mlp.up_proj", "mlp.gate_proj"], ["mlp.down_proj"]], model)

    def move_embed(self, device):
        self.model.embed_tokens = self.model.embed_tokens.to(device)
        self.model.rotary_emb = self.model.rotary_emb.to(device)

    model.move_embed = types.MethodType(move_embed, model)
    return model


def linear_weights(model):
    return {n: bits(p) for n, p in model.state_dict().items() if n.endswith("proj.weight")}


def main():
    # gptq first: it installs the reference's `sys.path.append(<llm_compressor dir>)` hack (gptq/core.py:12-19)
    import llm_compressor.quantization.calibrations.gptq.core as G  # isort: skip
    import llm_compressor.pruning.magnitude.core as MAG
    import llm_compressor.pruning.ria.core as RIA
    import llm_compressor.pruning.sparsegpt.core as SGP
    import llm_compressor.pruning.wanda.core as WAN
    import llm_compressor.quantization.calibrations.gptaq.core as GA
    import llm_compressor.quantization.calibrations.rtn.core as RTN

    fake = lambda name, tokenizer_path, nsamples=128, seqlen=2048, seed=0: (loader(seed), None)  # noqa: E731
    for mod in (G, GA, SGP, WAN, RIA):
        mod.get_loaders = fake
    cpu = torch.device("cpu")
    out = {}
    base = tiny_llama()
    for n, p in base.state_dict().items():
        out["init/" + n] = bits(p) if p.dtype == torch.bfloat16 else p.numpy()
    runs = [
        ("rtn_int4_g128_zp", "int4-g[128]-zp-rw", None, lambda m: RTN.rtn(m, cpu, mse=False, verbose=False)),
        ("gptq_int4_g128", "int4-g[128]-rw", None, lambda m: G.gptq(m, cpu, NSAMPLES, SEQLEN, False, False)),
        ("gptq_nvfp4_g16", "nvfp4_e2m1-g[16]-rw", None, lambda m: G.gptq(m, cpu, NSAMPLES, SEQLEN, False, False)),
        ("gptaq_int4_g128_a8", "int4-g[128]-rw", "int8-g[-1]-rw", lambda m: GA.gptaq(m, cpu, NSAMPLES, SEQLEN, False, False)),
        ("sparsegpt_50", None, None, lambda m: SGP.sparsegpt(m, cpu, 0.5, NSAMPLES, SEQLEN, False)),
        ("wanda_50", None, None, lambda m: WAN.wanda(m, cpu, 0.5, NSAMPLES, SEQLEN, False)),
        ("ria_50", None, None, lambda m: RIA.ria(m, cpu, 0.5, 0.5, NSAMPLES, SEQLEN, False)),
        ("magnitude_50", None, None, lambda m: MAG.magnitude(m, cpu, 0.5, False)),
    ]
    meta = []
    for name, w, a, fn in runs:
        m = wrap_reference(tiny_llama(), w, a)
        m.config._name_or_path = "tiny-llama"
        fn(m)
        ws = linear_weights(m)
        for n, v in ws.items():
            out[name + "/" + n] = v
        meta.append((name, w, a))
        print(name, "done", len(ws), "tensors")
    out["__meta__"] = np.array(repr(dict(runs=meta, vocab=VOCAB, d=D, ffn=FFN, layers=LAYERS, heads=HEADS, kv=KV,
                                         nsamples=NSAMPLES, seqlen=SEQLEN)))
    np.savez_compressed(os.path.join(GOLD, "drivers.npz"), **out)
    print("wrote", os.path.join(GOLD, "drivers.npz"))


if __name__ == "__main__":
    main()

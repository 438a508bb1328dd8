"""Model duck-type the layer-wise drivers expect, attached to a plain HF causal LM.

The reference subclasses the HF models (`CompressLlamaForCausalLM` ... ref: models/llama.py:143-258,
models/opt.py) to (1) swap every `nn.Linear` for a `QLinear` (`_prepare_qmodule`, llama.py:177-208) and
(2) expose `get_layers / get_sequential / move_embed` (llama.py:232-258).  This helper does the same two
things to an already-built `transformers` model, so `llm_compressor_b200.drivers.*` (or the reference's own
drivers) can run on it.  Host-side glue only; no numerics here.
"""
import types

from torch import nn

from .config import build_quant_config
from .modules import QLinear

_LLAMA_SEQ = [
    ["self_attn.k_proj", "self_attn.v_proj", "self_attn.q_proj"],
    ["self_attn.o_proj"],
    ["mlp.up_proj", "mlp.gate_proj"],
    ["mlp.down_proj"],
]
_OPT_SEQ = [
    ["self_attn.k_proj", "self_attn.v_proj", "self_attn.q_proj"],
    ["self_attn.out_proj"],
    ["fc1"],
    ["fc2"],
]


def _is_opt(model):
    return hasattr(model, "model") and hasattr(model.model, "decoder")


def swap_linears(model, quant_config, qlinear_cls=QLinear, dtype=None):
    """Linear -> QLinear everywhere; `lm_head` gets the head config (ref: models/llama.py:177-208)."""
    dtype = dtype if dtype is not None else next(model.parameters()).dtype
    mods = dict(model.named_modules())
    for name, module in list(mods.items()):
        if not isinstance(module, nn.Linear) or isinstance(module, qlinear_cls):
            continue
        op_name = name.replace("model.", "")
        cfg = quant_config["head"] if "lm_head" in name else quant_config["linear"]
        q = qlinear_cls(linear=module, quant_config=cfg, dtype=dtype, op_name=op_name)
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(mods[parent], child, q)
        else:
            setattr(model, name, q)
    return model


def attach_duck_type(model):
    """get_layers / get_sequential / move_embed for Llama-like and OPT-like HF models."""
    if _is_opt(model):
        dec = model.model.decoder

        def get_layers(self):
            return dec.layers

        def move_embed(self, device):
            dec.embed_tokens = dec.embed_tokens.to(device)
            dec.embed_positions = dec.embed_positions.to(device)

        seq = _OPT_SEQ
    else:
        def get_layers(self):
            return self.model.layers

        def move_embed(self, device):
            self.model.embed_tokens = self.model.embed_tokens.to(device)
            self.model.rotary_emb = self.model.rotary_emb.to(device)

        seq = _LLAMA_SEQ

    def get_sequential(self, mode="true"):
        return [list(g) for g in seq] if mode == "true" else [[n for g in seq for n in g]]

    model.get_layers = types.MethodType(get_layers, model)
    model.get_sequential = types.MethodType(get_sequential, model)
    model.move_embed = types.MethodType(move_embed, model)
    return model


def prepare(model, weight, act_in=None, act_out=None, head=None):
    """`prepare(model, "int4-g[128]-rw")`: config strings of the reference CLI -> wrapped model."""
    return attach_duck_type(swap_linears(model, build_quant_config(weight, act_in, act_out, head)))

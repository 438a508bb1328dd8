"""Host-side glue that needs no GPU: the Linear -> QLinear swap and the model duck-type of adapters.py
(ref: models/llama.py:177-258, models/opt.py:242-266), the config-string grammar feeding it, and the generated
Hadamard blocks' orthogonality."""
import pytest
import torch


def _tiny(kind):
    if kind == "llama":
        from transformers import LlamaConfig, LlamaForCausalLM
        return LlamaForCausalLM(LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                            num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=32,
                                            tie_word_embeddings=False)).to(torch.bfloat16)
    from transformers import OPTConfig, OPTForCausalLM
    return OPTForCausalLM(OPTConfig(vocab_size=64, hidden_size=32, ffn_dim=64, num_hidden_layers=2, num_attention_heads=4,
                                    max_position_embeddings=32, word_embed_proj_dim=32)).to(torch.bfloat16)


@pytest.mark.parametrize("kind", ["llama", "opt"])
def test_prepare_swaps_every_linear_and_attaches_the_duck_type(kind):
    from llm_compressor_b200 import adapters
    from llm_compressor_b200.modules import QLinear
    from llm_compressor_b200.quantizers import DummyQuantizer, INTQuantizer
    m = _tiny(kind)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    adapters.prepare(m, "int4-g[128]-zp-rw", act_in="int8-g[-1]-rw")
    lin = [mod for mod in m.modules() if isinstance(mod, torch.nn.Linear)]
    assert lin and all(isinstance(mod, QLinear) for mod in lin)
    assert all(torch.equal(v, before[k]) for k, v in m.state_dict().items() if k in before)     # weights untouched
    body = [mod for name, mod in m.named_modules() if isinstance(mod, QLinear) and "lm_head" not in name]
    assert all(isinstance(q.weight_quantizer, INTQuantizer) and q.weight_quantizer.zero_point for q in body)
    assert all(isinstance(q.input_quantizer, INTQuantizer) and q.input_quantizer.group_size == -1 for q in body)
    assert isinstance(m.lm_head.weight_quantizer, DummyQuantizer)                               # head config None
    layers = m.get_layers()
    assert len(layers) == 2
    seq = m.get_sequential("true")
    from llm_compressor_b200.solvers import find_layers
    full = find_layers(layers[0])
    assert sorted(n for g in seq for n in g) == sorted(full)                                    # every Linear in one group
    assert m.get_sequential("false") == [[n for g in seq for n in g]]
    m.move_embed("cpu")


def test_quantized_forward_refuses_cpu_tensors():
    """No CPU fallback: the swapped model cannot run a quantised forward without the CUDA library's device."""
    from llm_compressor_b200 import _lib, adapters
    m = adapters.prepare(_tiny("llama"), "int4-g[128]-rw", act_in="int8-g[-1]-rw")   # activation QDQ in every forward
    with pytest.raises(_lib.LcbError):
        m(torch.zeros(1, 8, dtype=torch.long))


@pytest.mark.parametrize("K", [12, 20, 28, 36, 40, 44, 60])
def test_generated_hadamard_blocks_are_orthogonal(K):
    from llm_compressor_b200 import hadamard as H
    h = H._had(K)
    assert set(h.unique().tolist()) == {-1.0, 1.0}
    assert torch.equal(h @ h.T, K * torch.eye(K))

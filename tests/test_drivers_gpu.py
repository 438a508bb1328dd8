"""Driver-level (L2 loop) parity on the GPU: `llm_compressor_b200.drivers.*` on a tiny random-init Llama against
the UNMODIFIED reference drivers run on CPU (tests/golden/drivers.npz, written by oracle/gen_golden_drivers.py).

RTN and magnitude never look at activations -> bit-exact.  The calibrated methods see activations produced by
bf16 forwards on a different device (cuBLAS vs CPU GEMM rounding), so the Hessians / row norms differ in the last
bf16 bits: masks and integer codes are compared by agreement rate, with the measured values in the comments.
"""
import numpy as np
import pytest
import torch

from golden_io import load
from util import t_from_bits

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
Z, META = load("drivers")
RUNS = {r[0]: r for r in META["runs"]}


def _tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(vocab_size=META["vocab"], hidden_size=META["d"], intermediate_size=META["ffn"],
                      num_hidden_layers=META["layers"], num_attention_heads=META["heads"],
                      num_key_value_heads=META["kv"], max_position_embeddings=META["seqlen"],
                      tie_word_embeddings=False, attn_implementation="eager")
    m = LlamaForCausalLM(cfg).to(torch.bfloat16)
    sd = {k[len("init/"):]: t_from_bits(Z[k]) for k in Z.files if k.startswith("init/")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "rotary" not in k], missing
    return m


def _loader():
    from llm_compressor_b200.drivers import synthetic_loader

    return synthetic_loader(META["vocab"], META["nsamples"], META["seqlen"], seed=0)


def _run(name):
    from llm_compressor_b200 import adapters, drivers

    _, w, a = RUNS[name]
    m = adapters.prepare(_tiny_llama(), w, a)
    n, s = META["nsamples"], META["seqlen"]
    kind = name.split("_")[0]
    if kind == "rtn":
        drivers.rtn(m, DEV, mse=False, verbose=False)
    elif kind == "gptq":
        drivers.gptq(m, DEV, n, s, False, False, dataloader=_loader())
    elif kind == "gptaq":
        drivers.gptaq(m, DEV, n, s, False, False, dataloader=_loader())
    elif kind == "sparsegpt":
        drivers.sparsegpt(m, DEV, 0.5, n, s, False, dataloader=_loader())
    elif kind == "wanda":
        drivers.wanda(m, DEV, 0.5, n, s, False, dataloader=_loader())
    elif kind == "ria":
        drivers.ria(m, DEV, 0.5, 0.5, n, s, False, dataloader=_loader())
    elif kind == "magnitude":
        drivers.magnitude(m, DEV, 0.5, False)
    got = {k: v.detach().float().cpu().numpy() for k, v in m.state_dict().items() if k.endswith("proj.weight")}
    ref = {k[len(name) + 1:]: t_from_bits(Z[k]).float().numpy() for k in Z.files if k.startswith(name + "/")}
    assert set(got) == set(ref) and len(ref) == 7 * META["layers"]
    return got, ref


@pytest.mark.parametrize("name", ["rtn_int4_g128_zp", "magnitude_50"])
def test_data_free_drivers_bit_exact(name):
    got, ref = _run(name)
    for k in ref:
        assert np.array_equal(got[k], ref[k]), k
    if name.startswith("magnitude"):
        assert all(abs(float((v == 0).mean()) - 0.5) < 0.02 for v in got.values())


@pytest.mark.parametrize("name", ["wanda_50", "ria_50", "sparsegpt_50"])
def test_pruning_drivers_mask_agreement(name):
    got, ref = _run(name)
    worst = 1.0
    for k in ref:
        mg, mr = got[k] == 0, ref[k] == 0
        worst = min(worst, float((mg == mr).mean()))
        if name.startswith("wanda"):  # exactly int(K * 0.5) weights pruned in every row (wanda/core.py:121-125)
            assert np.all(mg.sum(1) >= mg.shape[1] // 2)
        assert abs(float(mg.mean()) - float(mr.mean())) < 5e-3, k
    print(f"{name}: worst per-Linear mask agreement with the reference {worst:.4f}")
    assert worst > 0.995  # measured on B200: wanda 0.9999, ria 1.0000, sparsegpt 0.9993


@pytest.mark.parametrize("name", ["gptq_int4_g128", "gptq_nvfp4_g16", "gptaq_int4_g128_a8"])
def test_gptq_drivers_vs_reference(name):
    got, ref = _run(name)
    init = {k[len("init/"):]: t_from_bits(Z[k]).float().numpy() for k in Z.files if k.startswith("init/")}
    worst_same, worst_ratio = 1.0, 0.0
    for k in ref:
        same = float((got[k] == ref[k]).mean())
        eg = float(np.linalg.norm(got[k] - init[k]))
        er = float(np.linalg.norm(ref[k] - init[k]))
        worst_same = min(worst_same, same)
        worst_ratio = max(worst_ratio, eg / er)
        # the quantisation perturbation has the same size as the reference's
        assert 0.9 < eg / er < 1.1, (k, eg, er)
    print(f"{name}: worst fraction of identical weights {worst_same:.4f}, worst |dW| ratio {worst_ratio:.4f}")
    # layer 0's q/k/v see bit-identical inputs (embedding rows through one RMSNorm): near-identical codes
    k0 = "model.layers.0.self_attn.q_proj.weight"
    assert float((got[k0] == ref[k0]).mean()) > 0.98
    assert worst_same > 0.98  # measured on B200: int4-g128 0.9995, nvfp4-g16 0.9913, gptaq int4 + a8 0.9978

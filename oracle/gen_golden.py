"""TEST INFRASTRUCTURE ONLY -- golden-vector generator.

Runs the UNMODIFIED reference (imported from /root/reference through oracle/ref_shim.py) on
seeded synthetic inputs and stores inputs + outputs as small fixtures under tests/golden/.
Run in the build container only:

    python oracle/gen_golden.py

The fixtures are what pins the oracle (and through it the CUDA path) to the reference; the
reference has no golden vectors of its own.
bf16 tensors are stored as their uint16 bit patterns.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def to_bits(t):
    """torch tensor -> numpy; bf16 as uint16 bit pattern, fp32 as float32."""
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def make_input(shape, dtype, kind, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    if kind == "weight":
        x = 0.02 * x
    elif kind == "act":
        x = x * torch.exp(1.5 * torch.randn(shape[-1], generator=g))  # per-channel outliers
    elif kind == "edge":
        x = 0.5 * x
        flat = x.reshape(-1)
        flat[::7] = 0.0                       # exact zeros
        flat[3] = 7.96875; flat[5] = -255.0   # bf16 log2 rounding quirk inputs
        flat[11] = 1e-30; flat[13] = -3e4; flat[17] = 0.2490234375; flat[19] = 1.9921875
        last = x.shape[-1]
        x.reshape(-1, last)[1, :] = 0.0       # an all-zero row
        x.reshape(-1, last)[2, :] = 1.25      # an all-equal row
    return x.to(dtype)


QDQ_CASES = []


def _add(name, cfg, shape, dtype, kind, seed):
    QDQ_CASES.append(dict(name=name, cfg=cfg, shape=shape, dtype=dtype, kind=kind, seed=seed))


def _cfg(t, f, g, axes=-1, zp=False):
    return dict(type=t, format=f, group_size=g, axes=axes, zero_point=zp, is_profile=False)


def build_cases():
    seed = 100
    for dt in (torch.bfloat16, torch.float32):
        dn = "bf16" if dt == torch.bfloat16 else "f32"
        base = [
            ("int4_g128_zp", _cfg("int", "int4", 128, zp=True), (8, 384)),
            ("int4_g128", _cfg("int", "int4", 128), (8, 384)),
            ("int8_g128", _cfg("int", "int8", 128), (1, 12, 256)),
            ("int8_g128_zp", _cfg("int", "int8", 128, zp=True), (1, 12, 256)),
            ("int4_tok", _cfg("int", "int4", -1), (1, 9, 320)),
            ("int4_tok_zp", _cfg("int", "int4", -1, zp=True), (1, 9, 320)),
            ("int8_chan", _cfg("int", "int8", -2), (40, 72)),
            ("int8_chan_zp", _cfg("int", "int8", -2, zp=True), (40, 72)),
            ("int8_tensor", _cfg("int", "int8", 0), (20, 96)),
            ("int8_tensor_zp", _cfg("int", "int8", 0, zp=True), (20, 96)),
            ("int4_g128_cw", _cfg("int", "int4", 128, axes=-2), (1, 2, 256, 24)),
            ("int8_g128_ragged", _cfg("int", "int8", 128, zp=True), (6, 200)),
            ("int8_g128_short", _cfg("int", "int8", 128, zp=True), (1, 2, 9, 64)),
            ("int4_g128_cw_ragged", _cfg("int", "int4", 128, axes=-2, zp=True), (1, 2, 200, 16)),
            ("fp8e4m3_tok", _cfg("fp", "fp8_e4m3", -1), (1, 9, 320)),
            ("fp8e4m3_tok_zp", _cfg("fp", "fp8_e4m3", -1, zp=True), (1, 9, 320)),
            ("fp8e5m2_g128", _cfg("fp", "fp8_e5m2", 128), (8, 384)),
            ("fp4_g128_zp", _cfg("fp", "fp4_e2m1", 128, zp=True), (8, 384)),
            ("fp4_g128", _cfg("fp", "fp4_e2m1", 128), (8, 384)),
            ("fp8e4m3_tensor", _cfg("fp", "fp8_e4m3", 0), (20, 96)),
            ("fp8e4m3_chan", _cfg("fp", "fp8_e4m3", -2), (40, 72)),
            ("mxfp4_g32", _cfg("mx", "fp4_e2m1", 32), (8, 384)),
            ("mxfp4_g32_zp", _cfg("mx", "fp4_e2m1", 32, zp=True), (8, 384)),
            ("mxfp8e4m3_g32", _cfg("mx", "fp8_e4m3", 32), (1, 12, 256)),
            ("mxfp8e5m2_g32", _cfg("mx", "fp8_e5m2", 32), (1, 12, 256)),
            ("mxint8_g32", _cfg("mx", "int8", 32), (8, 384)),
            ("mxint4_g32", _cfg("mx", "int4", 32), (8, 384)),
            ("mxfp4_g32_cw", _cfg("mx", "fp4_e2m1", 32, axes=-2), (1, 2, 96, 24)),
            ("mxfp4_g32_ragged", _cfg("mx", "fp4_e2m1", 32), (6, 200)),
            ("nvfp4_g16", _cfg("nvfp", "fp4_e2m1", 16), (8, 384)),
            ("nvfp4_g16_zp", _cfg("nvfp", "fp4_e2m1", 16, zp=True), (8, 384)),
            ("nvfp4_g16_act", _cfg("nvfp", "fp4_e2m1", 16), (1, 12, 256)),
            ("nvfp4_g16_cw", _cfg("nvfp", "fp4_e2m1", 16, axes=-2), (1, 2, 96, 24)),
            ("nvfp4_g16_ragged", _cfg("nvfp", "fp4_e2m1", 16), (6, 200)),
        ]
        for name, cfg, shape in base:
            for kind in ("weight", "act", "edge"):
                if kind == "edge" and cfg["type"] == "nvfp":
                    pass  # all-zero rows are fine for nvfp (global amax non-zero)
                seed += 1
                _add(f"{name}.{dn}.{kind}", cfg, shape, dt, kind, seed)


def gen_qdq():
    build_cases()
    out = {}
    meta = []
    for c in QDQ_CASES:
        x = make_input(c["shape"], c["dtype"], c["kind"], c["seed"])
        q = ref_shim.build_quantizer(c["cfg"])
        with torch.no_grad():
            try:
                s, z = ref_shim.build_quantizer(c["cfg"]).find_params(x.clone())
                nan_assert = False
            except AssertionError:
                nan_assert = True
            if nan_assert:
                continue  # reference refuses (NaN scales); covered by host-side error tests
            y = q(x.clone())
        n = c["name"]
        out[n + "/x"] = to_bits(x)
        out[n + "/y"] = to_bits(y)
        out[n + "/s"] = to_bits(s)
        out[n + "/z"] = to_bits(z)
        meta.append((n, c["cfg"], list(c["shape"]), "bf16" if c["dtype"] == torch.bfloat16 else "f32"))
    out["__meta__"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, "qdq.npz"), **out)
    print("qdq cases:", len(meta))


def gen_elem_core():
    """Exhaustive: every bf16 bit pattern through _quantize_elemwise_core for each float format."""
    ref_shim.install()
    from llm_compressor.quantization.quantizers.utils import _quantize_elemwise_core
    from llm_compressor.quantization.quantizers.formats import ElemFormat, _get_format_params

    bits = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16)
    a = bits.view(torch.bfloat16)
    out = {}
    for fmt in ("fp4_e2m1", "fp8_e4m3", "fp8_e5m2", "int4", "int8"):
        ebits, mbits, emax, max_norm, _ = _get_format_params(ElemFormat.from_str(fmt))
        y = _quantize_elemwise_core(a.clone(), torch.tensor(mbits), torch.tensor(ebits), torch.tensor(max_norm),
                                    round="nearest", allow_denorm=True, saturate_normals=True)
        out[fmt] = to_bits(y)
    # fp32: adversarial values around powers of two and rounding ties
    g = torch.Generator().manual_seed(7)
    base = torch.cat([
        torch.randn(4096, generator=g) * 4,
        torch.randn(4096, generator=g) * 1e-3,
        (2.0 ** torch.arange(-20, 12).float()).repeat_interleave(4) * torch.tensor([1.0, 1 - 2 ** -24, 1 + 2 ** -23, 1.5]).repeat(32),
        torch.arange(0, 64).float() * 0.25 + 0.125,
    ])
    out["f32_in"] = base.numpy()
    for fmt in ("fp4_e2m1", "fp8_e4m3", "fp8_e5m2", "int4", "int8"):
        ebits, mbits, emax, max_norm, _ = _get_format_params(ElemFormat.from_str(fmt))
        y = _quantize_elemwise_core(base.clone(), torch.tensor(mbits), torch.tensor(ebits), torch.tensor(max_norm),
                                    round="nearest", allow_denorm=True, saturate_normals=True)
        out["f32_" + fmt] = y.numpy()
    np.savez_compressed(os.path.join(GOLD, "elem_core.npz"), **out)
    print("elem core done")


class _Layer:
    pass


def _make_qlinear(W, cfg):
    ref_shim.install()
    from llm_compressor.modules.qlinear import QLinear

    N, K = W.shape
    lin = torch.nn.Linear(K, N, bias=False, dtype=W.dtype)
    lin.weight.data = W.clone()
    none = dict(type=None, is_profile=False)
    qc = ref_shim.EasyDict(weight=dict(cfg), act_in=none, act_out=none)
    return QLinear(lin, qc, W.dtype)


def _hessian_from(X_list, dev="cpu", fp_list=None):
    import math
    K = X_list[0].shape[-1]
    H = torch.zeros(K, K)
    D = torch.zeros(K, K) if fp_list is not None else None
    n = 0
    for j, x in enumerate(X_list):
        H *= n / (n + 1)
        if D is not None:
            D *= n / (n + 1)
        n += 1
        inp = math.sqrt(2 / n) * x.float().t()
        H += inp.matmul(inp.t())
        if D is not None:
            dX = math.sqrt(2 / n) * fp_list[j].float().t() - inp
            D += dX.matmul(inp.t())
    return H, D


def gen_solvers():
    ref_shim.install()
    import llm_compressor.quantization.calibrations.gptq.core as G
    import llm_compressor.quantization.calibrations.gptaq.core as GA
    import llm_compressor.pruning.sparsegpt.core as SG

    out = {}
    meta = []
    N, K, T, ns = 48, 256, 192, 3
    cases = [
        ("gptq_int4_g128", _cfg("int", "int4", 128), "gptq"),
        ("gptq_int4_g128_zp", _cfg("int", "int4", 128, zp=True), "gptq"),
        ("gptq_int4_row", _cfg("int", "int4", -1, zp=True), "gptq"),
        ("gptq_mxfp4_g32", _cfg("mx", "fp4_e2m1", 32), "gptq"),
        ("gptq_nvfp4_g16", _cfg("nvfp", "fp4_e2m1", 16), "gptq"),
        ("gptaq_int4_g128", _cfg("int", "int4", 128), "gptaq"),
        ("gptaq_int4_row", _cfg("int", "int4", -1), "gptaq"),
        ("sparsegpt_50", None, "sparsegpt"),
    ]
    g = torch.Generator().manual_seed(499)
    chan = torch.exp(0.8 * torch.randn(K, generator=g))
    Xs = [(torch.randn(T, K, generator=g) * chan).to(torch.bfloat16) for _ in range(ns)]
    for x in Xs:
        x[:, 5] = 0  # dead input channel -> diag(H) == 0
    Xfp = [(x.float() + 0.05 * torch.randn(T, K, generator=g) * chan).to(torch.bfloat16) for x in Xs]
    for x in Xfp:
        x[:, 5] = 0
    out["X"] = to_bits(torch.stack(Xs))
    out["Xfp"] = to_bits(torch.stack(Xfp))
    H0, D0 = _hessian_from(Xs, fp_list=Xfp)
    out["H"] = H0.numpy().copy()
    out["dXXT"] = D0.numpy().copy()
    for idx, (name, cfg, kind) in enumerate(cases):
        g = torch.Generator().manual_seed(500 + idx)
        W = (0.05 * torch.randn(N, K, generator=g)).to(torch.bfloat16)
        W[:, 7] = 0  # column of zeros in W (MASK path)
        out[name + "/W"] = to_bits(W)
        H, D = H0.clone(), (D0.clone() if kind == "gptaq" else None)
        with torch.no_grad():
            if kind == "sparsegpt":
                lay = _Layer()
                lay.module = torch.nn.Linear(K, N, bias=False, dtype=torch.bfloat16)
                lay.module.weight.data = W.clone()
                lay.H = H.clone()
                SG.prune_weight(lay, "cpu", 0.5, block_size=128, percdamp=0.01)
                Wn = lay.module.weight.data
            else:
                ql = _make_qlinear(W, cfg)
                ql.weight_quantizer.H = H.clone()
                if kind == "gptaq":
                    ql.weight_quantizer.dXXT = D.clone()
                    GA.update_weight(ql, "cpu", block_size=128, percdamp=0.01, actorder=True, alpha=0.25)
                else:
                    G.update_weight(ql, "cpu", block_size=128, percdamp=0.01, actorder=True)
                Wn = ql.weight.data
        out[name + "/Wnew"] = to_bits(Wn)
        meta.append((name, cfg, kind))
    out["__meta__"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, "solvers.npz"), **out)
    print("solver cases:", len(meta))


def gen_masks():
    out = {}
    g = torch.Generator().manual_seed(900)
    N, K, T = 40, 192, 96
    W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16)
    W[3, :10] = W[3, 10:20]  # ties inside a row
    W[:, 50] = 0
    out["W"] = to_bits(W)
    # row norms exactly as the hook does (wanda/core.py:92-105)
    s = torch.zeros(K)
    n = 0
    Xs = []
    for _ in range(3):
        x = (torch.randn(T, K, generator=g) * torch.exp(torch.randn(K, generator=g))).to(torch.bfloat16)
        Xs.append(x)
        xt = x.t()
        s *= n / (n + 1)
        n += 1
        s += torch.norm(xt.float(), p=2, dim=1) ** 2 / n
    out["X"] = to_bits(torch.stack(Xs))
    out["scaler_row"] = s.numpy().copy()
    for ratio in (0.3, 0.5):
        tag = str(int(ratio * 100))
        Wm = torch.abs(W) * torch.sqrt(s.reshape((1, -1)))
        mask = torch.zeros_like(Wm) == 1
        idx = torch.sort(Wm, dim=-1, stable=True)[1][:, : int(Wm.shape[1] * ratio)]
        mask.scatter_(1, idx, True)
        out["wanda_" + tag] = mask.numpy()
        for alpha in (0.5, 1.0):
            Wr = (torch.abs(W) / torch.sum(torch.abs(W), dim=0)
                  + torch.abs(W) / torch.sum(torch.abs(W), dim=1).reshape(-1, 1)) * (torch.sqrt(s.reshape((1, -1)))) ** alpha
            th = torch.sort(Wr.flatten())[0][int(W.numel() * ratio)]
            out[f"ria_{tag}_{alpha}"] = (Wr <= th).numpy()
        Wa = torch.abs(W)
        th = torch.sort(Wa.flatten())[0][int(W.numel() * ratio)]
        out["magnitude_" + tag] = (Wa <= th).numpy()
    np.savez_compressed(os.path.join(GOLD, "masks.npz"), **out)
    print("masks done")


if __name__ == "__main__":
    torch.set_num_threads(4)
    os.makedirs(GOLD, exist_ok=True)
    gen_elem_core()
    gen_qdq()
    gen_masks()
    gen_solvers()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))

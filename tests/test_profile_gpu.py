"""f4 (second half): the numerical profile `record_stats` computed on the GPU -- two reduction passes + an exact order
statistic by radix select -- against (1) a plain fp32 computation with a full sort on the CPU and (2) the reference's own
record_stats (quantizers/base.py:30-113, from oracle/_ref) through its CSV line, on fp32 tensors (for bf16 the reference
does its normalisation arithmetic in bf16 on the CPU; the columns then agree to bf16 precision only)."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True, is_profile=True)


def _cpu_stats(x, q):
    x, q = x.float().cpu(), q.float().cpu()
    k = round(0.99 * (x.numel() - 1))
    pc99 = torch.sort(x.flatten())[0][k].item()
    t = (x - x.min()) / (x.max() - x.min())
    u = (q - q.min()) / (q.max() - q.min())
    sqnr = (-10 * torch.log10(torch.mean((t.double() - u.double()) ** 2) + 1e-10)).item()
    return pc99, x.max().item(), q.max().item(), sqnr, x.max().item() - q.max().item()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(4, 300, 1024), (1, 2048, 3072), (7, 129)])
def test_device_stats_match_a_cpu_sort(dtype, shape):
    import llm_compressor_b200 as lc
    from llm_compressor_b200 import profile
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * torch.exp(0.5 * torch.randn(shape[-1], generator=g))).to(dtype).to(DEV)
    q = lc.FakeQuantizer.build(dict(CFG, is_profile=False, group_size=-1 if shape[-1] % 128 else 128)).to(DEV)(x)
    got = profile.device_stats(x, q)
    ref = _cpu_stats(x, q)
    assert got[0] == ref[0] and got[1] == ref[1] and got[2] == ref[2]          # order statistics and maxima: exact
    assert abs(got[3] - ref[3]) < 1e-3 and abs(got[4] - ref[4]) < 1e-6


@pytest.mark.skipif(not ref_shim.available(), reason="no reference copy (oracle/_ref) on this box")
def test_stats_csv_line_matches_the_reference(tmp_path):
    import llm_compressor_b200 as lc
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(1, 512, 1024, generator=g) * torch.exp(0.5 * torch.randn(1024, generator=g)))
    d_ref, d_our = tmp_path / "ref", tmp_path / "ours"
    d_ref.mkdir(); d_our.mkdir()
    ref_shim.install()
    from llm_compressor.quantization.quant import FakeQuantizer as RefFQ
    RefFQ.build(dict(CFG), op_name="layer0.q_proj.input", save_path=str(d_ref))(x)
    lc.FakeQuantizer.build(dict(CFG), op_name="layer0.q_proj.input", save_path=str(d_our)).to(DEV)(x.to(DEV))
    files = [next(p for p in d.iterdir() if p.suffix == ".csv") for d in (d_ref, d_our)]
    a, b = (f.read_text().strip().splitlines() for f in files)
    assert a[0] == b[0]                                           # identical header line (format of base.py:106-113)
    ra, rb = [c.strip() for c in a[1].split(",")], [c.strip() for c in b[1].split(",")]
    assert ra[0] == rb[0]
    for va, vb, tol in zip(ra[1:], rb[1:], (1e-5, 1e-5, 1e-5, 1e-3, 1e-3, 0, 1e-6)):
        fa, fb = float(va), float(vb)
        assert math.isclose(fa, fb, rel_tol=max(tol, 1e-12), abs_tol=1e-6), (a[1], b[1])

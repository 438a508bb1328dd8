"""Packed export: nibble packing round trip and  unpack_weight(pack_weight(W)) == quantizer(W)  bit for bit
for every weight format -- the externally checkable form of "integer codes and scales bit-exact"."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _q(t, f, g, zp=False):
    from llm_compressor_b200 import FakeQuantizer
    cfg = dict(type=t, format=f, group_size=g, axes=-1, zero_point=zp, is_profile=False)
    if t == "mx":
        cfg["scale_ebits"] = 8
    return FakeQuantizer.build(cfg).to(DEV)


@pytest.mark.parametrize("numel", [2, 14, 16, 4096, 100002, 3072 * 1024])
def test_pack4_unpack4_round_trip(numel):
    from llm_compressor_b200 import export
    g = torch.Generator().manual_seed(numel)
    codes = torch.randint(0, 16, (numel,), generator=g, dtype=torch.uint8).to(DEV)
    packed = export.pack4(codes)
    c = codes.cpu()
    assert torch.equal(packed.cpu(), (c[0::2] | (c[1::2] << 4)))
    assert torch.equal(export.unpack4(packed, numel, signed=False), codes)
    sx = export.unpack4(packed, numel, signed=True).view(torch.int8).cpu()
    assert torch.equal(sx, ((c.to(torch.int16) ^ 8) - 8).to(torch.int8))
    hi = (codes | 0xF0)   # garbage in the high nibbles of the input must not leak
    assert torch.equal(export.pack4(hi), packed)


CASES = [("int", "int4", 128, False), ("int", "int4", 128, True), ("int", "int8", -1, True), ("int", "int4", -1, False),
         ("fp", "fp8_e4m3", -1, False), ("fp", "fp4_e2m1", 128, True), ("mx", "fp4_e2m1", 32, False),
         ("mx", "fp8_e4m3", 32, False), ("nvfp", "fp4_e2m1", 16, False),
         # MX with fixed-point int elements: codes are in units of 2^-(mbits-2) (unpack decoded them 4x / 64x too large)
         ("mx", "int4", 32, False), ("mx", "int8", 32, False)]


@pytest.mark.parametrize("t,f,g,zp", CASES, ids=["%s-%s-g%d%s" % (t, f, g, "-zp" if zp else "") for t, f, g, zp in CASES])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_packed_blob_decodes_to_the_fake_quantised_weight(t, f, g, zp, dtype):
    from llm_compressor_b200 import export
    gen = torch.Generator().manual_seed(7)
    W = (0.02 * torch.randn(512, 3072, generator=gen)).to(dtype).to(DEV)
    W[3, :64] = 0
    q = _q(t, f, g, zp)
    blob, dq = export.pack_weight(W, q)
    assert torch.equal(dq, _q(t, f, g, zp)(W))                      # same call as the fake-quant forward
    out = export.unpack_weight(blob)
    assert out.dtype == W.dtype and torch.equal(out, dq), int((out != dq).sum())
    nbytes = blob["codes"].numel()
    assert nbytes == W.numel() // (2 if blob["bits"] == 4 else 1)
    if t == "mx":
        # the all-zero blocks of row 3 hit the reference's clamp(min=1e-5) (mx_quant.py:151): not a power of two any
        # more, the blob keeps the full scales; without such blocks the shared exponents are stored as E8M0 bytes
        assert blob["scales"] is not None
        W2 = W.clone()
        W2[3, :64] = 0.01
        blob2, dq2 = export.pack_weight(W2, _q(t, f, g, zp))
        assert blob2["scales"] is None and blob2["scales_e8m0"].dtype == torch.uint8
        assert torch.equal(export.unpack_weight(blob2), dq2)


def test_per_tensor_quantizer_is_rejected_not_divided_by_zero():
    from llm_compressor_b200 import export
    W = torch.zeros(8, 128, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(NotImplementedError):
        export.pack_weight(W, _q("int", "int8", 0, False))

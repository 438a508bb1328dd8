"""Hadamard / rotation stage of the SpinQuant-style pipeline (BASELINE config 5), with the reference's names.

ref: quantization/calibrations/spinquant/hadamard_utils.py (get_hadK :17-85, matmul_hadU :88-111, hadamard_matrix
:118-121, apply_exact_had_to_linear :135-172), rotation_utils.py (random_hadamard_matrix :40-45, rotate_* :57-113,
rotate_model :116-161), fuse_norm_utils.py (fuse_layer_norms :29-61).

The reference builds R = diag(s) * matmul_hadU(I) as a dense fp64 matrix and rotates every weight with an fp64 GEMM.
Here a randomised Hadamard rotation stays structured (`HadamardRotation`: the sign vector and n) and is applied by
the fast Walsh-Hadamard kernel `lcb_hadamard_rows` (csrc/hadamard.cu) with fp64 accumulation; only rotations that
are NOT Hadamard-structured (random orthogonal / trained R.bin matrices) take the dense fp64 GEMM (library GEMM).

The fixed H_K blocks are generated, not tabulated: the reference's had12 / had20 / had44 / had60 are Paley-I
matrices [[1, -1^T], [1, I - Q^T]] (Q the Jacobsthal matrix of GF(q), q = K - 1), had28 / had36 are Paley-II
[[S+I, S-I], [S-I, -S-I]] (S the symmetric conference matrix of order K/2) and had40 is the Sylvester double of
had20; had108 / had140 (Llama-1 13B / 30B intermediate sizes) are Paley-I again (q = 107, 139); had52 / had156 / had172
(Llama-1 13B hidden, 30B 3x hidden, Llama-2 7B+ intermediate 11008 = 172 * 64) are Williamson matrices
[[A, B, C, D], [-B, A, -D, C], [-C, D, A, -B], [-D, -C, B, A]] of four symmetric circulants whose first rows are the
classical Williamson sequences of order 13 / 39 / 43 (`_WILLIAMSON`) -- all checked entry by entry against the
reference's tables by tests/test_hadamard.py.
"""
import ctypes
import functools

import torch

from . import _lib

F64 = 2  # LCB_F64


def is_pow2(n):
    return (n & (n - 1) == 0) and (n > 0)


def _chi(a, q):
    a %= q
    if a == 0:
        return 0
    return 1 if pow(a, (q - 1) // 2, q) == 1 else -1


def _paley1(q):
    n = q + 1
    H = torch.ones(n, n)
    H[0, 1:] = -1
    for i in range(q):
        for j in range(q):
            if i != j:
                H[1 + i, 1 + j] = -_chi(j - i, q)
    return H


def _paley2(q):
    n = q + 1
    S = torch.zeros(n, n)
    S[0, 1:] = 1
    S[1:, 0] = 1
    for i in range(q):
        for j in range(q):
            S[1 + i, 1 + j] = _chi(j - i, q)
    eye = torch.eye(n)
    return torch.cat([torch.cat([S + eye, S - eye], 1), torch.cat([S - eye, -S - eye], 1)], 0)


# first halves (entries 0 .. (q - 1) / 2) of the symmetric first rows of the four circulants A, B, C, D, q = K / 4
_WILLIAMSON = {
    52: ("+-+--++", "+---+++", "++-+--+", "+----+-"),
    156: ("+++--+-+-----+--++--", "++++---+--++----+-+-", "+++--++-+---+-+--+--", "+---++-+-+-----+++-+"),
    172: ("+---++--++++-+-+++-++-", "++-++++++----+-+--++-+", "+++-+-++--+-+-++++-+--", "++---++++-+--+--++----"),
}


def _williamson(K):
    q = K // 4
    blocks = []
    for half in _WILLIAMSON[K]:
        h = [1.0 if c == "+" else -1.0 for c in half]
        row = torch.tensor(h + h[1:][::-1][: q - len(h)])          # symmetric: r[i] == r[q - i]
        assert row.numel() == q and torch.equal(row[1:], row[1:].flip(0))
        blocks.append(torch.stack([torch.roll(row, i) for i in range(q)]))
    A, B, C, D = blocks
    return torch.cat([torch.cat([A, B, C, D], 1), torch.cat([-B, A, -D, C], 1),
                      torch.cat([-C, D, A, -B], 1), torch.cat([-D, -C, B, A], 1)], 0)


@functools.lru_cache(maxsize=None)
def _had(K):
    if K in (12, 20, 44, 60, 108, 140):
        return _paley1(K - 1)
    if K in _WILLIAMSON:
        return _williamson(K)
    if K in (28, 36):
        return _paley2(K // 2 - 1)
    if K == 40:
        h = _had(20)
        return torch.cat([torch.cat([h, h], 1), torch.cat([h, -h], 1)], 0)
    raise NotImplementedError(f"no {K} x {K} Hadamard block in the reference (hadamard_utils.py:17-85)")


def get_hadK(n, transpose=False):
    """(H_K, K) with n = K * 2^m; same precedence of K as the reference (hadamard_utils.py:17-85)."""
    for K in (172, 156, 140, 108, 60, 52, 36, 28, 44, 40, 20, 12):
        if n % K == 0:
            assert is_pow2(n // K)
            h = _had(K)
            return (h.T.contiguous() if transpose else h.clone()), K
    assert is_pow2(n)
    return None, 1


def _dt(t):
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return F64
    raise NotImplementedError("Hadamard transform: dtype must be bfloat16, float32 or float64, got %s" % t.dtype)


def _bits(hadK):
    """host sign table of lcb_hadamard_rows: one 64-bit word per row for K <= 64, three per row above (include/lcb200.h)"""
    K = hadK.shape[0]
    wpr = 1 if K <= 64 else 3
    words = (ctypes.c_uint64 * (K * wpr))()
    neg = (hadK < 0).tolist()
    for r in range(K):
        for c in range(K):
            if neg[r][c]:
                words[r * wpr + (c >> 6)] |= 1 << (c & 63)
    return words


def sqrt_divisor(n):
    """`torch.tensor(n).sqrt()` of the reference: a float32 scalar (hadamard_utils.py:111)."""
    return float(torch.tensor(n).sqrt())


def hadamard_rows(x, signs=None, transpose=False, out_dtype=None, acc64=True, out=None):
    """y[r, :] = T(signs * x[r, :]) / float32(sqrt(n)) over the last dim of a contiguous CUDA tensor."""
    if not x.is_cuda:
        raise _lib.LcbError("Hadamard transform needs a CUDA tensor; this package has no CPU fallback")
    x = x.contiguous()
    n = x.shape[-1]
    rows = x.numel() // n
    hadK, K = get_hadK(n, transpose)
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    assert out.is_contiguous() and out.shape == x.shape
    if signs is not None:
        signs = signs.to(device=x.device, dtype=torch.float32).contiguous()
        assert signs.numel() == n
    stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    words = _bits(hadK) if K > 1 else None  # host table, copied into the kernel parameters by the call
    with torch.cuda.device(x.device):
        rc = _lib.lib().lcb_hadamard_rows(
            ctypes.c_void_p(x.data_ptr()), _dt(x), ctypes.c_void_p(out.data_ptr()), _dt(out), rows, n,
            ctypes.c_void_p(signs.data_ptr()) if signs is not None else None,
            ctypes.cast(words, ctypes.c_void_p) if K > 1 else None, K, sqrt_divisor(n), int(bool(acc64)), stream)
    _lib.check(rc, "lcb_hadamard_rows")
    return out


def matmul_hadU(X, transpose=False):
    """ref: hadamard_utils.py:88-111 (same dtype out as in)."""
    return hadamard_rows(X, None, transpose)


def matmul_hadUt(X):
    return matmul_hadU(X, transpose=True)


def hadamard_matrix(size, device):
    """ref: hadamard_utils.py:118-121 -- matmul_hadU(eye), float32."""
    return matmul_hadU(torch.eye(size, device=device))


class HadamardRotation:
    """R = diag(signs) @ matmul_hadU(I_n) kept structured.  `W @ R` and `R.T @ W` are fast transforms."""

    def __init__(self, signs, device):
        self.signs = signs.to(device=device, dtype=torch.float32)
        self.n = int(signs.numel())
        self.device = torch.device(device)

    def dense(self):
        """the reference's fp64 matrix (rotation_utils.py:40-45)"""
        q = torch.diag(self.signs.to(torch.float64))
        return matmul_hadU(q)

    def right(self, W, out_dtype=None):
        """W @ R for W [..., n]"""
        return hadamard_rows(W.to(self.device), self.signs, out_dtype=out_dtype)

    def left_t(self, W, out_dtype=None):
        """R.T @ W for W [n, c] (or a bias [n]): the same transform down the columns"""
        if W.dim() == 1:
            return hadamard_rows(W.to(self.device).unsqueeze(0), self.signs, out_dtype=out_dtype).squeeze(0)
        return hadamard_rows(W.to(self.device).t().contiguous(), self.signs, out_dtype=out_dtype).t().contiguous()


def random_hadamard_matrix(size, device, structured=False):
    """ref: rotation_utils.py:40-45.  Draws the signs with the same torch.randint call (same RNG stream).
    structured=True returns the HadamardRotation instead of the dense fp64 matrix."""
    q = torch.randint(low=0, high=2, size=(size,)).to(torch.float64)
    q = q * 2 - 1
    rot = HadamardRotation(q, device)
    return rot if structured else rot.dense()


def random_orthogonal_matrix(size, device):
    """ref: rotation_utils.py:21-37"""
    random_matrix = torch.randn(size, size, dtype=torch.float64).to(device)
    q, r = torch.linalg.qr(random_matrix)
    q *= torch.sign(torch.diag(r)).unsqueeze(0)
    return q


def get_orthogonal_matrix(size, mode, device, structured=True):
    if mode == "random":
        return random_orthogonal_matrix(size, device)
    if mode == "hadamard":
        return random_hadamard_matrix(size, device, structured=structured)
    raise ValueError(f"Unknown mode {mode}")


def _right(W, R, device):
    """(W @ R) in fp64 precision, cast to W's dtype, back on the CPU like the reference's rotate_* helpers."""
    dtype = W.dtype
    if isinstance(R, HadamardRotation):
        return R.right(W.to(device), out_dtype=dtype).to("cpu")
    return torch.matmul(W.to(device=device, dtype=torch.float64), R).to(device="cpu", dtype=dtype)


def _left_t(W, R, device):
    dtype = W.dtype
    if isinstance(R, HadamardRotation):
        return R.left_t(W.to(device), out_dtype=dtype).to("cpu")
    return torch.matmul(R.T, W.to(device=device, dtype=torch.float64)).to(device="cpu", dtype=dtype)


def rotate_embeddings(model, R1, device):
    for W in [model.model.embed_tokens]:
        W.weight.data = _right(W.weight.data, R1, device)


def rotate_attention_inputs(layer, R1, device):
    for W in [layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj]:
        W.weight.data = _right(W.weight.data, R1, device)


def rotate_attention_output(layer, R1, device):
    W = layer.self_attn.o_proj
    W.weight.data = _left_t(W.weight.data, R1, device)
    if W.bias is not None:
        W.bias.data = _left_t(W.bias.data, R1, device)


def rotate_mlp_input(layer, R1, device):
    for W in [layer.mlp.up_proj, layer.mlp.gate_proj]:
        W.weight.data = _right(W.weight.data, R1, device)


def rotate_mlp_output(layer, R1, device):
    W = layer.mlp.down_proj
    W.weight.data = _left_t(W.weight.data, R1, device)
    if W.bias is not None:
        W.bias.data = _left_t(W.bias.data, R1, device)


def rotate_head(model, R1, device):
    W = model.lm_head
    W.weight.data = _right(W.weight.data, R1, device)


def apply_exact_had_to_linear(module, had_dim=-1, output=False, R2=None, device="cuda"):
    """ref: hadamard_utils.py:135-172.  had_dim == -1: full-width transform of the input (or output) features;
    had_dim > 0: every block of had_dim features is multiplied by hadamard_matrix(had_dim) or by R2."""
    assert isinstance(module, torch.nn.Linear)
    if had_dim != -1:
        assert is_pow2(had_dim), "Hadamard dimension must be a power of 2!"
    W_ = module.weight.data
    dtype, dev = W_.dtype, W_.device
    W_ = W_.to(device)
    if had_dim == -1:
        if output:
            W_ = hadamard_rows(W_.t().contiguous(), None, out_dtype=dtype).t().contiguous()
        else:
            W_ = hadamard_rows(W_, None, out_dtype=dtype)
    else:
        if R2 is None:
            R2 = HadamardRotation(torch.ones(had_dim), device)
        Wt = W_.t().contiguous() if output else W_
        shape = Wt.shape
        blocks = Wt.reshape(-1, shape[-1] // had_dim, had_dim)
        if isinstance(R2, HadamardRotation):
            blocks = R2.right(blocks, out_dtype=dtype)
        else:
            blocks = (blocks.to(torch.float64) @ R2.to(device=device, dtype=torch.float64)).to(dtype)
        Wt = blocks.reshape(shape)
        W_ = Wt.t().contiguous() if output else Wt
    module.weight.data = W_.to(device=dev, dtype=dtype)


def rotate_ov_proj(layer, head_dim, R2=None, device="cuda"):
    apply_exact_had_to_linear(layer.self_attn.v_proj, had_dim=head_dim, output=True, R2=R2, device=device)
    apply_exact_had_to_linear(layer.self_attn.o_proj, had_dim=head_dim, output=False, R2=R2, device=device)


def fuse_ln_linear(layernorm, linear_layers):
    """ref: fuse_norm_utils.py:5-26 (elementwise fp64 glue, stays PyTorch)"""
    for linear in linear_layers:
        linear_dtype = linear.weight.dtype
        W_ = linear.weight.data.double()
        linear.weight.data = (W_ * layernorm.weight.double()).to(linear_dtype)
        if hasattr(layernorm, "bias"):
            if linear.bias is None:
                linear.bias = torch.nn.Parameter(torch.zeros(linear.out_features, dtype=torch.float64))
            linear.bias.data = linear.bias.data.double() + torch.matmul(W_, layernorm.bias.double())
            linear.bias.data = linear.bias.data.to(linear_dtype)


def fuse_layer_norms(model):
    """ref: fuse_norm_utils.py:29-61"""
    for W in [model.model.embed_tokens]:
        W_ = W.weight.data.double()
        W.weight.data = (W_ - W_.mean(dim=-1, keepdim=True)).to(W.weight.data.dtype)
    for layer in model.get_layers():
        fuse_ln_linear(layer.post_attention_layernorm, [layer.mlp.up_proj, layer.mlp.gate_proj])
        fuse_ln_linear(layer.input_layernorm, [layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj])
        layer.post_attention_layernorm.weight.data = torch.ones_like(layer.post_attention_layernorm.weight.data)
        layer.input_layernorm.weight.data = torch.ones_like(layer.input_layernorm.weight.data)
    fuse_ln_linear(model.model.norm, [model.lm_head])
    model.model.norm.weight.data = torch.ones_like(model.model.norm.weight.data)


@torch.inference_mode()
def rotate_model(model, rotate_mode, device, R1=None, R2s=None):
    """ref: rotation_utils.py:116-161.  R1 / R2s (dense fp64 matrices, e.g. a trained R.bin, or HadamardRotation
    objects) override the random draw; the RNG is consumed in the reference's order (R1, then one R2 per layer)."""
    config = model.config
    head_dim = config.hidden_size // config.num_attention_heads
    drawn = get_orthogonal_matrix(config.hidden_size, rotate_mode, device)
    R1 = drawn if R1 is None else R1
    rotate_embeddings(model, R1, device)
    rotate_head(model, R1, device)
    layers = model.get_layers()
    for i in range(len(layers)):
        drawn = get_orthogonal_matrix(head_dim, rotate_mode, device)
        R2 = drawn if R2s is None else R2s[i]
        rotate_attention_inputs(layers[i], R1, device)
        rotate_attention_output(layers[i], R1, device)
        rotate_mlp_input(layers[i], R1, device)
        rotate_mlp_output(layers[i], R1, device)
        rotate_ov_proj(layers[i], head_dim, R2=R2, device=device)
    return R1

"""Randomised-Hadamard rotation bandwidth: lcb_hadamard_rows (fp32 / fp64 accumulation) beside the reference's vendored
third-party FWHT (oracle/_ref/fast_hadamard_transform_cuda.so, built unmodified by oracle/make_fht.py), same box, same
tensor, bf16 in / bf16 out = 4 B per element.  The third-party kernel applies no sign vector (the reference multiplies
by the signs in a separate eager op), so its time is a lower bound for the reference's rotation.

    python scripts/had_bench.py [--once N]   # --once: one launch of the fp32-acc kernel at width N (for ncu)
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llm_compressor_b200 import hadamard as H  # noqa: E402


def tri_dao():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    try:
        import fast_hadamard_transform_cuda as F
        return F
    except Exception as e:  # not built on this box
        print("third-party FWHT not available:", e)
        return None


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", type=int, default=0)
    ap.add_argument("--acc64", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.once:
        n = args.once
        x = (0.02 * torch.randn(8 * 8192, n, device=dev)).to(torch.bfloat16)
        s = (torch.randint(0, 2, (n,), device=dev) * 2 - 1).float()
        y = torch.empty_like(x)
        for _ in range(2):
            H.hadamard_rows(x, s, acc64=bool(args.acc64), out=y)
        torch.cuda.synchronize()
        return
    F = tri_dao()
    for n in (3072, 2560, 8192, 4096, 128):
        rows = 8 * 8192 if n > 128 else 64 * 8192 * 4
        g = torch.Generator(device=dev).manual_seed(n)
        x = (0.02 * torch.randn(rows, n, generator=g, device=dev)).to(torch.bfloat16)
        s = (torch.randint(0, 2, (n,), generator=g, device=dev) * 2 - 1).float()
        y = torch.empty_like(x)
        gb = x.numel() * 4 / 1e6
        y64 = H.hadamard_rows(x, s, acc64=True)
        y32 = H.hadamard_rows(x, s, acc64=False)
        diff = (y64.float() - y32.float()).abs().max().item() / y64.float().abs().max().item()
        neq = (y64 != y32).float().mean().item()
        line = "n=%5d rows=%7d  fp32acc %7.1f GB/s  fp64acc %7.1f GB/s  (fp32 vs fp64: max rel %.2e, %.4f%% bf16 outputs differ)" % (
            n, rows, gb / timeit(lambda: H.hadamard_rows(x, s, acc64=False, out=y)),
            gb / timeit(lambda: H.hadamard_rows(x, s, acc64=True, out=y)), diff, 100 * neq)
        if F is not None:
            K = next(k for k in (1, 12, 20, 28, 40) if n % k == 0 and ((n // k) & (n // k - 1)) == 0)
            fn = {1: F.fast_hadamard_transform, 12: F.fast_hadamard_transform_12N, 20: F.fast_hadamard_transform_20N,
                  28: F.fast_hadamard_transform_28N, 40: F.fast_hadamard_transform_40N}[K]
            scale = 1.0 / (n ** 0.5)
            try:
                z = fn(x, scale)
                # same transform up to the sign vector: compare on the unsigned input
                ours = H.hadamard_rows(x, None, acc64=True)
                rel = (z.float() - ours.float()).abs().max().item() / ours.float().abs().max().item()
                line += "  | third-party FWHT %7.1f GB/s (no signs; max rel diff vs ours %.1e)" % (gb / timeit(lambda: fn(x, scale)), rel)
            except Exception as e:
                line += "  | third-party FWHT failed: %s" % str(e)[:80]
        print(line, flush=True)
        del x, y, y64, y32


if __name__ == "__main__":
    main()

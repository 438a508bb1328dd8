"""Helpers to read tests/golden/*.npz (written by oracle/gen_golden.py from the reference)."""
import ast
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits_to_f32(a):
    """uint16 bf16 bit patterns -> float32; float32 passes through."""
    if a.dtype == np.uint16:
        return (a.astype(np.uint32) << 16).view(np.float32)
    return a


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = ast.literal_eval(str(z["__meta__"])) if "__meta__" in z.files else None
    return z, meta


def same(a, b):
    """Bit-exact comparison up to the sign of zero; NaNs must coincide."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return bool(np.all((a == b) | na))


def n_diff(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return int(np.sum(~(((a == b) & ~na & ~nb) | (na & nb))))

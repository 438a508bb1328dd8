import numpy as np
import torch


def t_from_bits(a, device=None):
    """golden array (uint16 bf16 bits or float32) -> torch tensor of that dtype"""
    if a.dtype == np.uint16:
        t = torch.from_numpy(a.astype(np.int32).astype(np.int16) if False else a.view(np.int16).copy()).view(torch.bfloat16)
    else:
        t = torch.from_numpy(np.array(a, copy=True))
    return t.to(device) if device is not None else t


def to_f32_np(t):
    return t.detach().float().cpu().numpy()


def same(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.all((a == b) | na))


def n_diff(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return int(np.sum(~(((a == b) & ~na & ~nb) | (na & nb))))

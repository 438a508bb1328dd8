"""Numerical profile of a fake-quant op (`--profile`), host side.

ref: llm_compressor/quantization/quantizers/base.py:30-113 (record_stats).  Off the hot path:
the reference copies both tensors to the CPU and sorts them; this mirror keeps that behaviour
and appends the same columns to `<save_path>/stats.csv`.
"""
import csv
import os

import torch

KEYS = ("Op Name", "PC99%", "Max", "QDQ(Max)", "SQNR", "ClipError", "Elem", "BPV")


def _sqnr(t, q):
    t_ = (t - t.min()) / (t.max() - t.min())
    q_ = (q - q.min()) / (q.max() - q.min())
    return (-10 * torch.log10(torch.mean((t_ - q_) ** 2) + 1e-10)).item()


def _percentile(t, q):
    k = round(q * (t.numel() - 1))
    return torch.sort(t.flatten())[0][k].item()


def record_stats(quantizer, x, qdq_x):
    x_ = x.detach().float().cpu()
    q_ = qdq_x.detach().float().cpu()
    gs = quantizer.group_size if isinstance(quantizer.group_size, int) else 0
    bits = {"INT4": 4, "INT8": 8, "FP4_E2M1": 4, "FP8_E4M3": 8, "FP8_E5M2": 8}.get(quantizer.str_format, 16)
    bpv = bits + ((16 * (2 if quantizer.zero_point else 1)) / gs if gs and gs > 0 else 0)
    row = {
        "Op Name": quantizer.op_name,
        "PC99%": _percentile(x_, 0.99),
        "Max": x_.max().item(),
        "QDQ(Max)": q_.max().item(),
        "SQNR": _sqnr(x_, q_),
        "ClipError": (x_.abs().max() - q_.abs().max()).abs().item(),
        "Elem": x_.numel(),
        "BPV": bpv,
    }
    path = os.path.join(str(quantizer.save_path), "stats.csv")
    new = not os.path.exists(path)
    with open(path, "a", newline="") as f:
        w = csv.DictWriter(f, fieldnames=KEYS)
        if new:
            w.writeheader()
        w.writerow(row)
    return row

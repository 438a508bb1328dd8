"""Bandwidth of the Hadamard rotation kernel (lcb_hadamard_rows) at model shapes; prints JSON lines."""
import json
import sys

import torch

sys.path.insert(0, ".")
from llm_compressor_b200 import hadamard as H  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rows, n in ((8192, 3072), (3072, 8192), (10240, 2560), (2560, 10240), (128256, 3072), (8192 * 24, 128)):
    for dtype, acc64 in ((torch.bfloat16, True), (torch.bfloat16, False), (torch.float32, False)):
        x = torch.randn(rows, n, device=dev).to(dtype)
        s = (torch.randint(0, 2, (n,), device=dev) * 2 - 1).float()
        y = torch.empty_like(x)
        for _ in range(3):
            H.hadamard_rows(x, s, acc64=acc64, out=y)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            H.hadamard_rows(x, s, acc64=acc64, out=y)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        gb = 2 * x.numel() * x.element_size() / 1e9
        print(json.dumps(dict(rows=rows, n=n, dtype=str(dtype), acc64=acc64, ms=round(ms, 4), GBs=round(gb / ms * 1e3, 1))))

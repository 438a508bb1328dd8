"""One Hadamard rotation launch series for ncu: had_once.py ROWS N {bf16|f32} {64|32}"""
import sys

import torch

sys.path.insert(0, ".")
from llm_compressor_b200 import hadamard as H  # noqa: E402

rows, n = int(sys.argv[1]), int(sys.argv[2])
dtype = torch.bfloat16 if sys.argv[3] == "bf16" else torch.float32
acc64 = sys.argv[4] == "64"
x = torch.randn(rows, n, device="cuda:0").to(dtype)
s = (torch.randint(0, 2, (n,), device="cuda:0") * 2 - 1).float()
y = torch.empty_like(x)
for _ in range(4):
    H.hadamard_rows(x, s, acc64=acc64, out=y)
torch.cuda.synchronize()
print("ok")

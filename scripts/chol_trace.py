"""Task trace of the tile-task Cholesky kernel (LCB_CHOL_TRACE=1): where the panel chain spends its time."""
import os
import sys
os.environ["LCB_CHOL_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from llm_compressor_b200 import _lib, ops

dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(2 * K, K, generator=g, device=dev).to(torch.bfloat16).float()
H = (X.T @ X / K).contiguous()
U = torch.empty_like(H)
L = _lib.lib()
ws = torch.empty(L.lcb_chol_ws_bytes(K), dtype=torch.uint8, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(3):
    rc = L.lcb_chol_inv_upper(H.data_ptr(), U.data_ptr(), K, None, 0.01, ws.data_ptr(), ws.numel(), status.data_ptr(),
                              torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
nblk = (K + 127) // 128
off = L.lcb_chol_trace_offset(K)
tr = ws[off: off + nblk * nblk * 64].cpu().numpy().view(np.uint64).reshape(-1, 8).astype(np.int64)
t0 = tr[:, 1][tr[:, 1] > 0].min()
rows = []
for r in tr:
    typ, i, c = int(r[0] >> 32), int((r[0] >> 16) & 0xffff), int(r[0] & 0xffff)
    rows.append((typ, i, c, *[(int(v) - t0) / 1e3 if v else -1.0 for v in r[1:6]]))
print("total span us: %.1f" % max(x[7] for x in rows))
print("diag tasks: c fetch acc_done factor_done flag_set(end)   [us]")
for x in rows:
    if x[0] == 0 and x[1] == x[2]:
        print("  c=%2d fetch %8.1f acc %8.1f factor %8.1f end %8.1f  (factor %.1f, invert+store %.1f)" %
              (x[2], x[3], x[4], x[6], x[7], x[6] - x[4], x[7] - x[6]))
print("sub-diagonal tasks L(c+1, c): fetch acc_done diagflag end")
for x in rows:
    if x[0] == 0 and x[1] == x[2] + 1:
        print("  c=%2d fetch %8.1f acc %8.1f diagflag %8.1f end %8.1f (trsm %.1f)" % (x[2], x[3], x[4], x[5], x[7], x[7] - x[5]))
ys = [x for x in rows if x[0] == 1]
if ys:
    print("Y tasks: last end %.1f; of column nblk-1: " % max(x[7] for x in ys))
    for x in ys:
        if x[2] == nblk - 1 and x[1] in (0, nblk // 2, nblk - 2):
            print("  Y(%d,%d) fetch %.1f acc %.1f diagflag %.1f end %.1f" % (x[1], x[2], x[3], x[4], x[5], x[7]))
# per-step clock64 probes inside factor_invert_la of diagonal task c = 1 (cycles, relative to step 0 start)
dbg = ws[off + nblk * nblk * 64: off + nblk * nblk * 64 + 16 * 12 * 8].cpu().numpy().view(np.int64).reshape(16, 12)
b0 = dbg[0, 0]
print("joint form (LCB_CHOL_FI=0): chain[P1start P1end P3Async LAend] worker0[P1sync P2Wend LAsync first P3end barW]")
print("split form: chain[start factored P3Async LAend] Tworker0[P1sync solved LAsync first trailing_end] Wworker0[start end]  (cycles since step-0 start)")
for p in range(16):
    print("  p=%2d " % p + " ".join("%7d" % (int(v - b0) if v else -1) for v in dbg[p, :11]))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of time per kernel.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
    name = row["Kernel Name"]
    m = re.search(r"((?:lcb|at|cub|nccl)[\w:]*::)?(\w+)\s*(?:<|\()", name)
    short = (m.group(1) or "") + m.group(2) if m else name[:60]
    tot[short] += v
    cnt[short] += 1
T = sum(tot.values())
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {T / 1e6:.3f} ms of kernel time (ncu, serialised, cold cache)")
print(f"{'share':>7} {'total ms':>10} {'count':>6} {'avg us':>9}  kernel")
for k, v in tot.most_common(40):
    print(f"{v / T * 100:6.2f}% {v / 1e6:10.3f} {cnt[k]:6d} {v / cnt[k] / 1e3:9.1f}  {k}")

"""Fake-quant bandwidth table of bench.py (all FQ_CASES) without the rest of the benchmark.  Development aid."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import llm_compressor_b200 as lc
dev = torch.device("cuda:0")
print(json.dumps(bench.fake_quant_bandwidth(lc, torch, dev, bench.peaks()[2])))

"""bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`, the unmodified reference timed
on the host cores) prints ONE JSON line with the keys the driver reads, and the product arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("GPTQ W4g128 sec/model") and d["unit"] == "s/model"
    assert d["higher_is_better"] is False and d["value"] > 0 and d["steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "s/model", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    # "reference" = the unmodified reference from oracle/_ref (or /root/reference); "port" only when neither is present
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "llm_compressor_ref.zip")) or os.path.isdir("/root/reference/llm_compressor")
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]

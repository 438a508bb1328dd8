"""Headline benchmark: GPTQ W4 g128 sec/model on Llama-3.2-3B shapes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the compression hot path over one whole model: for each of the 28
decoder layers, the calibration Hessians of the 4 distinct Linear inputs accumulated sample by
sample over 128 x 2048 synthetic bf16 tokens (tcgen05 kernel), then for each of the 7 Linears
the act-order GPTQ solve (blocked Cholesky-inverse, 128-column quantize-and-propagate blocks,
lazy-batch GEMM updates) with int4-g[128]-rw fake-quant, random-init bf16 weights.
Calibration FORWARD passes are outside the path (north_star: they stay PyTorch); activations are
synthetic, created once on the device, as the reference's hooks receive them (device tensors).

Numbers printed (one JSON line on rank 0):
  value  sec/model, device timed (CUDA events), weights resident in HBM
  e2e    sec/model through the module API with the weights in pinned HOST memory: per layer
         H2D copy of the bf16 weights, solve, D2H copy of the compressed weights
  roofline   dominant kernel = hessian_umma_kernel (tensor bound): algorithmic 2*T*K^2 flop per
             launch / measured average launch duration, against the measured sustained bf16 peak
  cpu_baseline  the CPU oracle (port of the reference's PyTorch ops, numpy/LAPACK + C) timed on
             a bounded sample of the same workload and extrapolated to sec/model
Multi-GPU (torchrun, one rank per GPU): calibration samples are sharded over ranks, raw X^T X
sums are all-reduced over NCCL, Linear output rows are sharded for the solve and all-gathered.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Llama-3.2-3B (public HF config): hidden 3072, ffn 8192, 28 layers, 24 heads / 8 kv heads x 128
D_MODEL, D_FFN, N_LAYERS, D_KV = 3072, 8192, 28, 1024
N_SAMPLES, SEQ_LEN = 128, 2048
# (input width K, [(name, out rows N), ...]) for the 4 sequential groups (ref: models/llama.py:236-242)
GROUPS = [
    (D_MODEL, [("q_proj", D_MODEL), ("k_proj", D_KV), ("v_proj", D_KV)]),
    (D_MODEL, [("o_proj", D_MODEL)]),
    (D_MODEL, [("gate_proj", D_FFN), ("up_proj", D_FFN)]),
    (D_FFN, [("down_proj", D_MODEL)]),
]
WCFG = dict(type="int", format="int4", group_size=128, axes=-1, zero_point=False, is_profile=False)


def peaks():
    """(sustained bf16 TF/s, burst bf16 TF/s, HBM GB/s, SM MHz the sustained figure was taken at, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return (d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0),
                (d.get("clocks_under_load") or {}).get("sm_mhz_median"), "measured")
    return 1400.0, 1590.0, 6650.0, None, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- ours
def hessian_executed_fraction(K, BM=128, BN=256):
    """share of the K x K tile grid the symmetric-half kernel actually computes (hessian.cu launch_xtx)"""
    tm, tn = (K + BM - 1) // BM, (K + BN - 1) // BN
    return sum(tn - (mt * BM) // BN for mt in range(tm)) / float(tm * tn)


FQ_CASES = [
    ("int4_g128_asym", dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True), 4),
    ("int8_token", dict(type="int", format="int8", group_size=-1, axes=-1, zero_point=False), 4),
    ("mxfp4_g32", dict(type="mx", format="fp4_e2m1", group_size=32, axes=-1, zero_point=False), 4),
    ("mxfp8_g32", dict(type="mx", format="fp8_e4m3", group_size=32, axes=-1, zero_point=False), 4),
    ("nvfp4_g16", dict(type="nvfp", format="fp4_e2m1", group_size=16, axes=-1, zero_point=False), 6),
    # the other granularities north_star lists (per-tensor, per-channel, per-token FP8) and an fp32 tensor
    ("fp8_e4m3_token", dict(type="fp", format="fp8_e4m3", group_size=-1, axes=-1, zero_point=False), 4),
    ("int8_tensor", dict(type="int", format="int8", group_size=0, axes=-1, zero_point=False), 6),
    ("int8_channel_axis-2", dict(type="int", format="int8", group_size=-2, axes=-2, zero_point=False), 6),
    ("int4_g128_asym_fp32", dict(type="int", format="int4", group_size=128, axes=-1, zero_point=True), 8),
]


def fake_quant_bandwidth(lc, torch, dev, hbm_gbs):
    """Second half of the metric: fused fake-quant HBM GB/s on a [65536, 3072] bf16 tensor (403 MB, larger
    than L2).  Algorithmic bytes per element: 4 (read + write bf16); NVFP4 6 (the whole-tensor amax needs
    a first pass, ref: nvfp_quant.py:87); per-tensor and per-channel (reduction down the rows) scales likewise need
    the statistics before the first element can be written: 6 B / element; the fp32 case moves 8 B / element."""
    g = torch.Generator(device=dev).manual_seed(2)
    x = (0.02 * torch.randn(8 * 8192, 3072, generator=g, device=dev)).to(torch.bfloat16)
    out = {}
    x32 = None
    for name, cfg, bpe in FQ_CASES:
        if name.endswith("_fp32"):
            if x32 is None:
                x32 = x[: x.shape[0] // 2].float()
            xin = x32
        else:
            xin = x
        q = lc.FakeQuantizer.build(dict(cfg, is_profile=False)).to(dev)
        q.check_nan = False   # the NaN-scale assert (int_quant.py:165) is a host sync; checked in the tests
        for _ in range(3):
            q(xin)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            q(xin)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        gbs = xin.numel() * bpe / ms / 1e6
        out[name] = {"GBs": gbs, "frac_of_hbm_peak": gbs / hbm_gbs, "bytes_per_elem": bpe, "ms": ms,
                     "tensor": "%s %s" % (list(xin.shape), str(xin.dtype).replace("torch.", ""))}
    del x, x32
    return out


def fused_qlinear_times(lc, torch, dev):
    """F.linear(input_quantizer(x), W) of one calibration forward call ([1, 2048, K] bf16 activation, int8
    per-token and int8-g128 activation quantisers) as quantizer kernel + cuBLAS (the reference's two-op form on our
    quantizer) against the fused lcb_qlinear_fwd (find-only pass + tcgen05 GEMM with the QDQ in its operand prologue)."""
    import torch.nn.functional as F
    from llm_compressor_b200 import ops
    out = {}
    g = torch.Generator(device=dev).manual_seed(4)
    for name, K, N in (("qkv_3072x5120", D_MODEL, D_MODEL + 2 * D_KV), ("gate_up_3072x16384", D_MODEL, 2 * D_FFN), ("down_8192x3072", D_FFN, D_MODEL)):
        x = torch.randn(1, SEQ_LEN, K, generator=g, device=dev).to(torch.bfloat16)
        W = (0.02 * torch.randn(N, K, generator=g, device=dev)).to(torch.bfloat16)
        for gs in (-1, 128):
            q = lc.FakeQuantizer.build(dict(type="int", format="int8", group_size=gs, axes=-1, zero_point=False, is_profile=False)).to(dev)
            q.check_nan = False
            res = {}
            for key, fn in (("two_kernel_ms", lambda: F.linear(q(x), W)), ("fused_ms", lambda: ops.qlinear_forward(x, W, None, q)),
                            ("gemm_only_cublas_ms", lambda: F.linear(x, W))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    fn()
                b.record()
                torch.cuda.synchronize()
                res[key] = a.elapsed_time(b) / 20
            res["fused_tflops"] = 2.0 * SEQ_LEN * K * N / res["fused_ms"] / 1e9
            out["%s_int8_%s" % (name, "token" if gs == -1 else "g128")] = res
    return out


def rotation_bandwidth(torch, dev, hbm_gbs):
    """randomised Hadamard rotation W @ R1 (lcb_hadamard_rows) on a [65536, n] bf16 tensor (larger than
    L2), n = 3072 (Llama-3.2-3B hidden, K = 12), 2560 (Gemma-3 / Qwen3 hidden, K = 40), 8192 (Llama-3.2-3B FFN).
    4 B / element (read + write).  Beside it the reference's vendored third-party FWHT (oracle/_ref, built unmodified for
    sm_100a by oracle/make_fht.py) on the same tensor: no sign vector there, so its time is a lower bound for the
    reference's rotation."""
    from llm_compressor_b200 import hadamard as H
    F = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        import fast_hadamard_transform_cuda as F
    except Exception:
        F = None

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 10

    out = {}
    for n in (3072, 2560, 8192):
        g = torch.Generator(device=dev).manual_seed(n)
        x = (0.02 * torch.randn(8 * 8192, n, generator=g, device=dev)).to(torch.bfloat16)
        s = (torch.randint(0, 2, (n,), generator=g, device=dev) * 2 - 1).float()
        y = torch.empty_like(x)
        for acc64 in (True, False):
            ms = timed(lambda: H.hadamard_rows(x, s, acc64=acc64, out=y))
            gbs = x.numel() * 4 / ms / 1e6
            out["n%d_%s" % (n, "fp64acc" if acc64 else "fp32acc")] = {"GBs": gbs, "frac_of_hbm_peak": gbs / hbm_gbs, "ms": ms}
        if F is not None:
            K = next(k for k in (1, 12, 20, 28, 40) if n % k == 0 and ((n // k) & (n // k - 1)) == 0)
            fn = {1: F.fast_hadamard_transform, 12: F.fast_hadamard_transform_12N, 20: F.fast_hadamard_transform_20N,
                  28: F.fast_hadamard_transform_28N, 40: F.fast_hadamard_transform_40N}[K]
            try:
                ms = timed(lambda: fn(x, 1.0 / n ** 0.5))
                out["n%d_third_party_fwht" % n] = {"GBs": x.numel() * 4 / ms / 1e6, "ms": ms}
            except Exception as e:
                out["n%d_third_party_fwht" % n] = {"unavailable": str(e)[:80]}
        del x, y
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import llm_compressor_b200 as lc
    from llm_compressor_b200 import _lib, ops, parallel, solvers

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    # ---- synthetic inputs (identical on every rank)
    g = torch.Generator(device=dev).manual_seed(0)
    T_total = N_SAMPLES * SEQ_LEN
    xbuf = torch.randn(T_total * D_FFN, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    # four distinct activation tensors carved from one 4.3 GB buffer (larger than L2 by far)
    offs = [0, 7 * 3072 * 2048, 19 * 3072 * 2048, 0]
    acts = []
    for gi, (K, _) in enumerate(GROUPS):
        acts.append(xbuf[offs[gi]: offs[gi] + T_total * K].view(N_SAMPLES, SEQ_LEN, K))
    my_samples = parallel.sample_shard(N_SAMPLES)

    # weights: one stacked [sum N, K] bf16 matrix per sequential group (q/k/v and gate/up share their input)
    def make_weights(host):
        ws = []
        gw = torch.Generator(device=dev).manual_seed(1)
        for layer in range(args.layers):
            lw = []
            for K, lins in GROUPS:
                ntot = sum(N for _, N in lins)
                w = (0.02 * torch.randn(ntot, K, generator=gw, device=dev)).to(torch.bfloat16)
                if host:
                    hw = torch.empty((ntot, K), dtype=torch.bfloat16, pin_memory=True)
                    hw.copy_(w)
                    w = hw
                lw.append(w)
            ws.append(lw)
        return ws

    weights_dev = make_weights(host=False)
    weights_host = make_weights(host=True)
    out_host = [[torch.empty(w.shape, dtype=w.dtype, pin_memory=True) for w in lw] for lw in weights_host]
    copy_stream = torch.cuda.Stream(device=dev)   # H2D of the next group's weights
    d2h_stream = torch.cuda.Stream(device=dev)    # D2H of finished groups (second copy engine), off the compute stream

    class Lin(torch.nn.Module):
        def __init__(self, w):
            super().__init__()
            self.weight = torch.nn.Parameter(w, requires_grad=False)
            self.weight_quantizer = lc.FakeQuantizer.build(WCFG).to(dev)

    stage_evs = []  # (hessian end, factorize end, update end) events of the last step
    stage_bufs, stage_free, results = {}, {}, []   # e2e: static H2D staging buffers, their "free again" events, live results
    want_sum, chk = [False], {"sum_abs": 0.0, "sum_sq": 0.0, "zeros": 0, "elements": 0}

    def one_model(from_host):
        """One step. Returns the (start, end, flops) CUDA events of every Hessian accumulation."""
        evs = []
        stage_evs.clear()
        cur = torch.cuda.current_stream(dev)
        staged = {}

        def prefetch(layer, gi):  # H2D of one group's weights on the copy stream (pinned -> HBM)
            if not from_host or layer >= args.layers:
                return
            wh = weights_host[layer][gi]
            # rows are sharded over the ranks for the solve: each rank stages only its own rows -- into one of two
            # STATIC device buffers per group (allocated once): a tensor allocated on the copy stream per prefetch and handed to
            # the compute stream makes the caching allocator track cross-stream events, and blocks whose events are still
            # pending are not reusable, so the timed steps kept falling back to cudaMalloc (a device synchronisation each)
            src = wh[parallel.row_shard(wh.shape[0])]
            key = (gi, layer & 1)
            if key not in stage_bufs:
                stage_bufs[key] = torch.empty(src.shape, dtype=src.dtype, device=dev)
                # the block comes from the compute stream's pool and may be a just-freed temporary that kernels already
                # enqueued there still write: the copy stream must not touch it before they are done
                first = torch.cuda.Event()
                first.record(cur)
                stage_free[key] = first
            w = stage_bufs[key]
            with torch.cuda.stream(copy_stream):
                if key in stage_free:                      # the solve that last read this buffer (two layers ago) is done
                    copy_stream.wait_event(stage_free[key])
                w.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            staged[(layer, gi)] = (w, ev)

        prefetch(0, 0)
        for layer in range(args.layers):
            for gi, (K, lins) in enumerate(GROUPS):
                nxt = (layer, gi + 1) if gi + 1 < len(GROUPS) else (layer + 1, 0)
                prefetch(*nxt)
                H = torch.zeros(K, K, device=dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                X = acts[gi]
                acc = ops.HessianAccumulator(H, args.hessian_defer)
                for j in my_samples:  # one hook call per calibration sample ([1, 2048, K]), like the reference
                    acc.add(X[j])
                # raw partial sums over NVLink (no-op on 1 GPU); the split of the calibration set is static, so the total
                # count is known without an all-reduce + host read-back per Hessian
                # (N > 1: as the packed upper blocks, half the bytes; the same call finalizes: scale 2/n + mirror)
                n = parallel.reduce_finalize_hessian_(H, acc.flush(), n_total=N_SAMPLES)
                e1.record()
                evs.append((e0, e1, 2.0 * SEQ_LEN * K * K * len(my_samples), K))
                e2 = torch.cuda.Event(enable_timing=True)
                fac = solvers.factorize(H, WCFG["group_size"], actorder=True, percdamp=0.01)
                e2.record()
                del H
                ntot = weights_dev[layer][gi].shape[0]
                rows = parallel.row_shard(ntot)  # rows are independent given U: shard them over ranks
                if from_host:
                    w, ev = staged.pop((layer, gi))   # this rank's rows only (a static staging buffer)
                    cur.wait_event(ev)
                else:
                    w = weights_dev[layer][gi][rows]
                # the group's Linears as one stacked problem (solvers.update_weights_shared does this stacking
                # for separate modules; here the weights are stored stacked)
                lin = Lin(w)
                solvers.update_weight(lin, dev, block_size=128, percdamp=0.01, actorder=True, factor=fac)
                e3 = torch.cuda.Event(enable_timing=True)
                e3.record()
                out = parallel.gather_rows(lin.weight.data, ntot)
                stage_evs.append((e1, e2, e3))
                if want_sum[0]:   # untimed pass: checksum of the gathered result, comparable across N
                    o64 = out.double()
                    chk["sum_abs"] += float(o64.abs().sum())
                    chk["sum_sq"] += float((o64 * o64).sum())
                    chk["zeros"] += int((out == 0).sum())
                    chk["elements"] += out.numel()
                if from_host:  # device -> pinned host buffer on the D2H stream, ordered after the solve by an event
                    done = torch.cuda.Event()
                    done.record(cur)
                    mine = lin.weight.data  # every rank writes its own rows of the result to the host buffer
                    stage_free[(gi, layer & 1)] = done   # ... and the staging buffer may take the weights of layer + 2
                    with torch.cuda.stream(d2h_stream):
                        d2h_stream.wait_event(done)
                        out_host[layer][gi][rows].copy_(mine, non_blocking=True)
                    results.append(mine)   # kept alive until the D2H stream has been joined (no record_stream bookkeeping)
        if from_host:
            cur.wait_stream(d2h_stream)  # the step ends when the last result has reached the host buffer
        results.clear()
        return evs

    def timed(from_host, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = L.lcb_launch_count()
        s.record()
        evs = []
        for _ in range(steps):
            evs += one_model(from_host)
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, evs, L.lcb_launch_count() - launches0

    for _ in range(args.warmup):
        one_model(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, evs, launches = timed(False, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    hess_flops = sum(f for _, _, f, _ in evs)
    hess_exec = sum(f * hessian_executed_fraction(K) for _, _, f, K in evs)
    hess_ms_total = sum(a.elapsed_time(b) for a, b, _, _ in evs)
    fact_ms = sum(a.elapsed_time(b) for a, b, _ in stage_evs)      # last step (rank 0's view)
    upd_ms = sum(b.elapsed_time(c) for _, b, c in stage_evs)
    hess_last_ms = sum(a.elapsed_time(b) for a, b, _, _ in evs[-len(stage_evs):]) if stage_evs else 0.0
    n_hess_launch = len(evs) * ((len(my_samples) + args.hessian_defer - 1) // args.hessian_defer)

    want_sum[0] = True
    one_model(False)   # untimed: checksum of the whole gathered result (same inputs at every N)
    want_sum[0] = False
    checksum = dict(chk, note="over the gathered compressed weights of one model (untimed extra pass); N > 1 sums the Hessians "
                              "in a different order, so the sums agree to ~1e-4 relative with N = 1, not bit for bit")
    e2e_steps = max(1, min(args.steps, 2))
    one_model(True)  # warm the pinned-memory / copy-stream path
    e2e_ms, _, _ = timed(True, e2e_steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, burst_tf, hbm_gbs, peak_mhz, src = peaks()
    fq = rot = fq_clocks = None
    if not args.no_fake_quant:
        # The bandwidth sections time single kernels ALONE, like the copy peak they are quoted against
        # (MEASURED_PEAKS.json).  Right after the timed GPTQ region the SM clock is still held down by the power cap
        # of the tensor-core phase (1.6-1.7 GHz), which the ALU-heavy MX / NVFP / FP8 formats feel: give the governor
        # a moment to release it, and record the clock these sections actually ran at.
        torch.cuda.synchronize()
        time.sleep(3.0)
        s2 = ClockSampler(local)
        s2.start()
        fq = fake_quant_bandwidth(lc, torch, dev, hbm_gbs)
        rot = rotation_bandwidth(torch, dev, hbm_gbs)
        fq_clocks = s2.stop()
    achieved = hess_flops / (hess_ms_total * 1e-3) / 1e12
    executed = hess_exec / (hess_ms_total * 1e-3) / 1e12
    wbytes = sum(N * K * 2 for K, lins in GROUPS for _, N in lins) * args.layers
    traffic = None
    tp = os.path.join(ROOT, "profiles", "hessian_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch_avg")
    fused = fused_qlinear_times(lc, torch, dev) if not args.no_fake_quant else None
    chol_cmp = cholesky_vs_cusolver(torch, dev) if not args.no_fake_quant else None
    eager = reference_eager_b200(torch, dev) if not args.no_fake_quant else None
    cpu = None
    if not args.no_cpu_baseline:
        cpu = reference_cpu_sample() or cpu_baseline_port()
    out = {
        "metric": "GPTQ W4g128 sec/model (Llama-3.2-3B)", "value": ms / 1e3 / args.steps, "unit": "s/model",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16 x bf16 -> f32 (Hessian, tcgen05), 3xTF32 -> f32 (solver contractions, tcgen05), f32 (quantizer)",
        "data": "synthetic",
        "config": {"workload": "Llama-3.2-3B shapes (28 layers x 7 Linears), GPTQ int4-g[128]-rw act-order, "
                               "synthetic 128x2048-token bf16 activations per Linear input, random-init bf16 weights",
                   "calibration_forwards": "outside the path (north_star)", "layers": args.layers, "hessians_per_layer": 4,
                   "hessian_form": "one hook call per sample; raw sums of the symmetric half, %d hook inputs per kernel launch "
                                   "(by reference, no copy), one finalize (2/n, mirror) per Hessian" % args.hessian_defer,
                   "solve_form": "q/k/v and gate/up stacked into one [sum N, K] solve per shared Hessian",
                   "l2_note": "activation buffer 4.3 GB and Hessians 38-268 MB: inputs larger than L2",
                   "parallelism": "samples sharded + all-reduce(H), rows sharded + all-gather" if world > 1 else "single GPU"},
        "e2e": {"value": e2e_ms / 1e3 / e2e_steps, "unit": "s/model", "h2d_bytes_per_step": wbytes,
                "d2h_bytes_per_step": wbytes,
                "note": "weights in pinned host memory (each rank stages and returns its own row shard), H2D prefetched one group ahead on a copy stream, results "
                        "copied back to pinned host memory on a second copy stream (all inside the timed region); activations are produced on the device in the "
                        "reference flow too"},
        "gpu_launches": int(launches),
        "stages_s_per_model": {"hessian (incl. all-reduce + finalize)": hess_last_ms / 1e3,
                               "factorize (dead fix, act-order sort, Cholesky-inverse)": fact_ms / 1e3,
                               "update (find_params, permutes, block loop, lazy-batch GEMMs)": upd_ms / 1e3,
                               "note": "device time between CUDA events of the last timed step; the rest of `value` is "
                                       "the row all-gather (N > 1) and allocator / launch gaps"},
        "clocks": clocks,
        "roofline": {"kernel": "hessian_umma_kernel / hessian_umma_pair_kernel", "bound": "tensor",
                     "achieved": executed, "peak": peak_tf, "unit": "TFLOP/s", "frac": executed / peak_tf, "traffic": traffic,
                     "peak_source": src + " sustained bf16 (MEASURED_PEAKS.json, taken at a median %s MHz)" % peak_mhz,
                     "note": "achieved / frac count the MMAs the kernel ISSUES: it computes only the tiles touching the upper "
                             "triangle of the symmetric H (52-54 % of the 2*T*K^2 flop of SURVEY 8d); frac_algorithmic "
                             "divides the full algorithmic count by the same time and can therefore exceed 1",
                     "achieved_algorithmic": achieved, "frac_algorithmic": achieved / peak_tf,
                     "frac_of_burst_peak": executed / burst_tf, "burst_peak": burst_tf,
                     "sm_mhz_during_run": (clocks or {}).get("sm_mhz"),
                     "launches": n_hess_launch, "avg_launch_ms": hess_ms_total / max(n_hess_launch, 1),
                     "stage_share_of_step": hess_ms_total / ms},
        "fake_quant": fq,
        "hadamard_rotation": rot,
        "bandwidth_sections_clocks": fq_clocks,
        "fused_act_qdq_linear": fused,
        "cholesky_inverse": chol_cmp,
        "reference_eager_b200": eager,
        "result_checksum": checksum,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- reference legs
def _ref_timing():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_timing
    return ref_timing if ref_timing.available() else None


def _physical_cores():
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
        if n:
            return int(n)
    except Exception:
        pass
    return os.cpu_count() or 1


_REF_CACHE = {}


def reference_cpu_sample(budget_s=8.0):
    """The UNMODIFIED reference (oracle/_ref or /root/reference, oracle/ref_timing.py) on the host cores: its gptq()
    driver on a one-Linear block; timed inside it are the forward hook `cache_hessian_weight` (gptq/core.py:103-119) on one
    2048-token sample and `update_weight` (core.py:163-281) on an [N, K] slice at two N (fixed part: Cholesky chain, find_params;
    per-row part: block loop).  K = 3072 is re-measured every step, K = 8192 (a 25 s+ factorisation chain on a CPU) once per
    process.  Extrapolated to sec/model over the reference's REAL flow: 7 hooked Linears and 7 update_weight calls per layer
    (it does not share the q/k/v or gate/up Hessians), 28 layers, 128 samples."""
    import torch
    rt = _ref_timing()
    if rt is None:
        return None
    cores = _physical_cores()
    torch.set_num_threads(cores)       # torchrun exports OMP_NUM_THREADS=1: override it, the reference arm may use every core

    def fit(K, n1, n2):
        _, h1, u1 = rt.run_gptq(n1, K, 2, SEQ_LEN, "cpu")
        _, h2, u2 = rt.run_gptq(n2, K, 2, SEQ_LEN, "cpu")
        hook = min(h1[1:] + h2[1:])                       # first call allocates
        slope = max((u2 - u1) / (n2 - n1), 0.0)
        return hook, max(u1 - slope * n1, 0.0), slope, (u1, u2)

    t0 = time.perf_counter()
    k3 = fit(D_MODEL, 256, 1024)
    if D_FFN not in _REF_CACHE:
        _REF_CACHE[D_FFN] = fit(D_FFN, 128, 512)
    k8 = _REF_CACHE[D_FFN]
    hess = N_LAYERS * N_SAMPLES * (6 * k3[0] + k8[0])
    upd = 0.0
    for Kg, lins in GROUPS:
        hk, fixed, slope, _ = k3 if Kg == D_MODEL else k8
        for _, N in lins:
            upd += fixed + slope * N
    upd *= N_LAYERS
    return {"value": hess + upd, "unit": "s/model", "cores": cores, "kind": "reference",
            "torch_threads": torch.get_num_threads(),
            "sample": "unmodified reference gptq() on a one-Linear block (oracle/ref_timing.py): hook on one 2048-token sample "
                      "K=3072 %.3f s / K=8192 %.3f s; update_weight [256|1024, 3072] %.2f / %.2f s (every step), "
                      "[128|512, 8192] %.2f / %.2f s (once per process); extrapolated: 28 layers x (128 samples x 7 hooks + 7 "
                      "update_weight calls, fixed + per-row parts fitted from the two N)" %
                      (k3[0], k8[0], k3[3][0], k3[3][1], k8[3][0], k8[3][1]),
            "stages": {"hessian": hess, "update_weight": upd}, "sample_wall_s": time.perf_counter() - t0}


def reference_eager_b200(torch, dev):
    """Same-box bar: the unmodified reference functions in PyTorch eager on THIS GPU -- hook (fp32 SGEMM),
    update_weight (cuSOLVER potrf -> potri -> potrf + eager block loop) at the four distinct Linear shapes, eager fake-quant
    of the five headline formats; extrapolated to the reference's flow (7 hooks + 7 solves per layer)."""
    rt = _ref_timing()
    if rt is None:
        return {"unavailable": "oracle/_ref not present (run __graft_entry__.build() where /root/reference is mounted)"}
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    out = {"hook_ms": {}, "update_weight_ms": {}, "fake_quant_GBs": {}}
    shapes = [(D_MODEL, D_MODEL), (D_KV, D_MODEL), (D_FFN, D_MODEL), (D_MODEL, D_FFN)]
    for N, K in shapes:
        _, hs, us = rt.run_gptq(N, K, 4, SEQ_LEN, str(dev))
        _, hs2, us2 = rt.run_gptq(N, K, 4, SEQ_LEN, str(dev))      # second run: cuSOLVER / allocator warm
        out["hook_ms"]["K%d" % K] = min(hs2[1:]) * 1e3
        out["update_weight_ms"]["%dx%d" % (N, K)] = min(us, us2) * 1e3
    h3, h8 = out["hook_ms"]["K%d" % D_MODEL], out["hook_ms"]["K%d" % D_FFN]
    u = out["update_weight_ms"]
    per_layer_ms = N_SAMPLES * (6 * h3 + h8) + 2 * u["%dx%d" % (D_MODEL, D_MODEL)] + 2 * u["%dx%d" % (D_KV, D_MODEL)] \
        + 2 * u["%dx%d" % (D_FFN, D_MODEL)] + u["%dx%d" % (D_MODEL, D_FFN)]
    out["sec_per_model_extrapolated"] = N_LAYERS * per_layer_ms / 1e3
    g = torch.Generator(device=dev).manual_seed(3)
    x = (0.02 * torch.randn(8192, 3072, generator=g, device=dev)).to(torch.bfloat16)
    for name, cfg, bpe in FQ_CASES[:5]:
        q = ref_shim.build_quantizer(dict(cfg, is_profile=False)).to(dev)
        for _ in range(2):
            q(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            q(x)
        b.record()
        torch.cuda.synchronize()
        out["fake_quant_GBs"][name] = x.numel() * bpe / (a.elapsed_time(b) / 5) / 1e6
    out["note"] = ("unmodified reference code, device=cuda, torch %s; hook = one 2048-token call; update_weight includes its own "
                   "cuSOLVER Cholesky chain; fake-quant on [8192, 3072] bf16 at the same algorithmic bytes/element as ours" % torch.__version__)
    return out


def cholesky_vs_cusolver(torch, dev):
    """lcb_chol_inv_upper next to the reference's chain (torch.linalg.cholesky -> cholesky_inverse -> cholesky(upper),
    gptq/core.py:213-224) in PyTorch eager (cuSOLVER) on this GPU."""
    from llm_compressor_b200 import ops
    out = {}
    for K in (D_MODEL, D_FFN):
        g = torch.Generator(device=dev).manual_seed(K)
        X = torch.randn(2 * K, K, generator=g, device=dev).to(torch.bfloat16).float()
        H = ((1.0 / K) * X.T @ X).contiguous()
        del X

        def ours():
            U, pend = ops.chol_inv_upper(H, percdamp=0.01, defer=True)
            return pend

        def ref():
            Hf = H.clone()
            Hf.diagonal().add_(0.01 * torch.mean(torch.diag(Hf)))
            return torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hf)), upper=True)

        res = {}
        for name, fn, n in (("ours_ms", ours, 5), ("reference_eager_cusolver_ms", ref, 2)):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            res[name] = a.elapsed_time(b) / n
        out["K%d" % K] = res
        del H
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    vals, cpu = [], None
    for _ in range(args.warmup + args.steps):
        cpu = reference_cpu_sample()
        if cpu is None:
            break
        vals.append(cpu["value"])
    if cpu is None:
        # the reference copy is missing: time the oracle port instead (kind: "port")
        for _ in range(max(1, args.warmup > 0) + args.steps):
            cpu = cpu_baseline_port()
            vals.append(cpu["value"])
    v = sum(vals[-args.steps:]) / args.steps
    out = {
        "impl": "reference", "metric": "GPTQ W4g128 sec/model (Llama-3.2-3B)", "value": v, "unit": "s/model",
        "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Llama-3.2-3B shapes (28 layers x 7 Linears), GPTQ int4-g[128]-rw act-order, "
                               "synthetic 128x2048-token bf16 activations per Linear input, random-init bf16 weights",
                   "note": "each step is a bounded sample of the reference's own code on the host cores, extrapolated to "
                           "sec/model (the full job is hours on a CPU); see cpu_baseline.sample"},
        "cpu_baseline": dict(cpu, value=v),
        "e2e": {"value": v, "unit": "s/model", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def cpu_baseline_port():
    """Fallback when oracle/_ref is absent: the oracle (port of the reference's PyTorch ops) on the host cores, bounded sample:
    one 2048-token Hessian update at K=3072 and K=8192 and one full GPTQ solve of a [1024, 3072]
    slice, extrapolated to sec/model by the algorithmic counts."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as orc

    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    t_h = {}
    for K in (D_MODEL, D_FFN):
        X = orc.bf16_round(rng.standard_normal((SEQ_LEN, K), dtype=np.float32))
        H = np.zeros((K, K), np.float32)
        orc.hessian_accum(H, X, 0)
        t0 = time.perf_counter()
        orc.hessian_accum(H, X, 1)
        t_h[K] = time.perf_counter() - t0
    Ns, K = 1024, D_MODEL
    X = orc.bf16_round(rng.standard_normal((4096, K), dtype=np.float32))
    H = np.zeros((K, K), np.float32)
    orc.hessian_accum(H, X[:2048], 0)
    orc.hessian_accum(H, X[2048:], 1)
    W = orc.bf16_round(0.02 * rng.standard_normal((Ns, K), dtype=np.float32))
    t0 = time.perf_counter()
    Hc = H.copy()
    orc.damp_and_factor(Hc, 0.01)
    t_chol = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.gptq_update(W, H.copy(), WCFG)
    t_tot = time.perf_counter() - t0           # gptq_update = parameter search + factor + block loop
    t_upd = max(t_tot - t_chol, 0.05 * t_tot)
    # extrapolation: Hessians exactly (4 per layer, 128 samples); Cholesky ~ K^3; block loop ~ N*K^2
    hess = N_LAYERS * N_SAMPLES * (3 * t_h[D_MODEL] + t_h[D_FFN])
    chol = N_LAYERS * (3 * t_chol + t_chol * (D_FFN / D_MODEL) ** 3)
    upd = 0.0
    for Kg, lins in GROUPS:
        for _, N in lins:
            upd += t_upd * (N * Kg * Kg) / (Ns * K * K)
    upd *= N_LAYERS
    return {"value": hess + chol + upd, "unit": "s/model", "cores": cores, "kind": "port",
            "sample": "oracle: one 2048-token Hessian update each at K=3072/8192 (%.3f s / %.3f s), one Cholesky chain "
                      "K=3072 (%.2f s), one GPTQ block loop on [1024,3072] (%.2f s); extrapolated by 2TK^2, K^3, NK^2"
                      % (t_h[D_MODEL], t_h[D_FFN], t_chol, t_upd),
            "stages": {"hessian": hess, "cholesky": chol, "update": upd}}



if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    ap.add_argument("--no-fake-quant", action="store_true", help="skip the fake-quant bandwidth section")
    ap.add_argument("--hessian-defer", type=int, default=4,
                    help="hook inputs accumulated per Hessian kernel launch (solvers.HESSIAN_DEFER); 1 = one launch per call")
    ap.add_argument("--layers", type=int, default=N_LAYERS,
                    help="decoder layers per step (default: the full 28-layer model; smaller only for profiling runs, "
                         "the JSON line then says so in config.layers)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

"""torchrun script (N >= 2 GPUs, NCCL): row-sharded SparseGPT / RIA / magnitude against the unsharded calls.
Launched by tests/test_solvers_gpu.py::test_multi_gpu_sharded_pruning_script, or by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_sharded.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl")
    from llm_compressor_b200 import ops, parallel

    N, K = 768, 1024
    g = torch.Generator().manual_seed(11)
    W = (0.02 * torch.randn(N, K, generator=g)).to(torch.bfloat16).to(dev)
    X = (torch.randn(2048, K, generator=g) * torch.exp(0.5 * torch.randn(K, generator=g))).to(torch.bfloat16).to(dev)
    s = (X.float() ** 2).sum(0) / 8
    rows = parallel.row_shard(N)

    # calibration tokens sharded: raw sums + all-reduce == all tokens on one GPU (bit-level: fp32 adds re-associate)
    H = torch.zeros(K, K, device=dev)
    n = 0
    for j in parallel.sample_shard(8):
        n = ops.hessian_accum_raw(H, X[j * 256:(j + 1) * 256].unsqueeze(0), n)
    n = parallel.reduce_hessian_(H, n)
    assert n == 8
    ops.hessian_finalize(H, 2.0 / n, True)
    Hf = torch.zeros(K, K, device=dev)
    m = 0
    for j in range(8):
        m = ops.hessian_accum_raw(Hf, X[j * 256:(j + 1) * 256].unsqueeze(0), m)
    ops.hessian_finalize(Hf, 2.0 / m, True)
    assert float((H - Hf).norm() / Hf.norm()) < 1e-6

    # exact-fp32 solver GEMMs: one stream, fixed summation order -> the row shards reproduce the unsharded call bit
    # for bit.  tensor-core mode: trailing updates are L2 reduce-adds issued from look-ahead streams, their arrival
    # order (hence the last fp32 bit) is not fixed -> masks / values compared by agreement.
    for mode in (0, 1):
        prev = ops.set_gemm_mode(mode)
        U = ops.chol_inv_upper(Hf.clone(), percdamp=0.01)
        full = ops.sparsegpt_update(W.float().contiguous(), U, 0.5)
        part = parallel.sparsegpt_update_sharded(W[rows].float().contiguous(), U, 0.5)
        got = parallel.gather_rows(part, N)
        ops.set_gemm_mode(prev)
        if mode == 0:
            assert torch.equal(got, full), "sharded SparseGPT differs from the unsharded call (exact mode)"
        else:
            agree = float(((got == 0) == (full == 0)).float().mean())
            rel = float((got - full).norm() / full.norm())
            assert agree > 0.9999 and rel < 1e-3, (agree, rel)
        assert abs(float((got == 0).float().mean()) - 0.5) < 2e-3

    mm = parallel.mask_magnitude_sharded(W[rows].contiguous(), 0.5)
    assert torch.equal(mm, ops.mask_magnitude(W, 0.5)[rows])
    mr = parallel.mask_ria_sharded(W[rows].contiguous(), s, 0.5, 0.5)
    agree = float((mr == ops.mask_ria(W, s, 0.5, 0.5)[rows]).float().mean())
    assert agree > 0.9995, agree
    dist.barrier()
    if rank == 0:
        print("mgpu sharded ok: world", world, "ria agreement", agree)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

# round-2 profiles: launch list of one layer of the bench + full captures of the new / dominant kernels.
# Every ncu run follows a plain run of the same command (B200_PROFILING.md).
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > $O/r2_bench_layers1.json 2> $O/r2_bench_layers1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > /dev/null 2>&1
python scripts/summarize_launches.py $O/r2_launches.csv > $O/r2_launches_summary.txt; rm -f $O/r2_launches.csv
bash scripts/prof_one.sh r2_chol_tiles_k3072 chol_tiles_tc 1 1 python scripts/chol_once.py 3072
bash scripts/prof_one.sh r2_chol_tiles_k8192 chol_tiles_tc 1 1 python scripts/chol_once.py 8192
bash scripts/prof_one.sh r2_hessian_k3072 hessian_umma 2 2 python scripts/hess_once.py 3072 8 4
bash scripts/prof_one.sh r2_hessian_k8192 hessian_umma 2 2 python scripts/hess_once.py 8192 8 4
bash scripts/prof_one.sh r2_qdq_tensor tensor_apply_stream 1 1 python scripts/qdq_once.py int8_tensor
bash scripts/prof_one.sh r2_qdq_channel colgroup_apply_stream 1 1 python scripts/qdq_once.py int8_channel
bash scripts/prof_one.sh r2_qdq_f32 qdq_stream_f32 1 1 python scripts/qdq_once.py int4_g128_fp32
bash scripts/prof_one.sh r2_qdq_nvfp4 qdq_stream_kernel 1 1 python scripts/qdq_once.py nvfp4
bash scripts/prof_one.sh r2_qdq_mxfp4 qdq_stream_kernel 1 1 python scripts/qdq_once.py mxfp4

# round-2 final profiles (after the column-group task order and the 128-column GEMM tile): launch list of one layer of the
# bench and the K = 8192 tile-task Cholesky.  Every ncu run follows a plain run of the same command (B200_PROFILING.md).
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > $O/r2c_bench_layers1.json 2> $O/r2c_bench_layers1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c_launches.csv python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > /dev/null 2>&1
python scripts/summarize_launches.py $O/r2c_launches.csv > $O/r2c_launches_summary.txt; rm -f $O/r2c_launches.csv
python scripts/chol_once.py 8192 && bash scripts/prof_one.sh r2c_chol_tiles_k8192 chol_tiles_tc 1 1 python scripts/chol_once.py 8192

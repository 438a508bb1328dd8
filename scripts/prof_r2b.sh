# round-2 (second session) profiles: launch list of one layer of the bench + full captures of the kernels changed since
# prof_r2.sh (tile-task Cholesky with the split-role potf2, register-resident Hadamard kernel).
# Every ncu run follows a plain run of the same command (B200_PROFILING.md).
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > $O/r2b_bench_layers1.json 2> $O/r2b_bench_layers1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2b_launches.csv python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > /dev/null 2>&1
python scripts/summarize_launches.py $O/r2b_launches.csv > $O/r2b_launches_summary.txt; rm -f $O/r2b_launches.csv
python scripts/chol_once.py 3072 && bash scripts/prof_one.sh r2b_chol_tiles_k3072 chol_tiles_tc 1 1 python scripts/chol_once.py 3072
python scripts/chol_once.py 8192 && bash scripts/prof_one.sh r2b_chol_tiles_k8192 chol_tiles_tc 1 1 python scripts/chol_once.py 8192
python scripts/had_once.py 65536 3072 bf16 32 && bash scripts/prof_one.sh r2b_hadamard_reg_3072 hadamard_reg 2 1 python scripts/had_once.py 65536 3072 bf16 32
python scripts/had_once.py 65536 8192 bf16 32 && bash scripts/prof_one.sh r2b_hadamard_reg_8192 hadamard_reg 2 1 python scripts/had_once.py 65536 8192 bf16 32

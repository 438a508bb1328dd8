# round-1 v11 profiles: launch list of one layer + the reworked Cholesky panel kernel
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > $O/r1v11_bench_layers1.json 2> $O/r1v11_bench_layers1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1v11_launches.csv python bench.py --steps 1 --warmup 0 --layers 1 --no-cpu-baseline --no-fake-quant > /dev/null 2>&1
python scripts/summarize_launches.py $O/r1v11_launches.csv > $O/r1v11_launches_summary.txt; rm -f $O/r1v11_launches.csv
bash scripts/prof_one.sh r1v11_chol_panel chol_panel 70 2 python scripts/chol_once.py 8192

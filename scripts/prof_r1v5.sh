# round-1 final profiles.  Reports are summarised ON the box and deleted (gpurun_out is capped at 64 MiB).
set -x
O=gpurun_out
sum() {  # $1 = report stem
  python scripts/ncu_summary.py $O/$1.ncu-rep > $O/$1_summary.txt 2>&1
  ncu -i $O/$1.ncu-rep --page details > $O/$1_details.txt 2>&1
  rm -f $O/$1.ncu-rep
}
python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > $O/r1v5_bench_layers1.json 2> $O/r1v5_bench_layers1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1v5_launches.csv python bench.py --steps 1 --warmup 0 --layers 1 --no-cpu-baseline --no-fake-quant > /dev/null 2>&1
python scripts/summarize_launches.py $O/r1v5_launches.csv > $O/r1v5_launches_summary.txt; rm -f $O/r1v5_launches.csv
for K in 8192 3072; do
  timeout 300 ncu --set full --clock-control none -k regex:"hessian_umma" --launch-skip 10 --launch-count 2 -f -o $O/r1v5_hessian_k$K python scripts/hess_once.py $K > /dev/null 2>&1
  sum r1v5_hessian_k$K
done
timeout 300 ncu --set full --clock-control none -k regex:"tgemm" --launch-skip 40 --launch-count 3 -f -o $O/r1v5_tgemm python scripts/update_once.py 3072 8192 1 > /dev/null 2>&1
sum r1v5_tgemm
timeout 300 ncu --set full --clock-control none -k regex:"chol_panel" --launch-skip 70 --launch-count 2 -f -o $O/r1v5_chol_panel python scripts/chol_once.py 8192 > /dev/null 2>&1
sum r1v5_chol_panel
for c in int4_g128_zp int8_tok mxfp4 nvfp4; do
  timeout 300 ncu --set full --clock-control none -k regex:"qdq|nvfp" --launch-skip 6 --launch-count 2 -f -o $O/r1v5_qdq_$c python scripts/qdq_bw.py $c > /dev/null 2>&1
  sum r1v5_qdq_$c
done
ls -la $O

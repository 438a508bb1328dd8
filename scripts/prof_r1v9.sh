# round-1 v9 profiles (deferred lock-step Hessian, Hadamard, pack).  Reports are summarised ON the box and deleted.
set -x
O=gpurun_out
sum() {  # $1 = report stem
  python scripts/ncu_summary.py $O/$1.ncu-rep > $O/$1_summary.txt 2>&1
  ncu -i $O/$1.ncu-rep --page details > $O/$1_details.txt 2>&1
  rm -f $O/$1.ncu-rep
}
python bench.py --steps 1 --warmup 1 --layers 1 --no-cpu-baseline --no-fake-quant > $O/r1v9_bench_layers1.json 2> $O/r1v9_bench_layers1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1v9_launches.csv python bench.py --steps 1 --warmup 0 --layers 1 --no-cpu-baseline --no-fake-quant > /dev/null 2>&1
python scripts/summarize_launches.py $O/r1v9_launches.csv > $O/r1v9_launches_summary.txt; rm -f $O/r1v9_launches.csv
for K in 8192 3072; do
  python scripts/hess_once.py $K 8 4 > $O/r1v9_hess_once_k$K.txt 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:"hessian_umma" --launch-skip 2 --launch-count 1 -f -o $O/r1v9_hessian_k$K python scripts/hess_once.py $K 8 4 > /dev/null 2>&1
  sum r1v9_hessian_k$K
done
timeout 300 ncu --set full --clock-control none -k regex:"hadamard_tile" --launch-skip 3 --launch-count 1 -f -o $O/r1v9_hadamard_3072_acc64 python scripts/had_once.py 8192 3072 bf16 64 > /dev/null 2>&1
sum r1v9_hadamard_3072_acc64
timeout 300 ncu --set full --clock-control none -k regex:"hadamard_tile" --launch-skip 3 --launch-count 1 -f -o $O/r1v9_hadamard_8192_acc32 python scripts/had_once.py 3072 8192 bf16 32 > /dev/null 2>&1
sum r1v9_hadamard_8192_acc32
ls -la $O | tail -20

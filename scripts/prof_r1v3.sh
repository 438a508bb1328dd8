set -x
for c in int4_g128_zp mxfp4 nvfp4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"qdq|nvfp" --launch-skip 4 --launch-count 2 -f -o gpurun_out/r1v3_qdq_$c python scripts/qdq_bw.py $c > gpurun_out/r1v3_ncu_qdq_$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hessian_umma" --launch-skip 300 --launch-count 2 -f -o gpurun_out/r1v3_hessian python bench.py --steps 1 --warmup 0 --layers 1 --no-cpu-baseline --no-fake-quant > gpurun_out/r1v3_ncu_hessian.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tgemm" --launch-skip 40 --launch-count 3 -f -o gpurun_out/r1v3_tgemm python scripts/update_once.py 3072 8192 1 > gpurun_out/r1v3_ncu_tgemm.log 2>&1
ls -la gpurun_out/*.ncu-rep

# usage: bash scripts/prof_qdq.sh tag case...   -> gpurun_out/<tag>_qdq_<case>.ncu-rep  (one launch, no source: small reports)
tag=$1; shift
for c in "$@"; do
  timeout 300 ncu --set full --clock-control none -k regex:"qdq|nvfp" --launch-skip 4 --launch-count 1 -f -o gpurun_out/${tag}_qdq_$c python scripts/qdq_bw.py $c > gpurun_out/${tag}_ncu_qdq_$c.log 2>&1
done
ls -la gpurun_out

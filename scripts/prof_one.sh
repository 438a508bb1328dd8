# usage: prof_one.sh <stem> <kernel-regex> <skip> <count> <cmd...>  -> gpurun_out/<stem>_{summary,details}.txt (report deleted)
stem=$1; regex=$2; skip=$3; count=$4; shift 4
timeout 300 ncu --set full --clock-control none -k regex:"$regex" --launch-skip $skip --launch-count $count -f -o gpurun_out/$stem "$@" > gpurun_out/${stem}_run.log 2>&1
python scripts/ncu_summary.py gpurun_out/$stem.ncu-rep > gpurun_out/${stem}_summary.txt 2>&1
ncu -i gpurun_out/$stem.ncu-rep --page details > gpurun_out/${stem}_details.txt 2>&1
rm -f gpurun_out/$stem.ncu-rep

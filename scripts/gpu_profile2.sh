#!/bin/bash
# ncu --set full of selected kernels of one call.  usage: bash scripts/gpu_profile2.sh <tag> <kind> <kernel-regex> <count> [skip]
set -u
TAG=$1; K=$2; RE=$3; CNT=$4; SKIP=${5:-$4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE --launch-skip $SKIP -c $CNT -f -o $OUT/full_$K \
    python profiles/run_kernels.py $K 3 1 > $OUT/ncu_full_$K.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/full_$K.ncu-rep --page raw --csv > $OUT/full_${K}_raw.csv 2>/dev/null
ncu -i $OUT/full_$K.ncu-rep --page source --csv > $OUT/full_${K}_source.csv 2>/dev/null
python profiles/summarize_ncu.py $OUT/full_$K.ncu-rep > $OUT/full_${K}_summary.txt 2>&1
sz=$(stat -c %s $OUT/full_$K.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f $OUT/full_$K.ncu-rep; fi
cat $OUT/full_${K}_summary.txt

#!/bin/bash
# ncu --set full of selected kernels of one call; keeps gzip'd CSV exports (raw + per-SASS source page) and a summary,
# drops the .ncu-rep (too big for the 64 MiB gpurun_out limit).
# usage: bash scripts/gpu_profile2.sh <tag> <kind> <kernel-regex> <count> [skip]
set -u
TAG=$1; K=$2; RE=$3; CNT=$4; SKIP=${5:-$4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE --launch-skip $SKIP -c $CNT -f -o /tmp/full_$K \
    python profiles/run_kernels.py $K 3 1 > $OUT/ncu_full_$K.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/full_$K.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/full_${K}_raw.csv.gz
ncu -i /tmp/full_$K.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/full_${K}_source.csv.gz
python profiles/summarize_ncu.py /tmp/full_$K.ncu-rep > $OUT/full_${K}_summary.txt 2>&1
cat $OUT/full_${K}_summary.txt
du -sh $OUT

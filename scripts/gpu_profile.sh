#!/bin/bash
# ncu captures of one stage-2 DFA call per modality (forward + the backward kernels), second repetition
# (first one warms the instruction cache / smem attributes).  Keeps CSV exports; drops reports > 12 MiB.
# usage: bash scripts/gpu_profile.sh <tag> [kinds...]
set -u
TAG=${1:-prof}; shift
KINDS=${@:-det map plan}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for k in $KINDS; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dfa_ --launch-skip 5 -c 5 -f -o $OUT/full_$k \
      python profiles/run_kernels.py $k 3 1 > $OUT/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?" | tee -a $OUT/status.txt
  ncu -i $OUT/full_$k.ncu-rep --page raw --csv > $OUT/full_${k}_raw.csv 2>/dev/null
  python profiles/summarize_ncu.py $OUT/full_$k.ncu-rep > $OUT/full_${k}_summary.txt 2>&1
  sz=$(stat -c %s $OUT/full_$k.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12582912 ]; then
    ncu -i $OUT/full_$k.ncu-rep --page source --csv > $OUT/full_${k}_source.csv 2>/dev/null
    rm -f $OUT/full_$k.ncu-rep; echo "dropped full_$k.ncu-rep ($sz bytes)" | tee -a $OUT/status.txt
  fi
done
du -sh $OUT
cat $OUT/full_*_summary.txt

#!/bin/bash
# A/B of an environment knob on the per-stage timings.  usage: bash scripts/gpu_ab.sh <tag> <kinds> VAR v1 v2 ...
set -u
TAG=$1; KINDS=$2; VAR=$3; shift; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  echo "== $VAR=$v" | tee -a $OUT/ab.txt
  for k in $KINDS; do env $VAR=$v timeout 120 python profiles/run_kernels.py $k 4 1 2>&1 | tail -1 | tee -a $OUT/ab.txt; done
done

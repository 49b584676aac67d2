#!/bin/bash
# A/B of an environment knob on the per-stage timings.  usage: bash scripts/gpu_ab.sh <tag> VAR v1 v2 ...
set -u
TAG=$1; VAR=$2; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  echo "== $VAR=$v" | tee -a $OUT/ab.txt
  for k in det map plan; do env $VAR=$v timeout 120 python profiles/run_kernels.py $k 4 1 2>&1 | tail -1 | tee -a $OUT/ab.txt; done
  env $VAR=$v timeout 120 python profiles/run_kernels.py det 3 4 2>&1 | tail -1 | tee -a $OUT/ab.txt
  env $VAR=$v timeout 120 python profiles/run_kernels.py map 3 4 2>&1 | tail -1 | tee -a $OUT/ab.txt
done

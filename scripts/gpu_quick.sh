#!/bin/bash
# quick correctness + per-kernel timing.  usage: bash scripts/gpu_quick.sh <tag>
set -u
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -p timeout --timeout 60 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
tail -15 $OUT/pytest_gpu.log
for k in det map plan ego; do timeout 120 python profiles/run_kernels.py $k 4 1 2>&1 | tail -2 | tee -a $OUT/kernels.txt; done
timeout 120 python profiles/run_kernels.py det 3 4 2>&1 | tail -1 | tee -a $OUT/kernels.txt
timeout 120 python profiles/run_kernels.py map 3 4 2>&1 | tail -1 | tee -a $OUT/kernels.txt
for k in det plan; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dfa_ --csv --log-file $OUT/launches_$k.csv \
     python profiles/run_kernels.py $k 3 1 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches_$k.csv")) if len(r)>5]
h={n:i for i,n in enumerate(rows[0])}
names=[(r[h["Kernel Name"]].split("(")[0][-40:], float(r[h["Metric Value"]])) for r in rows[1:]]
n=len(names)//3
print("$k per-kernel us (last rep):", ", ".join("%s=%.1f"%(a,b) for a,b in names[-n:]))
PY
done
for a in "1 f32" "1 bf16" "4 f32"; do timeout 120 python profiles/run_format.py $a 2>&1 | tail -1 | tee -a $OUT/kernels.txt; done

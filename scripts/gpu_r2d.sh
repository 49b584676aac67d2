#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2d}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest_group.log 2>&1; echo "group pytest rc=$?" | tee -a $OUT/status.txt; tail -3 $OUT/pytest_group.log
HIPAD_DFA_GROUP_DEEP=0 timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest_group_shallow.log 2>&1; echo "group pytest shallow rc=$?" | tee -a $OUT/status.txt; tail -3 $OUT/pytest_group_shallow.log
for D in 1 0; do
  HIPAD_DFA_GROUP_DEEP=$D timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_deep$D.json 2> $OUT/group_deep$D.err; echo "run_group deep=$D rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_deep$D.json
done
HIPAD_DFA_BANDS=12 timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_bands12.json 2> $OUT/group_bands12.err; echo "run_group bands12 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_bands12.json
HIPAD_DFA_GROUP_DEEP=0 timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_bs4_deep0.json 2> $OUT/group_bs4.err; echo "run_group bs4 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_bs4_deep0.json

#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2e}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt; tail -3 $OUT/pytest.log
timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_f32.json 2> $OUT/group_f32.err; echo "run_group rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_f32.json
timeout 300 python profiles/run_group.py 1 bf16 > $OUT/group_bf16.json 2> $OUT/group_bf16.err; echo "run_group bf16 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_bf16.json
timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_bs4.json 2> $OUT/group_bs4.err; echo "run_group bs4 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_bs4.json

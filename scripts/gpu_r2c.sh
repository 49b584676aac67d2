#!/bin/bash
# round-2: quad-merging grouped kernels — parity first, then timings
set -u
OUT=gpurun_out/${1:-r2c}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest_group.log 2>&1; echo "group pytest rc=$?" | tee -a $OUT/status.txt; tail -12 $OUT/pytest_group.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -p timeout --timeout 600 > $OUT/pytest_parity.log 2>&1; echo "parity pytest rc=$?" | tee -a $OUT/status.txt; tail -6 $OUT/pytest_parity.log
timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_f32.json 2> $OUT/group_f32.err; echo "run_group rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_f32.json
timeout 300 python profiles/run_group.py 1 bf16 > $OUT/group_bf16.json 2> $OUT/group_bf16.err; echo "run_group bf16 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_bf16.json
timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_bs4.json 2> $OUT/group_bs4.err; echo "run_group bs4 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_bs4.json
HIPAD_DFA_GROUP_PS_FWD=48 HIPAD_DFA_GROUP_PS_BWD=48 timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_ps48.json 2> $OUT/group_ps48.err; echo "run_group ps48 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_ps48.json
timeout 600 python harness/parity_report.py > $OUT/decoder_parity.json 2> $OUT/decoder_parity.err; echo "decoder parity rc=$?" | tee -a $OUT/status.txt; cat $OUT/decoder_parity.json
timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt; cut -c1-300 $OUT/bench.json

#!/bin/bash
# round-2: new grouped kernels — GPU tests first (parity), then per-stage timings under a few knob settings
set -u
OUT=gpurun_out/${1:-r2b}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest_group.log 2>&1; echo "group pytest rc=$?" | tee -a $OUT/status.txt; tail -12 $OUT/pytest_group.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_decoder_harness.py -m gpu -q -p timeout --timeout 600 > $OUT/pytest_rest.log 2>&1; echo "rest pytest rc=$?" | tee -a $OUT/status.txt; tail -12 $OUT/pytest_rest.log
for W in 4 8; do
  HIPAD_DFA_GROUP_WARPS=$W timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_w${W}.json 2> $OUT/group_w${W}.err; echo "run_group w=$W rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_w${W}.json
done
HIPAD_DFA_GROUP_KERNEL=0 timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_old.json 2> $OUT/group_old.err; echo "run_group old rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_old.json
HIPAD_DFA_GROUP_PS_FWD=64 HIPAD_DFA_GROUP_PS_BWD=32 timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_ps64_32.json 2> $OUT/group_ps.err; echo "run_group ps rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_ps64_32.json
timeout 300 python profiles/run_group.py 1 bf16 > $OUT/group_bf16.json 2> $OUT/group_bf16.err; echo "run_group bf16 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_bf16.json
timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_bs4.json 2> $OUT/group_bs4.err; echo "run_group bs4 rc=$?" | tee -a $OUT/status.txt; cat $OUT/group_bs4.json
timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt; cut -c1-400 $OUT/bench.json

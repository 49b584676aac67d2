#!/bin/bash
# round-2: decoder harness tests, stage breakdown, new bench.py (all legs)
set -u
OUT=gpurun_out/${1:-r2f}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_decoder_harness.py -m gpu -x -q -p timeout --timeout 600 > $OUT/pytest_harness.log 2>&1; echo "harness pytest rc=$?" | tee -a $OUT/status.txt; tail -8 $OUT/pytest_harness.log
timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_f32.json 2> $OUT/group_f32.err; echo "run_group rc=$?" | tee -a $OUT/status.txt; python -c "import json;print(json.load(open('$OUT/group_f32.json'))['layer'])"
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt; tail -5 $OUT/bench.err; cat $OUT/bench.json

#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2i}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt; tail -15 $OUT/pytest.log
timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_f32.json 2> $OUT/group_f32.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/group_f32.json'));print({k:(d[k]['fwd_group1_us'],d[k]['bwd_group1_us']) for k in ('det','map','plan','ego')}, d['layer'])"

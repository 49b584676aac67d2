#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2o}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt; tail -3 $OUT/pytest.log
for cfg in "1 f32" "4 f32" "1 bf16"; do
n=$(echo $cfg | tr ' ' '_')
timeout 300 python profiles/run_group.py $cfg > $OUT/group_$n.json 2> $OUT/group_$n.err; echo "$cfg rc=$?"; python -c "import json;d=json.load(open('$OUT/group_$n.json'));print({k:(d[k]['fwd_group1_us'],d[k]['bwd_group1_us']) for k in ('det','map','plan','ego')}, d['layer']['fwd_grouped_us'], d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
done
